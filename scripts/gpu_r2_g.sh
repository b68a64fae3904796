#!/bin/bash
# call G: strided dgrads as parity classes, epilogue prefetch depth 2 (variant library), checkpoint test tolerance
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2g_kernels.log 2>&1; echo "kernels rc=$?"
VLFB_DEBUG=1 timeout 200 python scripts/bench_gemm_shapes.py strided epi > $O/r2g_strided.txt 2> $O/r2g_strided.err; echo "strided rc=$?"
VLFB_NO_PAR=1 timeout 200 python scripts/bench_gemm_shapes.py strided > $O/r2g_strided_nopar.txt 2>/dev/null
VLFB_LIB=$CS/libvlfb_d2.so timeout 200 python scripts/bench_gemm_shapes.py epi > $O/r2g_epi_d2.txt 2>/dev/null
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r2g_model.log 2>&1; echo "model rc=$?"
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --dump-gemms $O/r2g_gemm_table.txt > $O/r2g_bench.log 2>&1
VLFB_NO_PAR=1 timeout 300 $B > $O/r2g_bench_nopar.log 2>&1
VLFB_LIB=$CS/libvlfb_d2.so timeout 300 $B --dump-gemms $O/r2g_gemm_table_d2.txt > $O/r2g_bench_d2.log 2>&1
for f in bench bench_nopar bench_d2; do echo "== $f"; tail -1 $O/r2g_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
tail -n 3 $O/r2g_kernels.log; tail -n 3 $O/r2g_model.log
cat $O/r2g_strided.txt; grep "kinds 3" $O/r2g_strided.err | sort -u | head; echo "-- no par"; cat $O/r2g_strided_nopar.txt; echo "-- depth 2"; cat $O/r2g_epi_d2.txt
