#!/bin/bash
# call H: ReLU sign bits (producer epilogue + consumer dgrad mask), look-ahead ring of the lean epilogue (depth 2 default,
# depth 1 variant library)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2h_kernels.log 2>&1; echo "kernels rc=$?"
timeout 200 python scripts/bench_gemm_shapes.py epi strided > $O/r2h_epi.txt 2> $O/r2h_epi.err; echo "epi rc=$?"
VLFB_LIB=$CS/libvlfb_d1.so timeout 200 python scripts/bench_gemm_shapes.py epi > $O/r2h_epi_d1.txt 2>/dev/null
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r2h_model.log 2>&1; echo "model rc=$?"
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --dump-gemms $O/r2h_gemm_table.txt > $O/r2h_bench.log 2>&1
VLFB_RELU_BITS=0 timeout 300 $B > $O/r2h_bench_nobits.log 2>&1
VLFB_LIB=$CS/libvlfb_d1.so timeout 300 $B > $O/r2h_bench_d1.log 2>&1
for f in bench bench_nobits bench_d1; do echo "== $f"; tail -1 $O/r2h_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
tail -n 3 $O/r2h_kernels.log; tail -n 3 $O/r2h_model.log
cat $O/r2h_epi.txt; echo "-- depth 1"; cat $O/r2h_epi_d1.txt
