#!/bin/bash
# call I: compile-time bit emission variant, rounded gradients stored by their producers (no copy + round passes)
# depth 1 variant library)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2i_kernels.log 2>&1; echo "kernels rc=$?"
timeout 200 python scripts/bench_gemm_shapes.py epi strided > $O/r2i_epi.txt 2> $O/r2i_epi.err; echo "epi rc=$?"
VLFB_LIB=$CS/libvlfb_d1.so timeout 200 python scripts/bench_gemm_shapes.py epi > $O/r2i_epi_d1.txt 2>/dev/null
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r2i_model.log 2>&1; echo "model rc=$?"
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --dump-gemms $O/r2i_gemm_table.txt > $O/r2i_bench.log 2>&1
true
VLFB_LIB=$CS/libvlfb_d1.so timeout 300 $B > $O/r2i_bench_d1.log 2>&1
export VLFB_LIB=$CS/libvlfb_trace.so
rm -f $O/r2i_trace.txt
for l in "res2_2c fwd" "res2_2c dgrad"; do timeout 120 python scripts/trace_gemm.py $l -1 -1 >> $O/r2i_trace.txt 2>&1; done
unset VLFB_LIB
for f in bench bench_d1; do echo "== $f"; tail -1 $O/r2i_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
tail -n 3 $O/r2i_kernels.log; tail -n 3 $O/r2i_model.log
cat $O/r2i_epi.txt; echo "-- depth 1"; cat $O/r2i_epi_d1.txt; grep -v slowest $O/r2i_trace.txt | grep -E "^#|item[01] " | head -40
