#!/bin/bash
# call F: conv1 patch tiles (fwd: 16 x 8 output patch per box; wgrad: one box of sH + kH input rows per chunk):
# kernel tests, conv1 microbench with the plan print, model tests (full-size configs), bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2f_kernels.log 2>&1; echo "kernels rc=$?"
VLFB_DEBUG=1 timeout 200 python scripts/bench_gemm_shapes.py conv1 > $O/r2f_conv1.txt 2> $O/r2f_conv1.err; echo "conv1 rc=$?"
VLFB_NO_PATCH=1 timeout 200 python scripts/bench_gemm_shapes.py conv1 > $O/r2f_conv1_nopatch.txt 2>/dev/null
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r2f_model.log 2>&1; echo "model rc=$?"
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --dump-gemms $O/r2f_gemm_table.txt > $O/r2f_bench.log 2>&1
VLFB_NO_PATCH=1 timeout 300 $B > $O/r2f_bench_nopatch.log 2>&1
for f in bench bench_nopatch; do echo "== $f"; tail -1 $O/r2f_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
tail -n 3 $O/r2f_kernels.log; tail -n 3 $O/r2f_model.log
cat $O/r2f_conv1.txt; sort -u $O/r2f_conv1.err | head -20; cat $O/r2f_conv1_nopatch.txt
