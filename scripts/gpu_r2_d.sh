#!/bin/bash
# call D: conv1 TMA path, FBO-NL stack v2 (deep loads), graph-key fix -- tests, timelines, bench A/Bs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2d_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_kernels.py > $O/r2d_model.log 2>&1; echo "model rc=$?"
export VLFB_LIB=$PWD/video-long-term-feature-banks_b200/csrc/libvlfb_trace.so
for layer in res5_2b res4_2b; do
  rm -f $O/r2d_trace_$layer.txt
  for v in "-1 -1" "-1 1"; do
    timeout 120 python scripts/trace_gemm.py $layer fwd $v >> $O/r2d_trace_$layer.txt 2>&1
  done
done
unset VLFB_LIB
VLFB_DEBUG=1 REPS=20 timeout 300 python scripts/bench_gemm_shapes.py conv1 > $O/r2d_conv1.txt 2> $O/r2d_conv1.err
grep "vlfb gemm_tc" $O/r2d_conv1.err | sort | uniq -c > $O/r2d_conv1_plans.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --dump-gemms $O/r2d_gemm_table.txt > $O/r2d_bench.log 2>&1
VLFB_FUSE_GRAD_FINISH=1 timeout 300 $B --dump-gemms $O/r2d_gemm_table_fuse.txt > $O/r2d_bench_fuse.log 2>&1
VLFB_FBO_STACK=0 timeout 300 $B > $O/r2d_bench_nostack.log 2>&1
for f in bench bench_fuse bench_nostack; do echo "== $f"; tail -1 $O/r2d_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
tail -n 3 $O/r2d_kernels.log $O/r2d_model.log
cat $O/r2d_conv1.txt $O/r2d_conv1_plans.txt
