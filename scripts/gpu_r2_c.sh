#!/bin/bash
# call C: lean epilogue + FBO-NL stack -- tests, timelines, variant table, bench A/Bs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2c_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_kernels.py > $O/r2c_model.log 2>&1; echo "model rc=$?"
export VLFB_LIB=$PWD/video-long-term-feature-banks_b200/csrc/libvlfb_trace.so
for layer in res5_2b res4_2b; do
  rm -f $O/r2c_trace_$layer.txt
  for v in "-1 -1" "-1 1" "1 1"; do
    timeout 120 python scripts/trace_gemm.py $layer fwd $v >> $O/r2c_trace_$layer.txt 2>&1
  done
done
unset VLFB_LIB
REPS=50 timeout 600 python scripts/bench_gemm_shapes.py > $O/r2c_shapes.txt 2>&1
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
VLFB_PAIR=-1 VLFB_SK=-1 timeout 300 $B --dump-gemms $O/r2c_gemm_table_off.txt > $O/r2c_bench_off.log 2>&1
timeout 300 $B --dump-gemms $O/r2c_gemm_table_auto.txt > $O/r2c_bench_auto.log 2>&1
VLFB_PAIR=-1 VLFB_SK=-1 VLFB_FUSE_GRAD_FINISH=1 timeout 300 $B --dump-gemms $O/r2c_gemm_table_off_fuse.txt > $O/r2c_bench_off_fuse.log 2>&1
VLFB_PAIR=-1 VLFB_SK=-1 VLFB_FBO_STACK=0 timeout 300 $B > $O/r2c_bench_off_nostack.log 2>&1
for f in off auto off_fuse off_nostack; do echo "== $f"; tail -1 $O/r2c_bench_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
tail -3 $O/r2c_kernels.log $O/r2c_model.log
cat $O/r2c_shapes.txt
