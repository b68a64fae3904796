#!/bin/bash
# call C: lean epilogue -- kernel tests, timelines, variant table, bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2c_kernels.log 2>&1; echo "kernels rc=$?"
export VLFB_LIB=$PWD/video-long-term-feature-banks_b200/csrc/libvlfb_trace.so
for layer in res5_2b res4_2b; do
  rm -f $O/r2c_trace_$layer.txt
  for v in "-1 -1" "-1 1" "1 1"; do
    timeout 120 python scripts/trace_gemm.py $layer fwd $v >> $O/r2c_trace_$layer.txt 2>&1
  done
done
unset VLFB_LIB
REPS=50 timeout 600 python scripts/bench_gemm_shapes.py > $O/r2c_shapes.txt 2>&1
VLFB_PAIR=-1 VLFB_SK=-1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms $O/r2c_gemm_table_off.txt > $O/r2c_bench_off.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms $O/r2c_gemm_table_auto.txt > $O/r2c_bench_auto.log 2>&1
tail -3 $O/r2c_kernels.log
cat $O/r2c_shapes.txt
