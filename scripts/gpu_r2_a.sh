#!/bin/bash
# Round 2, call A: bring-up of the CTA-pair (cta_group::2) + stream-K GEMM.  Each group runs in its own process under its
# own timeout so that a trap / hang in one variant does not hide the others.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
run() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  timeout $t "$@" > $O/r2a_$name.log 2>&1
  echo "$name rc=$?" >> $O/r2a_status.txt
}
rm -f $O/r2a_status.txt
# 1. refactor only (no pairs, no stream-K): the round-1 kernel set on the new 10-warp TMA-only / 17-warp builds
VLFB_PAIR=-1 VLFB_SK=-1 run base_kernels 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "not tiling_variants and not deterministic"
# 2. variants, one process each
run var_conv 900 python -m pytest tests/test_gpu_kernels.py -q -k "production_conv_shapes"
run var_matmul 600 python -m pytest tests/test_gpu_kernels.py -q -k "production_matmul_shapes or deterministic"
# 3. whole kernel suite with the default (auto) plan, then the model tests
run auto_kernels 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "not tiling_variants and not deterministic"
run model 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_kernels.py
# 4. bench A/B
VLFB_PAIR=-1 VLFB_SK=-1 run bench_off 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms $O/r2a_gemm_table_off.txt
run bench_auto 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms $O/r2a_gemm_table_auto.txt
cat $O/r2a_status.txt
tail -3 $O/r2a_base_kernels.log $O/r2a_var_conv.log $O/r2a_var_matmul.log $O/r2a_auto_kernels.log $O/r2a_model.log
tail -2 $O/r2a_bench_off.log $O/r2a_bench_auto.log
