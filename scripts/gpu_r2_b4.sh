#!/bin/bash
# call B4: per-CTA timelines (trace build) of res5 / res4 branch2b / res5 2a(1x1) forward under the tiling variants
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export VLFB_LIB=$PWD/video-long-term-feature-banks_b200/csrc/libvlfb_trace.so
for layer in res5_2b res4_2b res5_2a1; do
  rm -f $O/r2b_trace_$layer.txt
  for v in "-1 -1" "1 -1" "-1 1" "1 1" "-1 -1 192" "-1 -1 128"; do
    timeout 120 python scripts/trace_gemm.py $layer fwd $v >> $O/r2b_trace_$layer.txt 2>&1
  done
done
