#!/bin/bash
# call B3: timelines (trace build) + ncu full (CSV exported on the box) of res5 / res4 branch2b under the tiling variants
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export VLFB_LIB=$PWD/video-long-term-feature-banks_b200/csrc/libvlfb_trace.so
for layer in res5_2b res4_2b res5_2a1; do
  for v in "-1 -1" "1 -1" "-1 1" "1 1" "-1 -1 192" "-1 -1 128"; do
    timeout 120 python scripts/trace_gemm.py $layer fwd $v >> $O/r2b_trace_$layer.txt 2>&1
  done
done
unset VLFB_LIB
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 100 > $O/r2b_clocks.csv &
SMI=$!
REPS=200 timeout 300 python scripts/bench_gemm_shapes.py res5_2b res4_2b > $O/r2b_shapes_long.txt 2>&1
kill $SMI
for v in "off -1 -1" "pair 1 -1" "sk -1 1"; do
  set -- $v
  for layer in res5_2b; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o /tmp/rep_${layer}_$1 \
      python scripts/prof_pair.py $layer fwd $2 $3 > $O/r2b_ncu_${layer}_$1.log 2>&1
    echo "ncu $layer $1 rc=$?"
    ncu -i /tmp/rep_${layer}_$1.ncu-rep --page raw --csv > $O/r2b_${layer}_$1_raw.csv 2>/dev/null
    ncu -i /tmp/rep_${layer}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r2b_${layer}_$1_source.csv.gz
  done
done
du -sh $O
