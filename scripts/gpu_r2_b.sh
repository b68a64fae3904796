#!/bin/bash
# Round 2, call B: why are the CTA-pair / stream-K launches slower than the round-1 tiling?  Device-time table of the
# production shapes per variant + ncu --set full of res5 / res4 branch2b under three variants.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
VLFB_DEBUG=1 timeout 600 python scripts/bench_gemm_shapes.py > $O/r2b_shapes.txt 2> $O/r2b_shapes.err
echo "shapes rc=$?"
grep "vlfb gemm_tc" $O/r2b_shapes.err | sort | uniq -c | sort -rn | head -150 > $O/r2b_plans.txt
for v in "off -1 -1" "pair 1 -1" "pairsk 1 1" "sk -1 1"; do
  set -- $v
  for layer in res5_2b res4_2b; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o $O/r2b_${layer}_$1 \
      python scripts/prof_pair.py $layer fwd $2 $3 > $O/r2b_ncu_${layer}_$1.log 2>&1
    echo "ncu $layer $1 rc=$?"
  done
done
cat $O/r2b_shapes.txt
