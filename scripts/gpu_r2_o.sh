#!/bin/bash
# call O: where the e2e loop loses ~1.3 ms per step (copies / conversions / host), more bf16 scan occupancy points
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
timeout 300 python scripts/diag_e2e.py > $O/r2o_diag_e2e.txt 2>&1; echo "diag rc=$?"; grep -v Warning $O/r2o_diag_e2e.txt | tail -8
for v in default scan_r2b6 scan_r3b5 scan_r4b5; do
  L=$CS/libvlfb_$v.so; [ $v = default ] && L=$CS/libvlfb.so
  echo "== scan variant $v"
  VLFB_LIB=$L timeout 200 python bench_fbo.py --modes infer_fold_bf16 --R 64,256 --L 1200,3600 --layers 2 --steps 10 --out $O/r2o_fbo_$v.txt > /dev/null 2>&1
  cat $O/r2o_fbo_$v.txt
done
