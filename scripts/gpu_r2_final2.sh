#!/bin/bash
# final validation, part 2: ncu launch list of one eager training step (the first attempt stopped after 1250 kernels, which
# the parameter initialisation + the first step used up) -> profiles/ncu_traffic.json
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --launch-skip 700 -c 1500 \
  --log-file /tmp/final_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu-baseline --no-fbo --large-batch 0 > $O/final_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
python scripts/summarize_ncu_launches.py /tmp/final_ncu_launches.csv $O/final_ncu_launches_summary.txt $O/final_ncu_traffic.json > /dev/null 2>&1
gzip -c /tmp/final_ncu_launches.csv > $O/final_ncu_launches.csv.gz
grep -c sgd_k /tmp/final_ncu_launches.csv
head -30 $O/final_ncu_launches_summary.txt; tail -1 $O/final_ncu_launches_summary.txt
