#!/bin/bash
# final validation of round 2: whole GPU suite, smoke, bench (both arms), ncu launch list of one eager step ->
# profiles/ncu_traffic.json
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/final_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 $O/final_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/final_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --dump-gemms $O/final_gemm_table.txt > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err; echo "bench ref rc=$?"
tail -1 $O/final_bench.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic_source'][:60], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, 'e2e', d['e2e'], d['cpu_baseline'], d['large_batch'], d['clocks']); print([(f['mode'], f['R'], f['L'], f['ms'], f['hbm_frac'], (f['scan'] or {}).get('frac')) for f in d['fbo_microbench']])
except Exception as e: print('ERR', e)
"
tail -1 $O/final_bench_ref.json | cut -c1-400
timeout 700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -c 1250 \
  --log-file /tmp/final_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu-baseline --no-fbo --large-batch 0 > $O/final_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
python scripts/summarize_ncu_launches.py /tmp/final_ncu_launches.csv $O/final_ncu_launches_summary.txt $O/final_ncu_traffic.json > /dev/null 2>&1
gzip -c /tmp/final_ncu_launches.csv > $O/final_ncu_launches.csv.gz
head -24 $O/final_ncu_launches_summary.txt; tail -1 $O/final_ncu_launches_summary.txt
