#!/bin/bash
# call E: ncu launch list of one eager training step (device time + DRAM bytes per kernel), bench with the templated epilogue,
# conv / epilogue timelines with the current trace build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2e_kernels.log 2>&1; echo "kernels rc=$?"
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --dump-gemms $O/r2e_gemm_table.txt > $O/r2e_bench.log 2>&1
VLFB_FUSE_GRAD_FINISH=1 timeout 300 $B --dump-gemms $O/r2e_gemm_table_fuse.txt > $O/r2e_bench_fuse.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file /tmp/r2e_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu-baseline > $O/r2e_ncu_bench.log 2>&1
echo "ncu rc=$?"
python scripts/summarize_ncu_launches.py /tmp/r2e_ncu_launches.csv $O/r2e_ncu_launches_summary.txt $O/r2e_ncu_traffic.json > /dev/null 2>&1
gzip -c /tmp/r2e_ncu_launches.csv > $O/r2e_ncu_launches.csv.gz
export VLFB_LIB=$PWD/video-long-term-feature-banks_b200/csrc/libvlfb_trace.so
rm -f $O/r2e_trace.txt
for l in "res5_2b fwd" "res4_2b fwd" "res5_2a1 fwd"; do timeout 120 python scripts/trace_gemm.py $l -1 -1 >> $O/r2e_trace.txt 2>&1; done
unset VLFB_LIB
for f in bench bench_fuse; do echo "== $f"; tail -1 $O/r2e_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
tail -n 3 $O/r2e_kernels.log
cat $O/r2e_ncu_launches_summary.txt
