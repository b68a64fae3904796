#!/bin/bash
# call Q: three TMA issuer warps (default build, lazy stage wait of the helpers) against one (libvlfb_mi1.so) and two
# (libvlfb_mi2.so): kernel + model parity on the default build, per-launch tables, step time at 2 and 8 clips
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2q_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 3 $O/r2q_kernels.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0"
timeout 300 $B --dump-gemms $O/r2q_gemm_table.txt > $O/r2q_bench.log 2>&1
VLFB_LIB=$CS/libvlfb_mi1.so timeout 300 $B --dump-gemms $O/r2q_gemm_table_mi1.txt > $O/r2q_bench_mi1.log 2>&1
VLFB_LIB=$CS/libvlfb_mi2.so timeout 300 $B > $O/r2q_bench_mi2.log 2>&1
timeout 300 $B --clips-per-gpu 8 > $O/r2q_bench_c8.log 2>&1
for f in bench bench_mi1 bench_mi2 bench_c8; do echo "== $f"; tail -1 $O/r2q_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, {k:v['ms'] for k,v in d['roofline']['by_kind'].items()}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step_blocking_fetch'], d['clocks'])
except Exception as e: print('ERR', e)
"; done
grep -c "mbarrier timeout" $O/r2q_bench*.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q > $O/r2q_model.log 2>&1; echo "model tests rc=$?"; tail -n 3 $O/r2q_model.log
