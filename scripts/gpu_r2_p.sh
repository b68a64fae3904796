#!/bin/bash
# call P: three TMA issuer warps in the TMA-only GEMM builds (libvlfb_mi3.so) against the single issuer (libvlfb.so):
# kernel parity, per-launch table, step time; bench e2e with the long-lived clock sampler
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
VLFB_LIB=$CS/libvlfb_mi3.so timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2p_kernels_mi3.log 2>&1; echo "kernel tests (mi3) rc=$?"; tail -n 3 $O/r2p_kernels_mi3.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0"
timeout 300 $B --dump-gemms $O/r2p_gemm_table.txt > $O/r2p_bench.log 2>&1
VLFB_LIB=$CS/libvlfb_mi3.so timeout 300 $B --dump-gemms $O/r2p_gemm_table_mi3.txt > $O/r2p_bench_mi3.log 2>&1
VLFB_LIB=$CS/libvlfb_mi3.so timeout 300 $B --clips-per-gpu 8 > $O/r2p_bench_mi3_c8.log 2>&1
for f in bench bench_mi3 bench_mi3_c8; do echo "== $f"; tail -1 $O/r2p_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, {k:v['ms'] for k,v in d['roofline']['by_kind'].items()}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step_blocking_fetch'], d['clocks'])
except Exception as e: print('ERR', e)
"; done
VLFB_LIB=$CS/libvlfb_mi3.so timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "tiny or config2" > $O/r2p_model_mi3.log 2>&1; echo "model tests (mi3) rc=$?"; tail -n 3 $O/r2p_model_mi3.log
