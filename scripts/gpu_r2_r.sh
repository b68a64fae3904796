#!/bin/bash
# call R: MN-major operand tiles as ONE 4-D TMA box (dense MN and wgrad activations of convolutions without spatial taps)
# against one box per 32-element atom (VLFB_FUSE_ATOMS=0); sampler perturbation of the e2e loop
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r2r_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 3 $O/r2r_kernels.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0"
VLFB_DEBUG=1 timeout 300 $B --dump-gemms $O/r2r_gemm_table.txt > $O/r2r_bench.log 2> $O/r2r_bench.err
echo "launch plans by tma modes:"; grep -o "tma=[0-9],[0-9]" $O/r2r_bench.err | sort | uniq -c
VLFB_FUSE_ATOMS=0 timeout 300 $B --dump-gemms $O/r2r_gemm_table_nofuse.txt > $O/r2r_bench_nofuse.log 2>&1
timeout 300 $B --clips-per-gpu 8 > $O/r2r_bench_c8.log 2>&1
for f in bench bench_nofuse bench_c8; do echo "== $f"; tail -1 $O/r2r_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, {k:v['ms'] for k,v in d['roofline']['by_kind'].items()}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step_blocking_fetch'])
except Exception as e: print('ERR', e)
"; done
grep -c "mbarrier timeout" $O/r2r_bench*.log
timeout 300 python scripts/diag_e2e.py > $O/r2r_diag_e2e.txt 2>&1; grep -v Warning $O/r2r_diag_e2e.txt | tail -12
timeout 900 python -m pytest tests/test_gpu_model.py -x -q > $O/r2r_model.log 2>&1; echo "model tests rc=$?"; tail -n 3 $O/r2r_model.log
