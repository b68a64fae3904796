#!/bin/bash
# call S: CTA pairs for the wgrads with spatial taps (8 im2col boxes per chunk -> 4 per SM), pairs everywhere, e2e after the
# wall-clock fix
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0"
timeout 300 $B > $O/r2s_bench.log 2>&1
VLFB_WGRAD_PAIR=1 timeout 300 $B --dump-gemms $O/r2s_gemm_table_wpair.txt > $O/r2s_bench_wpair.log 2>&1
VLFB_PAIR=1 timeout 300 $B > $O/r2s_bench_pair.log 2>&1
for f in bench bench_wpair bench_pair; do echo "== $f"; tail -1 $O/r2s_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, {k:v['ms'] for k,v in d['roofline']['by_kind'].items()}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step_blocking_fetch'])
except Exception as e: print('ERR', e)
"; done
grep "conv_wgrad" $O/r2s_gemm_table_wpair.txt | head -12
