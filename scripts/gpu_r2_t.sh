#!/bin/bash
# call T: max-pool backward forms in the captured step: 1 = gather everywhere (default), 0 = scatter (fill + atomics + separate
# ReLU-backward pass), 2 = gather only for non-overlapping windows
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0 --no-roofline"
for m in 1 0 2; do
  VLFB_POOL_GATHER=$m timeout 120 $B > $O/r2t_bench_pool$m.log 2>&1
  echo "== VLFB_POOL_GATHER=$m"; tail -1 $O/r2t_bench_pool$m.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], 'e2e', d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
