#!/bin/bash
# call N2 (gpurun --gpus 2): does `bench.py --gpus 2` (overlapped all-reduce captured in the step) exit after printing?
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2n2_bench_n2.log 2> $O/r2n2_bench_n2.err; echo "bench n2 rc=$? after $(( $(date +%s) - t0 )) s"
tail -1 $O/r2n2_bench_n2.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['clocks'])
except Exception as e: print('ERR', e)
"
grep -i "shutdown\|error\|Traceback" $O/r2n2_bench_n2.err | head -5
