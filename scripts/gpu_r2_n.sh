#!/bin/bash
# call N (gpurun --gpus 2): NCCL data-parallel step == 1-GPU step on the concatenated batch; bench at N=2 (overlapped
# all-reduce inside the captured step) next to N=1 on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > $O/r2n_dist_test.log 2>&1; echo "dist test rc=$?"
grep -E "losses|update|passed|failed|skipped" $O/r2n_dist_test.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0 > $O/r2n_bench_n1.log 2> $O/r2n_bench_n1.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2n_bench_n2.log 2> $O/r2n_bench_n2.err; echo "bench n2 rc=$?"
VLFB_OVERLAP=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2n_bench_n2_nooverlap.log 2> $O/r2n_bench_n2_nooverlap.err; echo "bench n2 no-overlap rc=$?"
for f in bench_n1 bench_n2 bench_n2_nooverlap; do echo "== $f"; tail -1 $O/r2n_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['clocks'])
except Exception as e: print('ERR', e)
"; done
grep -i "falling back\|error" $O/r2n_bench_n2.err | head -5
