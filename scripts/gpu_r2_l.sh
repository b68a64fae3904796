#!/bin/bash
# call L: all GPU tests after the gather-form max-pool backward (pool1: ReLU mask + rounding folded in), split-K kept for
# the non-local gradient products, SpatialBN bounds; bench with the large-batch sub-measurement
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/r2l_gpu_tests.log 2>&1; echo "tests rc=$?"
grep -E "BN tiny|gradient cosine" $O/r2l_gpu_tests.log | cut -c1-500; tail -n 4 $O/r2l_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --dump-gemms $O/r2l_gemm_table.txt > $O/r2l_bench.log 2> $O/r2l_bench.err; echo "bench rc=$?"
tail -1 $O/r2l_bench.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, {k:v['ms'] for k,v in d['roofline']['by_kind'].items()}, d['e2e']['value'], d['cpu_baseline'], d['large_batch'])
except Exception as e: print('ERR', e)
"
