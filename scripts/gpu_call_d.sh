#!/bin/bash
# Last call of the round: bench with the one-pass clip dequeue, smoke, the tests the feed path touches, then the whole suite.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench rc=$?" > gpurun_out/d_status.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/d_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/d_status.txt
timeout 200 python -m pytest tests -m gpu -q -k "elementwise_and_layout or enqueue_blobs or tiny_fbo_nl_train_step or full_size_config2 or device_bank" > gpurun_out/d_tests_feed.log 2>&1
echo "tests_feed rc=$?" >> gpurun_out/d_status.txt
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/d_gpu_tests.log 2>&1
echo "gpu_tests rc=$?" >> gpurun_out/d_status.txt
cat gpurun_out/d_status.txt; tail -2 gpurun_out/d_tests_feed.log; tail -2 gpurun_out/d_gpu_tests.log; python -c "
import json; d=json.load(open('gpurun_out/d_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
