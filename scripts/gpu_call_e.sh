#!/bin/bash
# A/B of the optional tile widths (VLFB_BN_EXTRA=1): whole GPU suite with them on, then the bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
VLFB_BN_EXTRA=1 timeout 170 python -m pytest tests -m gpu -q -x > gpurun_out/e_gpu_tests_extra.log 2>&1
echo "gpu_tests_extra rc=$?" > gpurun_out/e_status.txt
VLFB_BN_EXTRA=1 timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms gpurun_out/e_gemm_table_extra.txt > gpurun_out/e_bench_extra.json 2> gpurun_out/e_bench_extra.err
echo "bench_extra rc=$?" >> gpurun_out/e_status.txt
cat gpurun_out/e_status.txt; tail -3 gpurun_out/e_gpu_tests_extra.log; python -c "
import json; d=json.load(open('gpurun_out/e_bench_extra.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['by_stage'])"
