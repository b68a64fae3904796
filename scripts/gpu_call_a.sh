#!/bin/bash
# GPU call A of the second session: new kernels / lowerings first, then the FBO sweep, the step bench, ncu of the scan.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "fbo_bank_scan or lfb_gather or device_bank or inference or r101" > gpurun_out/a_tests_new.log 2>&1
echo "tests_new rc=$?" >> gpurun_out/a_status.txt
timeout 420 python bench_fbo.py --R 4,64,256 --L 300,3600 --layers 2 --modes infer_fold,infer,train --steps 10 --out gpurun_out/a_fbo_sweep.txt > gpurun_out/a_fbo_sweep.jsonl 2> gpurun_out/a_fbo_sweep.err
echo "fbo_sweep rc=$?" >> gpurun_out/a_status.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-gemms gpurun_out/a_gemm_table.txt > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_status.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:fbo_bank_scan -c 2 -f -o gpurun_out/a_fbo_scan python bench_fbo.py --R 64 --L 300 --layers 2 --modes infer_fold --steps 1 --warmup 1 > gpurun_out/a_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/a_status.txt
tail -3 gpurun_out/a_tests_new.log; cat gpurun_out/a_status.txt; cat gpurun_out/a_fbo_sweep.txt; tail -c 1500 gpurun_out/a_bench.json
