#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "lfb_gather or device_bank or inference or r101 or fbo_bank_scan" > gpurun_out/b_tests_new.log 2>&1
echo "tests_new rc=$?" > gpurun_out/b_status.txt
timeout 300 python bench_fbo.py --R 4,16,64,256 --L 60,300,3600 --layers 3 --modes infer_fold,infer --steps 10 --out gpurun_out/b_fbo_sweep_3l.txt > gpurun_out/b_fbo_sweep_3l.jsonl 2> gpurun_out/b_fbo_sweep_3l.err
echo "fbo_sweep_3l rc=$?" >> gpurun_out/b_status.txt
timeout 200 python bench_fbo.py --R 4,16,64,256 --L 60,300,1200,3600 --layers 2 --modes infer_fold --steps 10 --out gpurun_out/b_fbo_sweep_2l.txt > gpurun_out/b_fbo_sweep_2l.jsonl 2> gpurun_out/b_fbo_sweep_2l.err
echo "fbo_sweep_2l rc=$?" >> gpurun_out/b_status.txt
timeout 120 python bench_fbo.py --R 4,16 --L 300 --layers 2 --modes train --steps 10 --out gpurun_out/b_fbo_train.txt > gpurun_out/b_fbo_train.jsonl 2> gpurun_out/b_fbo_train.err
echo "fbo_train rc=$?" >> gpurun_out/b_status.txt
tail -4 gpurun_out/b_tests_new.log; cat gpurun_out/b_status.txt; cat gpurun_out/b_fbo_sweep_3l.txt gpurun_out/b_fbo_sweep_2l.txt gpurun_out/b_fbo_train.txt
