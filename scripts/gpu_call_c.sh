#!/bin/bash
# Final validation of the round: whole GPU suite, smoke, bench (both arms), ncu launch list, ncu --set full of res4/res5 convs.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c_gpu_tests.log 2>&1
echo "gpu_tests rc=$?" > gpurun_out/c_status.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/c_status.txt
timeout 400 python bench.py --steps 20 --warmup 3 --dump-gemms gpurun_out/c_gemm_table.txt > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?" >> gpurun_out/c_status.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err
echo "bench_ref rc=$?" >> gpurun_out/c_status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -f -o gpurun_out/c_gemm3 python scripts/prof_gemm3.py > gpurun_out/c_ncu_gemm3.log 2>&1
echo "ncu_gemm3 rc=$?" >> gpurun_out/c_status.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/c_ncu_bench.log 2>&1
echo "ncu_launches rc=$?" >> gpurun_out/c_status.txt
tail -3 gpurun_out/c_gpu_tests.log; cat gpurun_out/c_status.txt; tail -2 gpurun_out/c_smoke.log; head -c 1200 gpurun_out/c_bench.json; echo; cat gpurun_out/c_bench_ref.json | head -c 600
