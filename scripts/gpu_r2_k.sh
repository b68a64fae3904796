#!/bin/bash
# call K: SpatialBN GPU tests, bf16 bank scan (48 KB tiles, re-unpacked rows), A/B of the last two epilogue changes
# (sign-bit masks, look-ahead depth) against the 12.49 ms of call G
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
timeout 400 python -m pytest tests/test_spatial_bn.py -m gpu -q -s > $O/r2k_bn_tests.log 2>&1; echo "bn tests rc=$?"; grep -E "rel err|cosine|passed|failed" $O/r2k_bn_tests.log | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "bank_scan or maxpool" -q > $O/r2k_scan_tests.log 2>&1; echo "scan tests rc=$?"; tail -n 4 $O/r2k_scan_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0"
timeout 300 $B --dump-gemms $O/r2k_gemm_table.txt > $O/r2k_bench.log 2>&1
VLFB_RELU_BITS=0 timeout 300 $B > $O/r2k_bench_nobits.log 2>&1
VLFB_LIB=$CS/libvlfb_d1.so timeout 300 $B --dump-gemms $O/r2k_gemm_table_d1.txt > $O/r2k_bench_d1.log 2>&1
VLFB_FUSE_GRAD_FINISH=0 timeout 300 $B > $O/r2k_bench_nofuse.log 2>&1
for f in bench bench_nobits bench_d1 bench_nofuse; do echo "== $f"; tail -1 $O/r2k_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], {k:(v['ms'], v['tensor_frac']) for k,v in d['roofline']['by_stage'].items()}, {k:v['ms'] for k,v in d['roofline']['by_kind'].items()}, d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
timeout 300 python bench_fbo.py --modes infer_fold,infer_fold_bf16 --R 64,256 --L 1200,3600 --layers 2 --steps 10 --out $O/r2k_fbo.txt > $O/r2k_fbo.jsonl 2>&1
cat $O/r2k_fbo.txt
