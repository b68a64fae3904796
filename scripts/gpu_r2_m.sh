#!/bin/bash
# call M: SpatialBN model test, bf16 scan tile / occupancy variants, bench with the smem job lookup of the weight transposes,
# ncu --set full of the res4 / res5 launches (2 and 8 clips per GPU) and of the bank scan
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CS=$PWD/video-long-term-feature-banks_b200/csrc
timeout 300 python -m pytest tests/test_spatial_bn.py -m gpu -q -s > $O/r2m_bn_tests.log 2>&1; echo "bn tests rc=$?"; grep -E "rel err vs|cosine vs|passed|failed" $O/r2m_bn_tests.log | grep -v print | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fbo --large-batch 0 --dump-gemms $O/r2m_gemm_table.txt > $O/r2m_bench.log 2>&1
tail -1 $O/r2m_bench.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['e2e']['value'])
except Exception as e: print('ERR', e)
"
for v in default scan_r4b4 scan_r6b3; do
  L=$CS/libvlfb_$v.so; [ $v = default ] && L=$CS/libvlfb.so
  echo "== scan variant $v"
  VLFB_LIB=$L timeout 200 python bench_fbo.py --modes infer_fold_bf16 --R 64,256 --L 1200,3600 --layers 2 --steps 10 --out $O/r2m_fbo_$v.txt > /dev/null 2>&1
  cat $O/r2m_fbo_$v.txt
done
for c in 2 8; do
  VLFB_PROF_CLIPS=$c timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -f -o /tmp/rep_res45_c$c \
    python scripts/prof_gemm3.py > $O/r2m_ncu_res45_c$c.log 2>&1
  echo "ncu res45 clips=$c rc=$?"
  ncu -i /tmp/rep_res45_c$c.ncu-rep --page raw --csv > /tmp/res45_c${c}_raw.csv 2>/dev/null
  python scripts/summarize_ncu_full.py /tmp/res45_c${c}_raw.csv $O/r2m_ncu_full_res45_clips${c}_summary.txt > /dev/null
  gzip -c /tmp/res45_c${c}_raw.csv > $O/r2m_ncu_full_res45_clips${c}_raw.csv.gz
  cat $O/r2m_ncu_full_res45_clips${c}_summary.txt | cut -c1-220
done
timeout 300 ncu --set full --clock-control none -k regex:fbo_bank_scan -c 4 -f -o /tmp/rep_scan \
  python bench_fbo.py --modes infer_fold,infer_fold_bf16 --R 256 --L 3600 --layers 2 --steps 1 --warmup 1 > $O/r2m_ncu_scan.log 2>&1
echo "ncu scan rc=$?"
ncu -i /tmp/rep_scan.ncu-rep --page raw --csv > /tmp/scan_raw.csv 2>/dev/null
python - <<'PY' > $O/r2m_ncu_full_scan_summary.txt
import csv
rows = list(csv.reader(open('/tmp/scan_raw.csv', errors='replace')))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, units = rows[h], rows[h + 1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size']
idx = [hdr.index(w) for w in want if w in hdr]
print(' | '.join('%s [%s]' % (hdr[i], units[i]) for i in idx))
for r in rows[h + 2:]:
    if len(r) >= len(hdr):
        print(' | '.join(r[i][:60] for i in idx))
PY
cat $O/r2m_ncu_full_scan_summary.txt | cut -c1-250
