"""Summarise an `ncu --set full --page raw --csv` export of the gemm_tc launches of scripts/prof_gemm3.py: one line per
launch with the metrics the north star quotes (tensor-pipe activity, DRAM traffic, duration)."""
import csv
import sys

LABELS = ['res5_2b fwd', 'res5_2b fwd', 'res5_2b dgrad', 'res5_2b dgrad', 'res5_2b wgrad', 'res5_2b wgrad', 'res5_2a fwd', 'res5_2a fwd',
          'res5_2c fwd+res', 'res5_2c fwd+res', 'res4_2a fwd', 'res4_2a fwd', 'res4_2b fwd', 'res4_2b fwd', 'res4_2c fwd+res',
          'res4_2c fwd+res']
WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']


def main(path, out=None):
    rows = list(csv.reader(open(path, errors='replace')))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units = rows[h], rows[h + 1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    lines = ['# %s' % path, '# ' + ' | '.join('%s [%s]' % (w, units[i]) for w, i in cols)]
    k = 0
    for r in rows[h + 2:]:
        if len(r) < len(hdr) or 'gemm_tc' not in r[hdr.index('Kernel Name')]:
            continue
        lab = LABELS[k] if k < len(LABELS) else 'launch %d' % k
        lines.append('%-16s %s' % (lab + (' (warm)' if k % 2 else ''), ' | '.join(r[i] for _, i in cols)))
        k += 1
    text = '\n'.join(lines)
    print(text)
    if out:
        open(out, 'w').write(text + '\n')


if __name__ == '__main__':
    main(*sys.argv[1:3])
