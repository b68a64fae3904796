"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`python bench.py --no-graph ...`: one training step = the launches between two consecutive `sgd_k` kernels (the fused
optimizer runs once per step); the last complete step of the capture is reported, grouped by kernel."""
import csv
import sys


def main(path, out=None, traffic_json=None):
    rows = list(csv.reader(open(path, errors='replace')))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    kn, mn, mv, idc, mu = (hdr.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Metric Unit'))
    launches = {}
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        d = launches.setdefault(int(r[idc]), {'name': r[kn]})
        v = float(r[mv].replace(',', ''))
        unit = r[mu]
        if r[mn] == 'gpu__time_duration.sum':
            d['us'] = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
        elif r[mn].startswith('dram__bytes'):
            scale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1e-6)
            d['r' if 'read' in r[mn] else 'w'] = v * scale
    ids = sorted(launches)
    sgd = [i for i in ids if 'sgd_k' in launches[i]['name']]
    assert len(sgd) >= 2, 'need two optimizer launches to delimit a step'
    step = [launches[i] for i in ids if sgd[-2] < i <= sgd[-1]]
    groups = {}
    for d in step:
        name = d['name'].split('(')[0].replace('void ', '').replace('vlfb::', '').replace('<unnamed>::', '')[:60]
        g = groups.setdefault(name, [0.0, 0, 0.0, 0.0])
        g[0] += d.get('us', 0.0)
        g[1] += 1
        g[2] += d.get('r', 0.0)
        g[3] += d.get('w', 0.0)
    total = sum(g[0] for g in groups.values())
    lines = ['%s: last complete training step of the capture (between two sgd_k launches); serialised, cold-cache per-launch times' % path,
             'kernels in step %d, total %.1f us' % (len(step), total)]
    for name, g in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        lines.append('%9.1f us %5.1f%% x%-4d dram R %8.1f MB W %8.1f MB  %s' % (g[0], 100 * g[0] / total, g[1], g[2], g[3], name))
    gem = [g for n, g in groups.items() if 'gemm_tc' in n]
    gt, gn, gr, gw = (sum(g[k] for g in gem) for k in range(4))
    lines.append('GEMM: %d launches, %.1f us (%.1f%% of the step), DRAM read %.1f MB + write %.1f MB = %.2f MB per launch' % (
        gn, gt, 100 * gt / total, gr, gw, (gr + gw) / max(gn, 1)))
    text = '\n'.join(lines)
    print(text)
    if out:
        open(out, 'w').write(text + '\n')
    if traffic_json:
        import hashlib
        import json
        import os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        src = os.path.join(root, 'video-long-term-feature-banks_b200', 'csrc', 'gemm_tc.cu')
        rec = {'gemm_tc_sha256': hashlib.sha256(open(src, 'rb').read()).hexdigest(),
               'dram_bytes_per_gemm_launch': (gr + gw) * 1e6 / max(gn, 1), 'gemm_launches_per_step': gn,
               'gemm_dram_read_mb': gr, 'gemm_dram_write_mb': gw, 'gemm_us': gt, 'step_kernels': len(step),
               'step_us_serialised': total,
               'step_dram_mb': sum(g[2] + g[3] for g in groups.values()),
               'source': '%s: dram__bytes_read.sum + dram__bytes_write.sum over the %d gemm_tc_kernel launches of one '
                         'training step (ncu launch list of `bench.py --steps 1 --no-graph`)' % (out or path, gn)}
        json.dump(rec, open(traffic_json, 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
