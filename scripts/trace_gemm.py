"""Per-CTA timeline of one gemm_tc launch (diagnostic build libvlfb_trace.so, -DVLFB_TRACE).
usage: VLFB_LIB=.../libvlfb_trace.so trace_gemm.py <res5_2b|res4_2b|res5_2a1> <fwd|dgrad> <pair> <stream_k> [tile_n]
Prints, in microseconds relative to the earliest CTA entry (globaltimer-aligned), the median / min / max over CTAs of
every recorded event, plus the slowest CTA's own timeline."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib'))
from vlfb import kernels as K  # noqa: E402
from vlfb import libvlfb as L  # noqa: E402

NAMES = {0: 'entry', 2: 'prologue done', 3: 'exit (before final sync)'}
for i in range(4):
    NAMES[8 + 2 * i] = 'prod item%d first copy' % i
    NAMES[9 + 2 * i] = 'prod item%d last copy issued' % i
    NAMES[16 + 2 * i] = 'mma  item%d first chunk landed' % i
    NAMES[17 + 2 * i] = 'mma  item%d last mma issued' % i
    NAMES[24 + 4 * i] = 'epi  item%d accumulator ready' % i
    NAMES[25 + 4 * i] = 'epi  item%d blocks done' % i
    NAMES[26 + 4 * i] = 'epi  item%d piece counted' % i
    NAMES[27 + 4 * i] = 'epi  item%d fix-up done' % i


def main():
    layer, op = sys.argv[1], sys.argv[2]
    K.GEMM_OPTS.update(dict(pair=int(sys.argv[3]), stream_k=int(sys.argv[4]), tile_n=int(sys.argv[5]) if len(sys.argv) > 5 else 0))
    ci, co, ker, pd, dil = {'res5_2b': (512, 512, (1, 3, 3), (0, 2, 2), (1, 2, 2)),
                            'res4_2b': (256, 256, (1, 3, 3), (0, 1, 1), (1, 1, 1)),
                            'res5_2a1': (2048, 512, (1, 1, 1), (0, 0, 0), (1, 1, 1))}[layer]
    shp = (2, 16, 14, 14)
    g = K.conv_geom(shp + (ci,), co, ker, (1, 1, 1), pd, dil)
    x = torch.randn(shp + (ci,), device='cuda')
    w = torch.randn((co,) + ker + (ci,), device='cuda') * 0.05
    taps = ker[0] * ker[1] * ker[2]
    wt = torch.randn((ci, taps, co), device='cuda') * 0.05
    s, b = torch.rand(co, device='cuda') + 0.5, torch.randn(co, device='cuda')
    y = torch.empty(K.out_shape(g), device='cuda')
    dx = torch.empty(shp + (ci,), device='cuda')
    lib = L.load()
    lib.vlfb_debug_set_trace.argtypes = [C.c_void_p]
    buf = torch.zeros(512 * 64, dtype=torch.int64, device='cuda')

    def run():
        if op == 'fwd':
            K.conv_fwd(x, w, y, g, scale=s, bias=b, relu=True, tf32_out=True)
        else:
            K.conv_dgrad(y, wt, dx, g)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    lib.vlfb_debug_set_trace(C.c_void_p(buf.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.vlfb_debug_set_trace(None)
    t = buf.cpu().numpy().astype(np.int64).reshape(512, 64)
    ctas = np.nonzero(t[:, 0])[0]
    t = t[ctas]
    clk_mhz = float(os.environ.get('SM_MHZ', '1965'))
    # align CTAs: globaltimer (ns) at entry gives each CTA's offset; clock64 deltas inside a CTA
    g0 = t[:, 1].min()
    base_us = (t[:, 1] - g0) / 1e3
    rel = np.where(t > 0, (t - t[:, [0]]) / clk_mhz + base_us[:, None], np.nan)
    print('# %s %s opts=%s: %d CTAs, entry spread %.1f us' % (layer, op, dict(K.GEMM_OPTS), len(ctas), base_us.max()))
    print('%-34s %9s %9s %9s %5s' % ('event', 'median', 'min', 'max', 'n'))
    for slot in sorted(NAMES):
        col = rel[:, slot]
        ok = ~np.isnan(col)
        if slot in (1,) or not ok.any():
            continue
        print('%-34s %9.2f %9.2f %9.2f %5d' % (NAMES[slot], np.median(col[ok]), col[ok].min(), col[ok].max(), ok.sum()))
    last = int(np.nanargmax(rel[:, 3]))
    print('# slowest CTA (block %d):' % ctas[last], ' '.join('%s=%.1f' % (NAMES[sl].replace(' ', '_'), rel[last, sl])
                                                           for sl in sorted(NAMES) if sl != 1 and not np.isnan(rel[last, sl])))


if __name__ == '__main__':
    main()
