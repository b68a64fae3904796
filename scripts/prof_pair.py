"""One production conv under a chosen tiling variant, for ncu (scripts/gpu_r2_b.sh).
usage: prof_pair.py <layer: res5_2b|res4_2b> <op: fwd|dgrad> <pair> <stream_k> [tile_n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib'))
from vlfb import kernels as K  # noqa: E402

layer, op = sys.argv[1], sys.argv[2]
K.GEMM_OPTS.update(dict(pair=int(sys.argv[3]), stream_k=int(sys.argv[4]), tile_n=int(sys.argv[5]) if len(sys.argv) > 5 else 0))
ci, co, ker, pd, dil = {'res5_2b': (512, 512, (1, 3, 3), (0, 2, 2), (1, 2, 2)),
                        'res4_2b': (256, 256, (1, 3, 3), (0, 1, 1), (1, 1, 1))}[layer]
shp = (2, 16, 14, 14)
g = K.conv_geom(shp + (ci,), co, ker, (1, 1, 1), pd, dil)
x = torch.randn(shp + (ci,), device='cuda')
w = torch.randn((co,) + ker + (ci,), device='cuda') * 0.05
wt = torch.randn((ci, 9, co), device='cuda') * 0.05
s, b = torch.rand(co, device='cuda') + 0.5, torch.randn(co, device='cuda')
y = torch.empty(K.out_shape(g), device='cuda')
dx = torch.empty(shp + (ci,), device='cuda')
for _ in range(4):
    if op == 'fwd':
        K.conv_fwd(x, w, y, g, scale=s, bias=b, relu=True, tf32_out=True)
    else:
        K.conv_dgrad(y, wt, dx, g)
torch.cuda.synchronize()
