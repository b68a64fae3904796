"""Second profiling set (for `ncu --set full -k regex:gemm_tc`): the launches that top the per-step table.
Order (2 launches each): [0,1] res2 branch2c fwd (pointwise K=64, +residual+ReLU, M=200704 N=256),
[2,3] conv1 fwd (stem, M=802816 N=64 K=1120), [4,5] conv1 wgrad (M=64 N=224 K=802816, 5 taps),
[6,7] res2 branch2b wgrad (1x3x3 gather, M=64 N=576 K=200704)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib'))
from vlfb import kernels as K  # noqa: E402


def conv_case(N, T, H, W, Ci, Co, ker, st, pd, dil=(1, 1, 1), residual=False, wshape=None):
    g = K.conv_geom((N, T, H, W, Ci), Co, ker, st, pd, dil)
    x = torch.randn((N, T, H, W, Ci), device='cuda')
    w = torch.randn(wshape or ((Co,) + tuple(ker) + (Ci,)), device='cuda') * 0.05
    y = torch.empty(K.out_shape(g), device='cuda')
    s = torch.rand(Co, device='cuda') + 0.5
    b = torch.randn(Co, device='cuda')
    res = torch.randn(K.out_shape(g), device='cuda') if residual else None
    return g, x, w, y, s, b, res


def main():
    which = sys.argv[1:] or ['k64', 'stemf', 'stemw', 'w3x3']
    reps = 2
    if 'k64' in which:
        g, x, w, y, s, b, r = conv_case(2, 32, 56, 56, 64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), residual=True)
        for _ in range(reps):
            K.conv_fwd(x, w, y, g, scale=s, bias=b, residual=r, relu=True, tf32_out=True)
    if 'stemf' in which or 'stemw' in which:
        g, x, w, y, s, b, _ = conv_case(2, 32, 224, 224, 4, 64, (5, 7, 7), (1, 2, 2), (2, 3, 3), wshape=(64, 5, 7, 8, 4))
        if 'stemf' in which:
            for _ in range(reps):
                K.conv_fwd(x, w, y, g, scale=s, bias=b, relu=True)
        if 'stemw' in which:
            dw = torch.zeros_like(w)
            for _ in range(reps):
                K.conv_wgrad(y, x, dw, g, row_scale=s)
    if 'w3x3' in which:
        g, x, w, y, s, b, _ = conv_case(2, 32, 56, 56, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        dw = torch.zeros_like(w)
        for _ in range(reps):
            K.conv_wgrad(y, x, dw, g, row_scale=s)
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
