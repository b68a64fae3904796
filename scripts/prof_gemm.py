"""Runs a handful of representative hot-path GEMM launches (for `ncu --set full -k regex:gemm_tc`).
Order of launches: [0,1] res5 2b conv fwd (3x3 dil 2, M=6272 N=512 K=4608), [2,3] res2 2c fwd (+residual+relu,
M=200704 N=256 K=64), [4,5] res4 2b wgrad (M=256 N=2304 K=6272), [6,7] res3 2a dgrad pointwise, [8,9] res2 2b 3x3 fwd (M=200704 N=64 K=576)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib'))
from vlfb import kernels as K  # noqa: E402


def conv_case(N, T, H, W, Ci, Co, ker, st, pd, dil, residual=False):
    g = K.conv_geom((N, T, H, W, Ci), Co, ker, st, pd, dil)
    x = torch.randn((N, T, H, W, Ci), device='cuda')
    w = torch.randn((Co,) + tuple(ker) + (Ci,), device='cuda') * 0.05
    y = torch.empty(K.out_shape(g), device='cuda')
    s = torch.rand(Co, device='cuda') + 0.5
    b = torch.randn(Co, device='cuda')
    res = torch.randn(K.out_shape(g), device='cuda') if residual else None
    return g, x, w, y, s, b, res


def main():
    reps = 2
    g, x, w, y, s, b, _ = conv_case(2, 16, 14, 14, 512, 512, (1, 3, 3), (1, 1, 1), (0, 2, 2), (1, 2, 2))
    for _ in range(reps):
        K.conv_fwd(x, w, y, g, scale=s, bias=b, relu=True, tf32_out=True)
    g2, x2, w2, y2, s2, b2, r2 = conv_case(2, 32, 56, 56, 64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1), True)
    for _ in range(reps):
        K.conv_fwd(x2, w2, y2, g2, scale=s2, bias=b2, residual=r2, relu=True, tf32_out=True)
    g3, x3, w3, y3, s3, b3, _ = conv_case(2, 16, 14, 14, 256, 256, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1))
    dw = torch.zeros_like(w3)
    for _ in range(reps):
        K.conv_wgrad(y3, x3, dw, g3, row_scale=s3)
    g4, x4, w4, y4, s4, b4, _ = conv_case(2, 16, 28, 28, 512, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1))
    wt = torch.empty((512, 1, 128), device='cuda')
    K.weight_transpose(w4, wt, s4)
    dx = torch.empty_like(x4)
    for _ in range(reps):
        K.conv_dgrad(y4, wt, dx, g4)
    g5, x5, w5, y5, s5, b5, _ = conv_case(2, 32, 56, 56, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1))
    for _ in range(reps):
        K.conv_fwd(x5, w5, y5, g5, scale=s5, bias=b5, relu=True, tf32_out=True)
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
