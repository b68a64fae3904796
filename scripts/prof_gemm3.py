"""Third profiling set (for `ncu --set full -k regex:gemm_tc`): the res4 / res5 convolutions the north star quotes
("tensor-pipe % on the res4/res5 3D-conv stages"), config-2 shapes (2 clips: M = 2*16*14*14 = 6272 output positions).
Two launches per case (the second one has warm instruction / L2 state).  Order:
 [0,1] res5 branch2b fwd  1x3x3 dil 2  512->512   (M=6272 N=512  K=4608, +affine+ReLU)
 [2,3] res5 branch2b dgrad                        (M=6272 N=512  K=4608)
 [4,5] res5 branch2b wgrad                        (M=512  N=4608 K=6272)
 [6,7] res5 branch2a fwd  3x1x1       2048->512   (M=6272 N=512  K=6144)
 [8,9] res5 branch2c fwd  1x1x1       512->2048   (M=6272 N=2048 K=512, +affine+residual+ReLU)
 [10,11] res4 branch2a fwd 3x1x1      1024->256   (M=6272 N=256  K=3072)
 [12,13] res4 branch2b fwd 1x3x3      256->256    (M=6272 N=256  K=2304)
 [14,15] res4 branch2c fwd 1x1x1      256->1024   (M=6272 N=1024 K=256, +residual+ReLU)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib'))
from vlfb import kernels as K  # noqa: E402


CLIPS = int(os.environ.get('VLFB_PROF_CLIPS', '2'))      # 8: the large-batch shapes (M = 25088)


def case(Ci, Co, ker, pd, dil=(1, 1, 1), residual=False):
    g = K.conv_geom((CLIPS, 16, 14, 14, Ci), Co, ker, (1, 1, 1), pd, dil)
    x = torch.randn((CLIPS, 16, 14, 14, Ci), device='cuda')
    w = torch.randn((Co,) + tuple(ker) + (Ci,), device='cuda') * 0.05
    y = torch.empty(K.out_shape(g), device='cuda')
    s = torch.rand(Co, device='cuda') + 0.5
    b = torch.randn(Co, device='cuda')
    res = torch.randn(K.out_shape(g), device='cuda') if residual else None
    return g, x, w, y, s, b, res


def fwd(c, reps=2):
    g, x, w, y, s, b, r = c
    for _ in range(reps):
        K.conv_fwd(x, w, y, g, scale=s, bias=b, residual=r, relu=True, tf32_out=True)


def main():
    c = case(512, 512, (1, 3, 3), (0, 2, 2), (1, 2, 2))
    fwd(c)
    g, x, w, y, s, b, _ = c
    taps = 9
    wt = torch.empty((512, taps, 512), device='cuda')
    K.weight_transpose(w, wt, s)
    dx = torch.empty_like(x)
    for _ in range(2):
        K.conv_dgrad(y, wt, dx, g, tf32_out=True)
    dw = torch.zeros_like(w)
    for _ in range(2):
        K.conv_wgrad(y, x, dw, g, row_scale=s)
    fwd(case(2048, 512, (3, 1, 1), (1, 0, 0)))
    fwd(case(512, 2048, (1, 1, 1), (0, 0, 0), residual=True))
    fwd(case(1024, 256, (3, 1, 1), (1, 0, 0)))
    fwd(case(256, 256, (1, 3, 3), (0, 1, 1)))
    fwd(case(256, 1024, (1, 1, 1), (0, 0, 0), residual=True))
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
