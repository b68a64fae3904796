"""Caffe2 operator semantics restated on PyTorch-CPU / numpy (test infrastructure).

Caffe2 itself is not in /root/reference (un-vendored dependency, INSTALL.md:24-31);
each function restates the published operator definition and cites the reference
call site that relies on it.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def affine_nd(x, s, b):
    """AffineNd fwd: y[n,c,...] = x*s[c]+b[c]  (caffe2_customized_ops/video/affine_nd_op.cu:32-44)."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    return x * s.view(shape) + b.view(shape)


def affine_nd_grad(dy, s):
    """AffineNdGradient: dX = dY*s[c]; no ds/db (affine_nd_op.cu:47-58, affine_nd_op.cc:45-53)."""
    shape = [1, -1] + [1] * (dy.dim() - 2)
    return dy * s.view(shape)


def spatial_bn(x, s, b, running_mean, running_var, eps=1e-5, momentum=0.9, is_test=False):
    """Caffe2 SpatialBN, order NCHW, any number of spatial dims (emitted by model_builder_video.py:186-190,
    resnet_video.py:185-188, nonlocal_helper.py:146-150).  Returns (y, new_running_mean, new_running_var, saved_mean,
    saved_inv_std).  Training mode normalises with the batch mean and the BIASED batch variance; the running statistics
    follow `running = running * momentum + batch * (1 - momentum)` with the UNBIASED batch variance; `saved_inv_std` =
    1 / sqrt(biased var + eps) is the '_bn_siv' blob lib/utils/bn_helper.py:170-176 inverts.  Test mode normalises
    with the running statistics ('_bn_riv' is a variance, bn_helper.py:216-219; lib/utils/checkpoints.py:108-110 folds
    it as scale / sqrt(riv + eps))."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    if is_test:
        f = s / torch.sqrt(running_var + eps)
        return x * f.view(shape) + (b - running_mean * f).view(shape), running_mean, running_var, None, None
    dims = [0] + list(range(2, x.dim()))
    m = x.numel() // x.shape[1]
    mean = x.mean(dim=dims)
    var = ((x - mean.view(shape)) ** 2).mean(dim=dims)
    inv_std = 1.0 / torch.sqrt(var + eps)
    y = (x - mean.view(shape)) * (inv_std * s).view(shape) + b.view(shape)
    new_rm = running_mean * momentum + mean.detach() * (1.0 - momentum)
    new_rv = running_var * momentum + var.detach() * (float(m) / max(m - 1, 1)) * (1.0 - momentum)
    return y, new_rm, new_rv, mean.detach(), inv_std.detach()


def conv_nd(x, w, b=None, strides=(1, 1, 1), pads=(0, 0, 0), dilations=(1, 1, 1)):
    """Caffe2 Conv (NCTHW cross-correlation, symmetric pads) as emitted at
    lib/models/resnet_video.py:169-179 and model_builder_video.py:211-217."""
    return F.conv3d(x, w, b, stride=tuple(strides), padding=tuple(pads), dilation=tuple(dilations))


def max_pool_nd(x, kernels, strides, pads):
    """Caffe2 MaxPool: padding ignored (-inf), floor output size (resnet_video.py:190-196)."""
    if len(kernels) == 3:
        return F.max_pool3d(x, tuple(kernels), tuple(strides), tuple(pads))
    return F.max_pool2d(x, tuple(kernels), tuple(strides), tuple(pads))


def avg_pool_nd(x, kernels, strides, pads):
    """Caffe2 AveragePool (head_helper.py:37-40, :92-98); pads are 0 at every call site."""
    return F.avg_pool3d(x, tuple(kernels), tuple(strides), tuple(pads))


def batch_matmul(a, b, trans_a=0, trans_b=0):
    """Caffe2 BatchMatMul (nonlocal_helper.py:94-95,121; lfb_helper.py:223-224,234)."""
    if trans_a:
        a = a.transpose(1, 2)
    if trans_b:
        b = b.transpose(1, 2)
    return torch.bmm(a, b)


def softmax_axis2(x):
    """Softmax(axis=2) on a 3-D blob: rows = dims[0:2] flattened (nonlocal_helper.py:104-105)."""
    return torch.softmax(x, dim=2)


def layer_norm_axis1(x, eps=1e-5):
    """Caffe2 LayerNorm(axis=1, epsilon=1e-5): normalise over all dims >= 1, biased
    variance, no learnable gamma/beta (lfb_helper.py:160-167)."""
    n = x.shape[0]
    flat = x.reshape(n, -1)
    mean = flat.mean(dim=1, keepdim=True)
    var = ((flat - mean) ** 2).mean(dim=1, keepdim=True)
    std = torch.sqrt(var + eps)
    return ((flat - mean) / std).reshape(x.shape), mean, std


def fc(x, w, b):
    """Caffe2 FC: y = x.W^T + b, W (out,in); input flattened from axis 1 (resnet_video.py:327-331)."""
    return x.reshape(x.shape[0], -1) @ w.t() + b


def sigmoid_cross_entropy_loss(logits, targets, scale=1.0):
    """Detectron SigmoidCrossEntropyLoss (resnet_video.py:336-337):
    per element (t != -1): -(x*(t-(x>=0)) - log(1+exp(x-2x(x>=0)))), summed,
    divided by max(#valid, 1e-5), times scale."""
    x = logits
    t = targets.to(x.dtype)
    valid = (t != -1).to(x.dtype)
    ge = (x >= 0).to(x.dtype)
    per = -(x * (t - ge) - torch.log(1 + torch.exp(x - 2 * x * ge)))
    total = (per * valid).sum()
    norm = torch.clamp(valid.sum(), min=1e-5)
    return total / norm * scale


def softmax_with_loss(logits, labels, scale=1.0):
    """Caffe2 SoftmaxWithLoss: mean over batch of -log p[label], times scale (resnet_video.py:339-340)."""
    logp = torch.log_softmax(logits, dim=1)
    n = logits.shape[0]
    loss = -logp[torch.arange(n), labels.long().view(-1)].sum() / n * scale
    return torch.softmax(logits, dim=1), loss


def nesterov_update(p, g, m, lr, momentum, wd):
    """WeightedSum(g,1,p,wd) then MomentumSGDUpdate(nesterov=1)
    (model_builder_video.py:375-388).  Returns (p', m')."""
    g = g + wd * p
    m_new = momentum * m + lr * g
    ng = (1.0 + momentum) * m_new - momentum * m
    return p - ng, m_new


def msra_std(w_shape):
    """Caffe2 MSRAFill: std = sqrt(2 / fan_out), fan_out = size / dim(1)."""
    fan_out = int(np.prod(w_shape)) // int(w_shape[1])
    return math.sqrt(2.0 / fan_out)


def bn_fold(scale, bias, mean, var, eps=1e-5):
    """BN -> Affine fold (lib/utils/checkpoints.py:108-110)."""
    s = scale / np.sqrt(var + eps)
    return s, bias - mean * s


def inflate_2d_to_3d(w2d, kt):
    """2D -> 3D weight inflation (lib/utils/checkpoints.py:359-362)."""
    return np.stack([w2d] * kt, axis=2) / float(kt)
