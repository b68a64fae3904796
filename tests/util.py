"""Test helpers (CPU side)."""
import numpy as np
import torch


def tf32_round(t):
    """Round-to-nearest (ties away) to TF32, like cvt.rna.tf32.f32."""
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def rel_err(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md section 7."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def to_cl(x):
    """NCTHW -> contiguous NDHWC."""
    return x.permute(0, 2, 3, 4, 1).contiguous()


def to_nc(x):
    """NDHWC -> contiguous NCTHW."""
    return x.permute(0, 4, 1, 2, 3).contiguous()


def w_to_cl(w):
    """(Co,Ci,kT,kH,kW) -> [Co,kT,kH,kW,Ci]."""
    return w.permute(0, 2, 3, 4, 1).contiguous()


def stem_pack(w):
    """(Co,3,kT,kH,kW<=8) -> [Co,kT,kH,8,4] zero padded (conv1 layout)."""
    co, ci, kt, kh, kw = w.shape
    out = torch.zeros((co, kt, kh, 8, 4), dtype=w.dtype)
    out[:, :, :, :kw, :ci] = w.permute(0, 2, 3, 4, 1)
    return out
