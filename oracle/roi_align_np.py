"""Legacy (non-"aligned") Caffe2 RoIAlign restated in numpy float32 (test infrastructure).

Follows the operator the reference emits at lib/models/lfb_helper.py:144-150
(pooled 7x7, spatial_scale 1/16, sampling_ratio 0).  Caffe2's roi_align_op is an
un-vendored dependency; the algorithm below is its published definition as
summarised in SURVEY.md section 8a row a11.  All index arithmetic is carried out in
IEEE fp32 with one rounding per operation (no FMA), evaluated in the order written,
so that the CUDA kernel (which uses __fmul_rn/__fadd_rn/__fdiv_rn in the same
order) can be compared bit-for-bit.
"""
import numpy as np

F32 = np.float32


def sample_table(rois, height, width, pooled_h=7, pooled_w=7, spatial_scale=1.0 / 16.0,
                 sampling_ratio=0):
    """Return the bilinear sample table of every RoI.

    Output: list over rois of dict(grid_h, grid_w, pos (PH,PW,GH,GW,4) int32 flat
    indices y*W+x [-1 where the sample is outside], w (PH,PW,GH,GW,4) float32).
    """
    rois = np.asarray(rois, dtype=F32)
    scale = F32(spatial_scale)
    out = []
    for r in range(rois.shape[0]):
        x1 = F32(rois[r, 1] * scale)
        y1 = F32(rois[r, 2] * scale)
        x2 = F32(rois[r, 3] * scale)
        y2 = F32(rois[r, 4] * scale)
        roi_w = np.maximum(F32(x2 - x1), F32(1.0))
        roi_h = np.maximum(F32(y2 - y1), F32(1.0))
        bin_h = F32(roi_h / F32(pooled_h))
        bin_w = F32(roi_w / F32(pooled_w))
        gh = int(sampling_ratio) if sampling_ratio > 0 else int(np.ceil(F32(roi_h / F32(pooled_h))))
        gw = int(sampling_ratio) if sampling_ratio > 0 else int(np.ceil(F32(roi_w / F32(pooled_w))))
        pos = np.full((pooled_h, pooled_w, gh, gw, 4), -1, dtype=np.int32)
        wts = np.zeros((pooled_h, pooled_w, gh, gw, 4), dtype=F32)
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                for iy in range(gh):
                    # y = y1 + ph*bin_h + ((iy+.5)*bin_h)/gh
                    yy = F32(F32(y1 + F32(F32(ph) * bin_h)) +
                             F32(F32(F32(F32(iy) + F32(0.5)) * bin_h) / F32(gh)))
                    for ix in range(gw):
                        xx = F32(F32(x1 + F32(F32(pw) * bin_w)) +
                                 F32(F32(F32(F32(ix) + F32(0.5)) * bin_w) / F32(gw)))
                        p, w = _bilinear(yy, xx, height, width)
                        pos[ph, pw, iy, ix] = p
                        wts[ph, pw, iy, ix] = w
        out.append(dict(batch=int(rois[r, 0]), grid_h=gh, grid_w=gw, pos=pos, w=wts))
    return out


def _bilinear(y, x, height, width):
    if y < F32(-1.0) or y > F32(height) or x < F32(-1.0) or x > F32(width):
        return (-1, -1, -1, -1), (F32(0), F32(0), F32(0), F32(0))
    if y <= 0:
        y = F32(0)
    if x <= 0:
        x = F32(0)
    y_low = int(y)
    x_low = int(x)
    if y_low >= height - 1:
        y_high = y_low = height - 1
        y = F32(y_low)
    else:
        y_high = y_low + 1
    if x_low >= width - 1:
        x_high = x_low = width - 1
        x = F32(x_low)
    else:
        x_high = x_low + 1
    ly = F32(y - F32(y_low))
    lx = F32(x - F32(x_low))
    hy = F32(F32(1.0) - ly)
    hx = F32(F32(1.0) - lx)
    w = (F32(hy * hx), F32(hy * lx), F32(ly * hx), F32(ly * lx))
    p = (y_low * width + x_low, y_low * width + x_high, y_high * width + x_low, y_high * width + x_high)
    return p, w


def roi_align(feat, rois, pooled_h=7, pooled_w=7, spatial_scale=1.0 / 16.0, sampling_ratio=0):
    """feat (N,C,H,W) float32, rois (R,5) -> (R,C,PH,PW) float32 (mean of samples per bin)."""
    feat = np.asarray(feat, dtype=F32)
    n, c, h, w = feat.shape
    table = sample_table(rois, h, w, pooled_h, pooled_w, spatial_scale, sampling_ratio)
    out = np.zeros((len(table), c, pooled_h, pooled_w), dtype=F32)
    flat = feat.reshape(n, c, h * w)
    for r, t in enumerate(table):
        f = flat[t['batch']]
        cnt = F32(t['grid_h'] * t['grid_w'])
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                acc = np.zeros((c,), dtype=F32)
                for iy in range(t['grid_h']):
                    for ix in range(t['grid_w']):
                        p = t['pos'][ph, pw, iy, ix]
                        if p[0] < 0:
                            continue
                        wt = t['w'][ph, pw, iy, ix]
                        acc += wt[0] * f[:, p[0]] + wt[1] * f[:, p[1]] + wt[2] * f[:, p[2]] + wt[3] * f[:, p[3]]
                out[r, :, ph, pw] = acc / cnt
    return out
