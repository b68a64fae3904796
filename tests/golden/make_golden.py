"""Generates tests/golden/*.npz from the oracle (the reference itself cannot run: Python 2 + Caffe2).
Regression fixtures: they freeze the oracle's outputs on seeded inputs so that neither the oracle nor
the CUDA path can drift silently.  Run from the repo root: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib')]

import harness as H  # noqa: E402
from oracle import model as OM  # noqa: E402
from oracle import roi_align_np  # noqa: E402

TINY = ['NUM_GPUS', 1, 'TRAIN.BATCH_SIZE', 2, 'TRAIN.CROP_SIZE', 64, 'TRAIN.VIDEO_LENGTH', 8,
        'LFB.WINDOW_SIZE', 4, 'TRAIN.DROPOUT_RATE', 0.0, 'FBO_NL.INPUT_DROPOUT_ON', False,
        'FBO_NL.LFB_DROPOUT_ON', False]


def main():
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', TINY)
    params = OM.make_params(ocfg, seed=2)
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    p64 = dict((k, v.double().requires_grad_(True)) for k, v in params.items())
    i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items())
    blobs, prob, loss = OM.forward(ocfg, p64, i64, 'train')
    loss.backward()
    out = {'loss': np.float64(loss.item())}
    for b in ['box_pooled', 'pool5', 'pred', 'prob', 'lfb_nl1_sum']:
        out['blob/' + b] = blobs[b].detach().numpy().astype(np.float32)
    for pn in ['pred_w', 'pred_b', 'conv1_w', 'lfb_nl1_theta_b']:
        out['grad/' + pn] = p64[pn].grad.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'tiny_ava_fbo_nl.npz'), **out)

    rois = np.array([[0, 0., 0., 223., 223.], [1, 10.5, 20.25, 100.75, 180.5], [0, 3.3, 4.4, 3.9, 5.0],
                     [1, 111., 7., 223., 60.], [0, 200., 200., 223., 223.]], dtype=np.float32)
    table = roi_align_np.sample_table(rois, 14, 14, 7, 7, 1.0 / 16, 0)
    np.savez_compressed(os.path.join(HERE, 'roi_table_14x14.npz'), rois=rois,
                        grid=np.array([[t['grid_h'], t['grid_w']] for t in table], dtype=np.int32),
                        **dict(('pos%d' % i, t['pos']) for i, t in enumerate(table)),
                        **dict(('w%d' % i, t['w']) for i, t in enumerate(table)))
    print('wrote', os.listdir(HERE))


if __name__ == '__main__':
    main()
