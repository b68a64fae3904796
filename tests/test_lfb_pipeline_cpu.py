"""The two-pass LFB recipe end to end on the CPU stand-in kernels (tools/lfb_loader.get_lfb :155-236 + training):
pass 1 runs the baseline net with lfb_infer_only=True and collects `box_pooled` + `metadata` into the bank
(construct_ava_lfb); the bank is packed into a device tensor (datasets.lfb_bank.DeviceLfb); pass 2 trains the FBO-NL
net with windows assembled on the device from index tables.  Checked against the oracle doing the same with the
reference's host-side dict + sample_lfb (oracle/lfb_sampling.py)."""
import numpy as np
import pytest
import torch

import harness as H
from oracle import lfb_sampling as OS

TINY = ['NUM_GPUS', 1, 'TRAIN.BATCH_SIZE', 2, 'TRAIN.CROP_SIZE', 64, 'TRAIN.VIDEO_LENGTH', 8,
        'TEST.BATCH_SIZE', 2, 'TEST.CROP_SIZE', 64, 'TEST.VIDEO_LENGTH', 8, 'LFB.WINDOW_SIZE', 4,
        'TRAIN.DROPOUT_RATE', 0.0, 'FBO_NL.INPUT_DROPOUT_ON', False, 'FBO_NL.LFB_DROPOUT_ON', False]


@pytest.fixture
def fake():
    import fake_kernels
    from vlfb import workspace
    fake_kernels.install()
    workspace.ResetWorkspace()
    yield fake_kernels
    workspace.ResetWorkspace()
    fake_kernels.uninstall()


def test_infer_bank_then_train_with_device_bank(fake):
    from datasets import lfb_bank as LB
    from oracle import model as OM
    from vlfb import workspace
    from core.config import config as cfg
    # ---- pass 1: baseline model, lfb_infer_only, three batches of two clips (3 boxes each) of one video
    ov = TINY
    H.setup_cfg('ava_r50_baseline.yaml', ov)
    ocfg = H.oracle_cfg('ava_r50_baseline.yaml', ov)
    base = OM.make_params(ocfg, seed=2, split='val', lfb_infer_only=True)
    model, sfx = H.build('val', False, suffix='_infer_train', lfb_infer_only=True)
    H.feed_params(base)
    feats, metas, ofeats = [], [], []
    for it in range(3):
        inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=3, crop=64, frames=8, seed=10 + it)
        secs = np.repeat([902 + 2 * it, 903 + 2 * it], 3)
        meta = np.stack([np.full(6, 7.0), secs.astype(np.float64), np.zeros(6), np.zeros(6)], 1)
        H.feed_inputs(inputs, sfx)
        workspace.FeedBlob('gpu_0/metadata' + sfx, meta.astype(np.float32))
        workspace.RunNet(model.net.Proto().name)
        feats.append([workspace.FetchBlob('gpu_0/box_pooled')])
        metas.append([workspace.FetchBlob('gpu_0/metadata' + sfx)])
        blobs, _, _ = OM.forward(ocfg, dict((k, v.double()) for k, v in base.items()),
                                 dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items()),
                                 'val', lfb_infer_only=True)
        ofeats.append([blobs['box_pooled'].numpy()])
        assert H.rel(feats[-1][0], ofeats[-1][0]) < 1e-9
    lfb = LB.construct_ava_lfb(feats, metas)
    olfb = OS.construct_ava_lfb(ofeats, metas)
    assert sorted(lfb) == [7] and sorted(lfb[7]) == list(range(902, 908)) and all(len(v) == 3 for v in lfb[7].values())
    dim = cfg.LFB.LFB_DIM
    bank = LB.DeviceLfb(lfb, dim)
    assert bank.rows == 18 and tuple(bank.bank.shape) == (18, dim)

    # ---- pass 2: the FBO-NL model trains on windows gathered from the device bank
    H.setup_cfg('ava_r50_lfb_nl.yaml', ov)
    ocfg2 = H.oracle_cfg('ava_r50_lfb_nl.yaml', ov)
    workspace.ResetWorkspace()
    params = OM.make_params(ocfg2, seed=2)
    model2, sfx2 = H.build('train', True)
    H.feed_params(params)
    inputs = OM.make_inputs(ocfg2, n_clips=2, rois_per_clip=2, crop=64, frames=8, seed=40)
    W, Kf = cfg.LFB.WINDOW_SIZE, cfg.AVA.LFB_MAX_NUM_FEAT_PER_STEP
    clip_secs = [903, 906]                                   # one window per clip, duplicated for each of its boxes
    np.random.seed(11)
    idx = np.stack([bank.sample_indices_ava(7, s, W, Kf) for s in clip_secs for _ in range(2)])
    np.random.seed(11)
    host = np.stack([OS.sample_lfb_ava(olfb[7], s, W, Kf, dim) for s in clip_secs for _ in range(2)])
    assert (idx >= 0).sum() > 0 and (idx < 0).sum() > 0        # real rows and zero padding both occur
    H.feed_inputs(dict((k, v) for k, v in inputs.items() if k != 'lfb'), sfx2)
    bank.feed('gpu_0/lfb' + sfx2, idx)
    net = workspace.current().nets[model2.net.Proto().name]
    net.update_ops = []
    workspace.RunNet(model2.net.Proto().name)
    oin = dict(inputs)
    oin['lfb'] = torch.from_numpy(host)
    p64 = dict((k, v.double().requires_grad_(True)) for k, v in params.items())
    blobs, _, loss = OM.forward(ocfg2, p64, dict((k, (v.double() if v.dtype in (torch.float32, torch.float64) else v))
                                                 for k, v in oin.items()), 'train')
    loss.backward()
    # the device bank stores fp32 rows (as the reference's pickle does); the oracle's dict holds its fp64 features
    assert H.rel(workspace.FetchBlob('gpu_0/loss'), loss.item()) < 1e-6
    assert H.rel(workspace.FetchBlob('gpu_0/lfb_nl1_sum'), blobs['lfb_nl1_sum'].detach().numpy()) < 1e-6
    assert H.rel(workspace.FetchBlob('gpu_0/lfb_1x1_w_grad'), p64['lfb_1x1_w'].grad.numpy()) < 1e-5
