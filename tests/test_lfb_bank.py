"""Device-resident feature bank: the index tables + gather reproduce the reference's window
sampling (oracle/lfb_sampling.py restates ava.py:300-323, charades.py:251-276, epic.py:310-331) bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import lfb_sampling as OS

DIM = 64


def _ava_dict(seed=0, videos=3, secs=30):
    rng = np.random.RandomState(seed)
    feats, meta = [], []
    for it in range(4):
        f, m = [], []
        for gpu in range(2):
            r = rng.randint(1, 9)
            f.append(rng.randn(r, DIM, 1, 1, 1).astype(np.float32))
            m.append(np.stack([rng.randint(0, videos, r), rng.randint(900, 900 + secs, r), np.zeros(r), np.zeros(r)], 1).astype(np.float64))
        feats.append(f)
        meta.append(m)
    return feats, meta


def _check_ava(bank_cls, lfb, gather):
    for video in lfb:
        for sec in list(lfb[video])[:6] + [850, 1000]:
            for W, K in [(5, 3), (4, 5), (60, 5)]:
                np.random.seed(7)
                ref = OS.sample_lfb_ava(lfb[video], sec, W, K, DIM)
                np.random.seed(7)
                idx = bank_cls.sample_indices_ava(video, sec, W, K)
                got = gather(idx[None])[0]
                assert got.shape == ref.shape and np.array_equal(got, ref.astype(np.float32)), (video, sec, W, K)


@pytest.fixture
def fake():
    import fake_kernels
    from vlfb import workspace
    fake_kernels.install()
    workspace.ResetWorkspace()
    yield fake_kernels
    workspace.ResetWorkspace()
    fake_kernels.uninstall()


def test_construct_and_roundtrip(tmp_path):
    from datasets import lfb_bank as LB
    feats, meta = _ava_dict()
    a, b = LB.construct_ava_lfb(feats, meta), OS.construct_ava_lfb(feats, meta)
    assert sorted(a) == sorted(b)
    for v in a:
        assert sorted(a[v]) == sorted(b[v])
        for s in a[v]:
            assert len(a[v][s]) == len(b[v][s]) and all(np.array_equal(x, y) for x, y in zip(a[v][s], b[v][s]))
    path = os.path.join(str(tmp_path), 'train_lfb.pkl')
    LB.write_lfb(a, path)
    c = LB.load_lfb(path)
    assert all(np.array_equal(c[v][s][0], a[v][s][0]) for v in a for s in a[v])
    md = [(v, f) for v in range(2) for f in range(5, 200, 6)]
    ff = [[np.random.RandomState(1).randn(len(md) + 3, DIM).astype(np.float32)]]
    fa, fb = LB.construct_frame_level_lfb(ff, md), OS.construct_frame_level_lfb(ff, md)
    assert all(np.array_equal(fa[v][f], fb[v][f]) for v in fb for f in fb[v]) and len(fa) == len(fb) == 2


def test_index_tables_reproduce_reference_sampling_cpu(fake):
    from datasets import lfb_bank as LB
    feats, meta = _ava_dict()
    lfb = LB.construct_ava_lfb(feats, meta)
    bank = LB.DeviceLfb(lfb, DIM)
    assert bank.rows == sum(len(r) for v in lfb.values() for r in v.values())
    _check_ava(bank, lfb, lambda idx: bank.gather(idx).float().numpy())
    # frame-level banks
    md = [(v, f) for v in range(2) for f in range(5, 400, 6)]
    ff = [[np.random.RandomState(1).randn(len(md), DIM).astype(np.float32)]]
    flfb = LB.construct_frame_level_lfb(ff, md)
    fbank = LB.DeviceLfb(flfb, DIM)
    for center in (0, 50, 211, 390, 1000):
        for W, cps in [(20, 2), (40, 2), (120, 2)]:
            ref = OS.sample_lfb_charades(flfb[1], center, W, cps, DIM)
            got = fbank.gather(fbank.sample_indices_charades(1, center, W, cps)[None])[0].float().numpy()
            assert np.array_equal(got, ref.astype(np.float32))
        ref = OS.sample_verb_lfb_epic(center, flfb[0], 10, DIM)
        got = fbank.gather(fbank.sample_indices_epic_verb(0, center, 10)[None])[0].float().numpy()
        assert np.array_equal(got, ref)


def test_epic_noun_bank_with_several_detections_per_frame(fake):
    from datasets import lfb_bank as LB
    rng = np.random.RandomState(4)
    video = {}
    for f in range(0, 900, 24):                          # NOUN_LFB_FRAMES_PER_SECOND = 1 at 24 fps
        n = int(rng.randint(0, 14))
        video[f] = rng.randn(n, DIM).astype(np.float32) if n else []
    bank = LB.DeviceLfb({'P01_01': video}, DIM)
    assert bank.rows == sum(v.shape[0] for v in video.values() if not isinstance(v, list))
    for center in (0, 100, 433, 880, 2000):
        for W, per_frame, fps in [(60, 10, 1), (40, 10, 1), (30, 3, 1), (7, 10, 1)]:
            ref = OS.sample_noun_lfb_epic(center, video, W, per_frame, fps, DIM)
            idx = bank.sample_indices_epic_noun('P01_01', center, W, per_frame, fps)
            got = bank.gather(idx[None])[0].float().numpy()
            assert got.shape == ref.shape and np.array_equal(got, ref.astype(np.float32)), (center, W, per_frame)


def test_feed_matches_feedblob_of_host_windows(fake):
    """DeviceLfb.feed(blob, indices) leaves the workspace exactly as FeedBlob(blob, host-assembled windows) does."""
    from datasets import lfb_bank as LB
    from vlfb import workspace
    feats, meta = _ava_dict(3)
    lfb = LB.construct_ava_lfb(feats, meta)
    bank = LB.DeviceLfb(lfb, DIM)
    video = sorted(lfb)[0]
    secs = sorted(lfb[video])[:4]
    np.random.seed(3)
    host = np.stack([OS.sample_lfb_ava(lfb[video], s, 5, 3, DIM) for s in secs]).astype(np.float32)
    np.random.seed(3)
    idx = np.stack([bank.sample_indices_ava(video, s, 5, 3) for s in secs])
    workspace.FeedBlob('gpu_0/lfb_a', host)
    bank.feed('gpu_0/lfb_b', idx)
    assert np.array_equal(workspace.FetchBlob('gpu_0/lfb_a'), workspace.FetchBlob('gpu_0/lfb_b'))


@pytest.mark.gpu
def test_device_bank_gather_gpu():
    from datasets import lfb_bank as LB
    from vlfb import workspace
    assert torch.cuda.is_available()
    workspace.ResetWorkspace()
    feats, meta = _ava_dict(5, videos=4, secs=40)
    lfb = LB.construct_ava_lfb(feats, meta)
    bank = LB.DeviceLfb(lfb, DIM)
    assert bank.bank.is_cuda
    _check_ava(bank, lfb, lambda idx: bank.gather(idx).cpu().numpy())
    video = sorted(lfb)[0]
    np.random.seed(3)
    secs = sorted(lfb[video])[:4]
    host = np.stack([OS.sample_lfb_ava(lfb[video], s, 5, 3, DIM) for s in secs]).astype(np.float32)
    np.random.seed(3)
    idx = np.stack([bank.sample_indices_ava(video, s, 5, 3) for s in secs])
    workspace.FeedBlob('gpu_0/lfb_a', host)
    bank.feed('gpu_0/lfb_b', idx)
    assert np.array_equal(workspace.FetchBlob('gpu_0/lfb_a'), workspace.FetchBlob('gpu_0/lfb_b'))
    workspace.ResetWorkspace()
