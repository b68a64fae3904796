"""world_size-2 data-parallel test on CPU (gloo): the N>1 path of vlfb.dist -- parameter broadcast,
bucketed gradient all-reduce (sum; loss pre-scaled by 1/NUM_GPUS) and the local fused SGD step --
must reproduce the single-process step on the concatenated batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))

TINY = ['TRAIN.CROP_SIZE', 32, 'TRAIN.VIDEO_LENGTH', 8, 'LFB.WINDOW_SIZE', 2, 'TRAIN.DROPOUT_RATE', 0.0,
        'FBO_NL.INPUT_DROPOUT_ON', False, 'FBO_NL.LFB_DROPOUT_ON', False]
CHECK = ['conv1_w', 'res3_1_branch2b_w', 'nonlocal_conv4_1_out_w', 'lfb_1x1_w', 'lfb_nl1_theta_b', 'pred_w', 'pred_b']


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _step(rank, world, clips_lo, clips_hi, all_inputs, params, out):
    import fake_kernels
    import harness as H
    from vlfb import dist as vdist
    from vlfb import workspace
    fake_kernels.install()
    workspace.ResetWorkspace()
    n = clips_hi - clips_lo
    ov = TINY + ['NUM_GPUS', world, 'TRAIN.BATCH_SIZE', n * world]
    H.setup_cfg('ava_r50_lfb_nl.yaml', ov)
    model, sfx = H.build('train', True)
    if rank == 0:
        H.feed_params(params)
    vdist.install(workspace.current())
    vdist.broadcast_params(workspace.current().params)
    rois = all_inputs['proposals']
    sel = (rois[:, 0] >= clips_lo) & (rois[:, 0] < clips_hi)
    inputs = {'data': all_inputs['data'][clips_lo:clips_hi], 'labels': all_inputs['labels'][sel],
              'lfb': all_inputs['lfb'][sel]}
    pr = rois[sel].clone()
    pr[:, 0] -= clips_lo
    inputs['proposals'] = pr
    H.feed_inputs(inputs, sfx)
    model.UpdateWorkspaceLr(10)
    workspace.RunNet(model.net.Proto().name)
    for name in CHECK:
        out[name] = workspace.FetchBlob('gpu_0/' + name).copy()
    out['loss'] = float(workspace.FetchBlob('gpu_0/loss'))


def _worker(rank, world, port, all_inputs, params, ret):
    sys.path.insert(0, HERE)
    import conftest  # noqa: F401  (sys.path setup)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from core import config as C
    C.reset_cfg()
    from vlfb import dist as vdist
    vdist.init_from_env('gloo')
    out = {}
    _step(rank, world, rank * 1, rank * 1 + 1, all_inputs, params, out)
    ret[rank] = out
    torch.distributed.destroy_process_group()


def test_two_rank_step_equals_single_process_step():
    import harness as H
    from oracle import model as OM
    ov = TINY + ['NUM_GPUS', 1, 'TRAIN.BATCH_SIZE', 2]
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', ov)
    params = OM.make_params(ocfg, seed=2)
    all_inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=32, frames=8)
    # single process, both clips
    single = {}
    _step(0, 1, 0, 2, all_inputs, params, single)
    import fake_kernels
    fake_kernels.uninstall()
    # two ranks, one clip each
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, all_inputs, params, ret), nprocs=2, join=True)
    r0, r1 = ret[0], ret[1]
    assert abs((r0['loss'] + r1['loss']) - single['loss']) < 1e-9 * abs(single['loss'])
    for name in CHECK:
        assert np.array_equal(r0[name], r1[name]), name                      # replicas stay in sync
        d = np.abs(r0[name] - single[name]).max() / max(np.abs(single[name]).max(), 1e-12)
        assert d < 1e-9, (name, d)                                           # == the big-batch step
        assert np.abs(single[name] - params[name].numpy()).max() > 0         # and the step moved the weights


def _bn_rank_worker(rank, world, port, ret):
    sys.path.insert(0, HERE)
    import conftest  # noqa: F401  (sys.path setup)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group('gloo', rank=rank, world_size=world)
    from utils import bn_helper
    ret[rank] = bn_helper._mean_over_ranks(np.array([1.0 + rank, 10.0 * rank, -3.0], dtype=np.float64))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_precise_bn_statistics_are_averaged_over_ranks():
    """lib/utils/bn_helper.py: the reference divides E[x], E[x^2] by ITER * NUM_GPUS inside its one process
    (bn_helper.py:186); with one process per GPU the per-rank means are all-reduced instead."""
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bn_rank_worker, args=(2, port, ret), nprocs=2, join=True)
    for r in (0, 1):
        assert np.allclose(ret[r], [1.5, 5.0, -3.0])
