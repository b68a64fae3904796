"""Checkpoint interchange: reference pickle format in and out, BN -> Affine fold, 2D -> 3D
inflation, classifier / momentum rules of lib/utils/checkpoints.py, against the oracle's closed forms (oracle/ops.py)."""
import os
import pickle

import numpy as np
import pytest
import torch

import harness as H

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


def test_save_load_round_trip_and_resume(fake, tmp_path):
    from oracle import model as OM
    from utils import checkpoints as CK
    from vlfb import workspace
    H.setup_cfg('ava_r50_lfb_nl.yaml', TINY)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', TINY)
    params = OM.make_params(ocfg, seed=2)
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    model, sfx = H.build('train', True)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    model.UpdateWorkspaceLr(10)
    workspace.RunNet(model.net.Proto().name)             # one SGD step: momentum becomes non-zero
    path = os.path.join(str(tmp_path), 'c2_model_iter10.pkl')
    CK.save_model_params(model, path, 9)
    with open(path, 'rb') as f:
        saved = pickle.load(f)['blobs']
    assert saved['model_iter'] == 10 and abs(float(saved['lr']) - float(workspace.FetchBlob('gpu_0/lr'))) < 1e-12
    assert saved['conv1_w'].shape == (64, 3, 5, 7, 7) and saved['res4_0_branch2a_w'].shape == (256, 512, 3, 1, 1)
    assert 'res2_0_branch2a_bn_s' in saved and 'res2_0_branch2a_bn_s_momentum' not in saved
    assert np.abs(saved['pred_w_momentum']).max() > 0
    want = dict((k, workspace.FetchBlob('gpu_0/' + k).copy()) for k in
                ['conv1_w', 'res3_1_branch2b_w', 'lfb_nl1_out_w', 'pred_b', 'res5_2_branch2c_bn_b', 'pred_w_momentum',
                 'conv1_w_momentum'])
    pred0 = workspace.FetchBlob('gpu_0/pred').copy()
    # fresh workspace + model: load restores parameters, momentum, lr and the iteration counter
    H.setup_cfg('ava_r50_lfb_nl.yaml', TINY)
    workspace.ResetWorkspace()
    model2, sfx = H.build('train', True)
    it, lr = CK.initialize_params_from_file(model2, path)
    assert it == 10 and abs(lr - float(saved['lr'])) < 1e-12
    for k, v in want.items():
        assert np.allclose(workspace.FetchBlob('gpu_0/' + k), v, rtol=1e-6, atol=1e-10), k   # fp32 file, fp64 test engine
    # a test net never receives momentum and reproduces the forward of the saved weights
    H.setup_cfg('ava_r50_lfb_nl.yaml', TINY)
    workspace.ResetWorkspace()
    model3, sfx3 = H.build('val', False)
    CK.load_model_from_params_file_for_test(model3, path)
    assert np.allclose(workspace.FetchBlob('gpu_0/res3_1_branch2b_w'), want['res3_1_branch2b_w'], rtol=1e-6, atol=1e-10)


def test_kinetics_conversion_fold_inflate_and_classifier_rules(fake, tmp_path):
    """A synthetic 2-D 'pre-trained classification' checkpoint with trainable-BN statistics -> convert_model ->
    load: BN folded (eps 1e-5), temporal kernels inflated (kT copies / kT), classifier dropped, lr = 0.00125."""
    from oracle import ops as O
    from utils import checkpoints as CK
    from vlfb import workspace
    H.setup_cfg('ava_r50_baseline.yaml', TINY)
    model, sfx = H.build('train', True)
    store = workspace.current().params
    rng = np.random.RandomState(0)
    src = {}
    for name in store.index:
        shape = store.logical_shape(name)
        if name.endswith('_bn_s') or name.endswith('_bn_b'):
            layer = name[:-len('_bn_s')]
            if layer + '_bn_rm' not in src:
                c = shape[0]
                src[layer + '_bn_s'], src[layer + '_bn_b'] = rng.rand(c) + 0.5, rng.randn(c) * 0.1
                src[layer + '_bn_rm'], src[layer + '_bn_riv'] = rng.randn(c) * 0.2, rng.rand(c) + 0.3
        elif len(shape) == 5 and shape[2] > 1 and 'conv1' not in name:
            src[name] = rng.randn(shape[0], shape[1], shape[3], shape[4]).astype(np.float32)     # 2-D kernel: inflate
        else:
            src[name] = rng.randn(*shape).astype(np.float32)
    src['pred_w'] = rng.randn(400, 2048).astype(np.float32)           # Kinetics classifier: 400 classes
    src['pred_b'] = rng.randn(400).astype(np.float32)
    src['conv1_w_momentum'] = np.ones((64, 3, 5, 7, 7), np.float32)
    src.update(lr=0.1, model_iter=1234, epoch=3)
    path = os.path.join(str(tmp_path), 'r50_k400.pkl')
    with open(path, 'wb') as f:
        pickle.dump({'blobs': src}, f, 2)
    init = CK.convert_model(path, out_dir=str(tmp_path))
    with open(init, 'rb') as f:
        conv = pickle.load(f)
    assert conv['lr'] == 0.00125 and not any('pred' in k or 'momentum' in k or k.endswith('_bn_rm') for k in conv)
    assert 'model_iter' not in conv and 'epoch' not in conv
    pw0 = workspace.FetchBlob('gpu_0/pred_w').copy()
    it, lr = CK.initialize_params_from_file(model, init)
    assert it == 0 and lr == 0.00125 and abs(float(workspace.FetchBlob('gpu_0/lr')) - 0.00125) < 1e-9
    layer = 'res3_0_branch2b'
    s, b = O.bn_fold(src[layer + '_bn_s'], src[layer + '_bn_b'], src[layer + '_bn_rm'], src[layer + '_bn_riv'])
    assert np.allclose(workspace.FetchBlob('gpu_0/' + layer + '_bn_s'), s, rtol=1e-6)
    assert np.allclose(workspace.FetchBlob('gpu_0/' + layer + '_bn_b'), b, rtol=1e-6, atol=1e-7)
    name = 'res4_0_branch2a_w'                                         # (256, 512, 3, 1, 1) from a (256, 512, 1, 1) file
    assert src[name].ndim == 4
    assert np.allclose(workspace.FetchBlob('gpu_0/' + name), O.inflate_2d_to_3d(src[name], 3), rtol=1e-6)
    assert np.allclose(workspace.FetchBlob('gpu_0/conv1_w'), src['conv1_w'], rtol=1e-6)
    assert np.array_equal(workspace.FetchBlob('gpu_0/pred_w'), pw0)   # classifier untouched (not in the file)
    # a file that still carries a 400-way classifier: found but unmatching -> skipped, not an error
    with open(init, 'rb') as f:
        blobs = pickle.load(f)
    blobs['pred_w'], blobs['pred_b'] = src['pred_w'], src['pred_b']
    p2 = os.path.join(str(tmp_path), 'with_pred.pkl')
    with open(p2, 'wb') as f:
        pickle.dump(blobs, f, 2)
    CK.initialize_params_from_file(model, p2)
    assert np.array_equal(workspace.FetchBlob('gpu_0/pred_w'), pw0)
    # no lr blob and no RESET_START_ITER -> the reference raises
    del blobs['lr']
    with open(p2, 'wb') as f:
        pickle.dump(blobs, f, 2)
    with pytest.raises(Exception):
        CK.initialize_params_from_file(model, p2)


def test_resume_rules_follow_the_reference(fake, tmp_path):
    """load_model_from_params_file (reference lib/utils/checkpoints.py:180-230): cases 1, 2a, 2b, 3a, 3b, the
    '<DIR>/checkpoints/c2_model_iter{N}.pkl' layout, no momentum from a pre-trained PARAMS_FILE, current_lr, and the
    RESUME_FROM_BATCH_SIZE / RESET_START_ITER corrections."""
    from oracle import model as OM
    from utils import checkpoints as CK
    from vlfb import workspace
    from core.config import config as cfg
    base = str(tmp_path)
    ov = TINY + ['CHECKPOINT.DIR', base]
    H.setup_cfg('ava_r50_lfb_nl.yaml', ov)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', TINY)
    params = OM.make_params(ocfg, seed=2)
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    model, sfx = H.build('train', True)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    model.UpdateWorkspaceLr(10)
    workspace.RunNet(model.net.Proto().name)
    ckdir = CK.create_and_get_checkpoint_directory()
    assert ckdir == os.path.join(os.path.abspath(base), 'checkpoints') and os.path.isdir(ckdir)
    assert not CK.find_checkpoint() and CK.get_checkpoint_resume_file() is None
    pre = os.path.join(base, 'pretrained.pkl')
    CK.save_model_params(model, pre, 99)                                # "pre-trained" file: iter 100, has momentum
    w_pre = workspace.FetchBlob('gpu_0/pred_w').copy()
    workspace.RunNet(model.net.Proto().name)                            # train on: the checkpoints differ from it
    CK.save_model_params(model, os.path.join(ckdir, 'c2_model_iter20.pkl'), 19)
    workspace.RunNet(model.net.Proto().name)
    CK.save_model_params(model, os.path.join(ckdir, 'c2_model_iter200.pkl'), 199)
    w_200 = workspace.FetchBlob('gpu_0/pred_w').copy()
    m_200 = workspace.FetchBlob('gpu_0/pred_w_momentum').copy()
    assert CK.find_checkpoint() and CK.get_checkpoint_resume_file().endswith('c2_model_iter200.pkl')   # numeric, not lexical

    def fresh(extra):
        H.setup_cfg('ava_r50_lfb_nl.yaml', ov + extra)
        workspace.ResetWorkspace()
        m, _ = H.build('train', True)
        return m

    # case 2a / 3a: RESUME and a checkpoint exists -> the latest checkpoint wins over PARAMS_FILE, momentum restored
    for extra in (['TRAIN.PARAMS_FILE', pre], []):
        m = fresh(extra)
        assert CK.load_model_from_params_file(m) == 200
        assert np.allclose(workspace.FetchBlob('gpu_0/pred_w'), w_200, rtol=1e-6, atol=1e-10)
        assert np.allclose(workspace.FetchBlob('gpu_0/pred_w_momentum'), m_200, rtol=1e-6, atol=1e-12)
        assert abs(m.current_lr - float(workspace.FetchBlob('gpu_0/lr'))) < 1e-9
    # case 1: RESUME False -> PARAMS_FILE, never its momentum; start iteration from the file
    m = fresh(['TRAIN.PARAMS_FILE', pre, 'CHECKPOINT.RESUME', False])
    assert CK.load_model_from_params_file(m) == 100
    assert np.allclose(workspace.FetchBlob('gpu_0/pred_w'), w_pre, rtol=1e-6, atol=1e-10)
    assert np.abs(workspace.FetchBlob('gpu_0/pred_w_momentum')).max() == 0
    # ... rescaled when the file was trained with another batch size, zeroed by RESET_START_ITER
    m = fresh(['TRAIN.PARAMS_FILE', pre, 'CHECKPOINT.RESUME', False, 'TRAIN.RESUME_FROM_BATCH_SIZE', 8])
    assert CK.load_model_from_params_file(m) == int(100 * 8 / cfg.TRAIN.BATCH_SIZE)
    m = fresh(['TRAIN.PARAMS_FILE', pre, 'CHECKPOINT.RESUME', False, 'TRAIN.RESET_START_ITER', True])
    assert CK.load_model_from_params_file(m) == 0
    # case 2b / 3b: RESUME but no checkpoint yet
    empty = os.path.join(base, 'empty')
    os.makedirs(empty)
    m = fresh(['TRAIN.PARAMS_FILE', pre, 'CHECKPOINT.DIR', empty])
    assert CK.load_model_from_params_file(m) == 100
    m = fresh(['CHECKPOINT.DIR', empty])
    pw = workspace.FetchBlob('gpu_0/pred_w').copy()
    assert CK.load_model_from_params_file(m) == 0 and np.array_equal(workspace.FetchBlob('gpu_0/pred_w'), pw)
    # a file from a differently named net loads nothing: that is an error, not a silent from-scratch run
    bad = os.path.join(base, 'renamed.pkl')
    with open(bad, 'wb') as f:
        pickle.dump({'blobs': {'lr': 0.1, 'something_else_w': np.zeros((3, 3), np.float32)}}, f, 2)
    m = fresh(['TRAIN.PARAMS_FILE', bad, 'CHECKPOINT.RESUME', False])
    with pytest.raises(Exception):
        CK.load_model_from_params_file(m)
