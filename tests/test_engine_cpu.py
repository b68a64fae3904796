"""Host-logic tests on CPU: the recorded net, its lowering/fusion, the autograd tape, the
ParamStore and the optimizer are run through a torch-CPU kernel stand-in (tests/fake_kernels.py)
and compared with the oracle (forward blobs, loss, every parameter gradient, one SGD step)."""
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


def _run(yaml_name, overrides, fake, check_blobs, reps=1):
    from oracle import model as OM
    from vlfb import workspace
    H.setup_cfg(yaml_name, overrides)
    ocfg = H.oracle_cfg(yaml_name, overrides)
    params = OM.make_params(ocfg, seed=2)
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    model, sfx = H.build('train', True)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    # oracle forward/backward in fp64
    p64 = dict((k, v.double().requires_grad_(True)) for k, v in params.items())
    i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items())
    blobs, prob, loss = OM.forward(ocfg, p64, i64, 'train')
    loss.backward()
    # product: forward + backward only (no update) to compare gradients
    net = workspace.current().nets[model.net.Proto().name]
    from vlfb import executor as X
    upd, net.update_ops = net.update_ops, []
    first = {}
    fuse0, X.FUSE_GRAD_FINISH = X.FUSE_GRAD_FINISH, True
    lazy0 = X.LAZY_GRAD_SUM
    # run 1: the default engine (deferred two-term gradient sums folded into the consumer's mask/round pass); it
    # also records how many contributions every gradient receives.  Run 2: the dgrad GEMM that delivers the last
    # contribution applies the ReLU backward + TF32 rounding itself (executor "grad finish" fusion, eager sums).
    # Both must match the oracle; they differ from each other only by the association of three-term sums.
    for rep in range(reps):
        X.LAZY_GRAD_SUM = lazy0 if rep == 0 else False
        fused0 = X.STATS['fused_grad_finish']
        workspace.RunNet(model.net.Proto().name)
        assert (X.STATS['fused_grad_finish'] > fused0) == (rep == 1)
        assert H.rel(workspace.FetchBlob('gpu_0/loss'), loss.item()) < 1e-9
        for b in check_blobs:
            assert H.rel(workspace.FetchBlob('gpu_0/' + b), blobs[b].detach().numpy()) < 1e-9, b
        trainable = model.TrainableParams()
        expected = [k for k in params if not (k.endswith('_bn_s') or k.endswith('_bn_b'))]
        assert sorted(trainable) == sorted(expected)
        worst = 0.0
        for name in trainable:
            g = workspace.FetchBlob('gpu_0/' + name + '_grad')
            ref = p64[name].grad.numpy()
            e = float(np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-5))   # e.g. phi_b has a zero gradient
            worst = max(worst, e)
            assert e < 1e-7, (name, e)
            if rep == 0:
                first[name] = g.copy()
            else:
                assert np.abs(g - first[name]).max() <= 1e-12 * max(np.abs(g).max(), 1e-5), name   # phi_b: zero gradient
        print("run %d: worst grad rel err %.2e" % (rep, worst))
    net.update_ops = upd
    X.FUSE_GRAD_FINISH, X.LAZY_GRAD_SUM = fuse0, lazy0
    return model, params, p64, worst


def test_ava_fbo_nl_forward_backward_matches_oracle(fake):
    _run('ava_r50_lfb_nl.yaml', TINY, fake,
         ['pool1', 'res2_2_branch2c_bn', 'nonlocal_conv3_1_sum', 'res3_3_branch2c_bn', 'nonlocal_conv4_1_sum',
          'res5_2_branch2c_bn', 'blob_pooled', 'roi_feat_3d', 'box_pooled', 'lfb_1x1', 'lfb_nl0_affinity_prob',
          'lfb_nl1_sum', 'pool5', 'pred', 'prob'], reps=2)


def test_charades_post_act_variant(fake):
    _run('charades_r50_lfb_nl.yaml', TINY + ['MODEL.NUM_CLASSES', 157], fake, ['pool5', 'pred'])


@pytest.mark.parametrize('yaml_name', ['ava_r50_lfb_avg.yaml', 'ava_r50_lfb_max.yaml', 'ava_r50_baseline.yaml',
                                       'ava_r101_lfb_nl_3l.yaml'])
def test_other_heads(fake, yaml_name):
    _run(yaml_name, TINY, fake, ['pool5', 'pred'])


def test_sgd_step_matches_oracle(fake):
    from oracle import ops as O
    from core.config import config as cfg
    from vlfb import workspace
    model, params, p64, _ = _run('ava_r50_lfb_nl.yaml', TINY, fake, [])
    model.UpdateWorkspaceLr(10)
    lr = float(workspace.FetchBlob('gpu_0/lr'))
    workspace.RunNet(model.net.Proto().name)          # full step (momentum starts at 0)
    for name in ['conv1_w', 'res4_3_branch2b_w', 'nonlocal_conv4_1_theta_b', 'lfb_nl1_out_w', 'pred_w', 'pred_b']:
        p_ref, m_ref = O.nesterov_update(params[name].double(), p64[name].grad, torch.zeros_like(p64[name]), lr,
                                         cfg.SOLVER.MOMENTUM, cfg.SOLVER.WEIGHT_DECAY)
        assert H.rel(workspace.FetchBlob('gpu_0/' + name), p_ref.detach().numpy()) < 1e-5, name
        assert H.rel(workspace.FetchBlob('gpu_0/' + name + '_momentum'), m_ref.detach().numpy()) < 1e-4, name
    # frozen affine parameters must not move
    assert np.array_equal(workspace.FetchBlob('gpu_0/res2_0_branch2a_bn_s'), params['res2_0_branch2a_bn_s'].numpy())


def test_test_split_graph_and_lfb_infer_only(fake):
    from oracle import model as OM
    from vlfb import workspace
    ov = TINY
    H.setup_cfg('ava_r50_lfb_nl.yaml', ov)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', ov)
    params = OM.make_params(ocfg, seed=2, split='val', lfb_infer_only=True)
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=3, crop=64, frames=8)
    model, sfx = H.build('val', False, lfb_infer_only=True)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    workspace.RunNet(model.net.Proto().name)
    blobs, _, _ = OM.forward(ocfg, params, inputs, 'val', lfb_infer_only=True)
    assert H.rel(workspace.FetchBlob('gpu_0/box_pooled'), blobs['box_pooled'].numpy()) < 1e-4
    assert not workspace.HasBlob('gpu_0/pred')


def test_fusion_plan(fake):
    """Every conv is lowered with its AffineNd (and Sum/Relu where the reference has them) fused."""
    from vlfb import executor as X
    from vlfb import workspace
    H.setup_cfg('ava_r50_lfb_nl.yaml', TINY)
    model, _ = H.build('train', True)
    net = workspace.current().nets[model.net.Proto().name]
    convs = [s for s in net.steps if isinstance(s, X.ConvStep)]
    by_name = dict((s.out, s) for s in convs)
    assert len([s for s in net.steps if isinstance(s, X.AffineStep)]) == 0
    assert by_name['res_conv1_bn'].relu and by_name['res_conv1_bn'].affine
    s = by_name['res3_1_branch2c_bn']
    assert s.relu and s.affine and s.res_key[0] == 'res3_0_branch2c_bn'
    s = by_name['res2_0_branch2c_bn']
    assert s.relu and s.res_key[0] == 'res2_0_branch1_bn'
    s = by_name['nonlocal_conv4_1_sum']
    assert s.affine and s.res_key[0] == 'res4_1_branch2c_bn' and not s.relu and s.b
    assert len([s for s in net.steps if isinstance(s, X.ScaleStep)]) == 0       # folded into the softmax
    # 53 backbone + 20 NL convs, reduc + lfb_1x1; the 8 theta / phi / g / out projections of the two FBO-NL layers
    # live inside ONE FboStackStep (B200.FBO_STACK), with the parameters the as-written graph creates
    assert len(convs) == 53 + 20 + 2
    stacks = [s for s in net.steps if isinstance(s, X.FboStackStep)]
    assert len(stacks) == 1 and len(stacks[0].params) == 16
    assert set(stacks[0].params) <= set(model.params) and 'lfb_nl1_out_w' in stacks[0].params
    from core.config import config as cfg
    H.setup_cfg('ava_r50_lfb_nl.yaml', TINY)
    cfg.B200.FBO_STACK = False
    workspace.ResetWorkspace()
    model2, _ = H.build('train', True)
    net2 = workspace.current().nets[model2.net.Proto().name]
    assert len([s for s in net2.steps if isinstance(s, X.ConvStep)]) == 53 + 20 + 2 + 8
    assert sorted(model2.params) == sorted(model.params)            # identical parameter inventory either way


def test_tf32_rounding_points_match_oracle_emulation(fake):
    """The engine rounds GEMM operands to TF32 at their producers.  With the
    CPU stand-in performing the same roundings, the forward must coincide with the oracle's
    `emulate_tf32` mode except for the rare elements that sit on a TF32 rounding boundary.
    (As-written lowering: the folded inference FBO, B200.FBO_FOLD, contracts in a different order and therefore
    rounds at different points; it is held to the fp64 oracle by test_inference_fbo_fold_*.)"""
    from oracle import model as OM
    from vlfb import workspace
    from core.config import config as cfg
    fake.EMULATE_TF32 = True
    try:
        H.setup_cfg('ava_r50_lfb_nl.yaml', TINY)
        cfg.B200.FBO_FOLD = False
        ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', TINY)
        params = OM.make_params(ocfg, seed=2)
        inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
        model, sfx = H.build('val', False)
        H.feed_params(dict((k, v) for k, v in params.items() if workspace.HasBlob(k)))
        H.feed_inputs(inputs, sfx)
        workspace.RunNet(model.net.Proto().name)
        p64 = dict((k, v.double()) for k, v in params.items())
        i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items())
        blobs, _, _ = OM.forward(ocfg, p64, i64, 'val', emulate_tf32=True)
        for name in ('pool1', 'res3_3_branch2c_bn', 'nonlocal_conv4_1_sum', 'res5_2_branch2c_bn', 'lfb_nl1_sum'):
            a = workspace.FetchBlob('gpu_0/' + name)
            b = blobs[name].numpy()
            differing = float((np.abs(a - b) > 1e-9 * np.abs(b).max()).mean())
            print(name, 'fraction of elements differing from the emulating oracle: %.2e' % differing)
            assert differing < 1e-3, (name, differing)          # one-ulp boundary cases only
            assert H.rel(a, b) < 2.5e-3, name                   # and never more than a TF32 ulp or two
    finally:
        fake.EMULATE_TF32 = False


def test_grad_finish_fusion_is_exact_under_tf32_emulation(fake):
    """Run 1 (separate ReLU-backward / rounding passes) and run 2 (folded into the last dgrad GEMM's epilogue)
    must produce bit-identical gradients also when the TF32 roundings are emulated."""
    from oracle import model as OM
    from vlfb import executor as X, workspace
    fake.EMULATE_TF32 = True
    fuse0, X.FUSE_GRAD_FINISH = X.FUSE_GRAD_FINISH, True
    lazy0, X.LAZY_GRAD_SUM = X.LAZY_GRAD_SUM, False      # same association of the sums in both runs
    try:
        H.setup_cfg('ava_r50_lfb_nl.yaml', TINY)
        ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', TINY)
        params = OM.make_params(ocfg, seed=2)
        inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
        model, sfx = H.build('train', True)
        H.feed_params(params)
        H.feed_inputs(inputs, sfx)
        net = workspace.current().nets[model.net.Proto().name]
        net.update_ops = []
        grads = []
        for rep in range(2):
            n0 = X.STATS['fused_grad_finish']
            workspace.RunNet(model.net.Proto().name)
            assert (X.STATS['fused_grad_finish'] - n0 > 40) == (rep == 1)
            grads.append(dict((n, workspace.FetchBlob('gpu_0/' + n + '_grad').copy()) for n in model.TrainableParams()))
        for n in grads[0]:
            assert np.array_equal(grads[0][n], grads[1][n]), n
    finally:
        fake.EMULATE_TF32 = False
        X.FUSE_GRAD_FINISH, X.LAZY_GRAD_SUM = fuse0, lazy0


@pytest.mark.parametrize('yaml_name,layers', [('ava_r50_lfb_nl.yaml', 2), ('ava_r50_lfb_nl_3l.yaml', 3),
                                               ('charades_r50_lfb_nl.yaml', 2)])
def test_inference_fbo_fold_matches_oracle_and_unfolded_graph(fake, yaml_name, layers):
    """Test-mode nets run every FBO-NL layer as one pass over the raw bank (executor.FboFoldStep); the result must
    equal the oracle's as-written graph and the engine's own unfolded lowering."""
    from oracle import model as OM
    from vlfb import workspace
    from vlfb import executor as X
    from core.config import config as cfg
    ov = TINY + (['MODEL.NUM_CLASSES', 157] if 'charades' in yaml_name else [])
    ocfg = H.oracle_cfg(yaml_name, ov)
    params = OM.make_params(ocfg, seed=2, split='val')
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=3, crop=64, frames=8)
    blobs, _, _ = OM.forward(ocfg, dict((k, v.double()) for k, v in params.items()),
                             dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items()), 'val')
    got = {}
    for fold in (True, False):
        H.setup_cfg(yaml_name, ov)
        cfg.B200.FBO_FOLD = fold
        workspace.ResetWorkspace()
        model, sfx = H.build('val', False)
        H.feed_params(params)
        H.feed_inputs(inputs, sfx)
        net = workspace.current().nets[model.net.Proto().name]
        assert sum(isinstance(s, X.FboFoldStep) for s in net.steps) == (layers if fold else 0)
        workspace.RunNet(model.net.Proto().name)
        for b in ['lfb_nl0_affinity_prob', 'lfb_nl%d_sum' % (layers - 1), 'pool5', 'pred', 'prob']:
            got[(fold, b)] = workspace.FetchBlob('gpu_0/' + b)
            assert H.rel(got[(fold, b)].reshape(-1), blobs[b].detach().numpy().reshape(-1)) < 1e-9, (fold, b)
        assert workspace.HasBlob('gpu_0/lfb_1x1') == (not fold)


def test_inference_fbo_fold_on_a_bf16_bank(fake):
    """B200.LFB_DTYPE 'bf16': feeding the bank also stores a bf16 copy, the folded FBO scans that copy.  The result
    equals the oracle evaluated on the bf16-rounded bank exactly (storage type only: fp32 arithmetic) and the fp32-bank
    oracle within the bf16 tolerance of BASELINE configs[3]/[4] (1e-2)."""
    from oracle import model as OM
    from vlfb import workspace
    from vlfb import executor as X
    from core.config import config as cfg
    yaml_name = 'ava_r50_lfb_nl.yaml'
    ocfg = H.oracle_cfg(yaml_name, TINY)
    params = OM.make_params(ocfg, seed=2, split='val')
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=3, crop=64, frames=8)
    p64 = dict((k, v.double()) for k, v in params.items())
    i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items())
    exact, _, _ = OM.forward(ocfg, p64, i64, 'val')
    i16 = dict(i64, lfb=inputs['lfb'].to(torch.bfloat16).double())
    rounded, _, _ = OM.forward(ocfg, p64, i16, 'val')
    H.setup_cfg(yaml_name, TINY)
    cfg.B200.LFB_DTYPE = 'bf16'
    workspace.ResetWorkspace()
    model, sfx = H.build('val', False)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    assert workspace.current().blobs['lfb' + sfx + '@bf16'].dtype == torch.bfloat16
    seen = []
    scan = fake.fbo_bank_scan
    fake.fbo_bank_scan = lambda bank, *a, **k: (seen.append(bank.dtype), scan(bank, *a, **k))[1]
    try:
        workspace.RunNet(model.net.Proto().name)
    finally:
        fake.fbo_bank_scan = scan
    assert seen == [torch.bfloat16, torch.bfloat16]
    for b in ['lfb_nl0_affinity_prob', 'lfb_nl1_sum', 'pred', 'prob']:
        got = workspace.FetchBlob('gpu_0/' + b).reshape(-1)
        assert H.rel(got, rounded[b].detach().numpy().reshape(-1)) < 1e-9, b
        assert H.rel(got, exact[b].detach().numpy().reshape(-1)) < 1e-2, b
    # switching the mode off drops the copy at the next feed
    cfg.B200.LFB_DTYPE = 'f32'
    H.feed_inputs(inputs, sfx)
    assert not workspace.HasBlob('gpu_0/lfb' + sfx + '@bf16')


def test_fbo_fold_is_skipped_for_bank_widths_the_scan_kernel_lacks(fake):
    """LFB.LFB_DIM = 512: vlfb_fbo_bank_scan only has 1024 / 2048 / 4096-float rows, so the test-mode graph must keep
    the as-written Conv / BatchMatMul lowering (and still match the oracle) instead of failing at run time."""
    from oracle import model as OM
    from vlfb import workspace
    from vlfb import executor as X
    ov = TINY + ['LFB.LFB_DIM', 512]
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', ov)
    params = OM.make_params(ocfg, seed=2, split='val')
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    assert inputs['lfb'].shape[-1] == 512
    blobs, _, _ = OM.forward(ocfg, dict((k, v.double()) for k, v in params.items()),
                             dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items()), 'val')
    H.setup_cfg('ava_r50_lfb_nl.yaml', ov)
    workspace.ResetWorkspace()
    model, sfx = H.build('val', False)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    net = workspace.current().nets[model.net.Proto().name]
    assert not any(isinstance(s, X.FboFoldStep) for s in net.steps)
    workspace.RunNet(model.net.Proto().name)
    assert H.rel(workspace.FetchBlob('gpu_0/pred').reshape(-1), blobs['pred'].detach().numpy().reshape(-1)) < 1e-9


def test_fetch_blob_async_equals_fetch_blob(fake):
    from vlfb import workspace
    _run('ava_r50_lfb_nl.yaml', TINY, fake, [])
    for name in ['loss', 'pred', 'pool5', 'conv1_w']:
        a, b = workspace.FetchBlob('gpu_0/' + name), workspace.FetchBlobAsync('gpu_0/' + name)
        assert b.ready() and np.array_equal(np.asarray(a), b.get()) and np.asarray(a).shape == b.get().shape, name


@pytest.mark.parametrize('stack', [True, False])
def test_training_nets_never_lower_to_the_inference_fold(fake, stack):
    """FboFoldStep has no backward (it is valid only without dropout between lfb_1x1 and phi / g): whatever the switches,
    a net with a loss must keep a differentiable lowering -- the FboNLStack operator or the as-written Conv / BatchMatMul
    graph -- and a forward-only TRAIN-split net (precise-BN's aux model, lfb_infer_only extraction aside) as well when
    dropout is on."""
    from core.config import config as cfg
    from vlfb import workspace
    from vlfb import executor as X
    ov = TINY + ['FBO_NL.LFB_DROPOUT_ON', True, 'FBO_NL.DROPOUT_RATE', 0.2]
    H.setup_cfg('ava_r50_lfb_nl.yaml', ov)
    cfg.B200.FBO_FOLD = True
    cfg.B200.FBO_STACK = stack
    workspace.ResetWorkspace()
    model, sfx = H.build('train', True)
    net = workspace.current().nets[model.net.Proto().name]
    assert net.train and net.losses
    assert not any(isinstance(s, X.FboFoldStep) for s in net.steps)
    assert sum(isinstance(s, X.FboStackStep) for s in net.steps) == (1 if stack else 0)
    assert any(isinstance(s, X.DropoutStep) for s in net.steps) or stack       # the stack applies the dropout itself
