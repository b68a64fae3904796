"""SpatialBN (trainable batch norm) and precise-BN (lib/utils/bn_helper.py).

CPU part: oracle/ops.spatial_bn pinned against torch.nn.functional.batch_norm (an external implementation of the same
published operator), the engine's lowering of a `MODEL.USE_AFFINE False` / `NONLOCAL.USE_BN True` net against the
oracle in fp64 (forward blobs, batch / running statistics, every gradient, test-mode forward), and the precise-BN
helper against a direct numpy computation of the population statistics.
GPU part (-m gpu): csrc/bn.cu against the oracle op on the production shapes; a tiny training step on the tcgen05 path.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import harness as H
from oracle import ops as OPS

TINY = ['NUM_GPUS', 1, 'TRAIN.BATCH_SIZE', 2, 'TRAIN.CROP_SIZE', 64, 'TRAIN.VIDEO_LENGTH', 8,
        'TEST.BATCH_SIZE', 2, 'TEST.CROP_SIZE', 64, 'TEST.VIDEO_LENGTH', 8, 'LFB.WINDOW_SIZE', 4,
        'TRAIN.DROPOUT_RATE', 0.0, 'FBO_NL.INPUT_DROPOUT_ON', False, 'FBO_NL.LFB_DROPOUT_ON', False]
BN = TINY + ['MODEL.USE_AFFINE', False, 'NONLOCAL.USE_AFFINE', False, 'NONLOCAL.USE_BN', True,
             'TRAIN.ITER_COMPUTE_PRECISE_BN', 3,
             # the reference's Conv3dBN drops `dilations` (model_builder_video.py:176-185), so its SpatialBN nets only
             # build without the dilated res5 (the Kinetics pre-training setting)
             'MODEL.DILATIONS_AFTER_CONV5', False]


@pytest.fixture
def fake():
    import fake_kernels
    from vlfb import workspace
    fake_kernels.install()
    workspace.ResetWorkspace()
    yield fake_kernels
    workspace.ResetWorkspace()
    fake_kernels.uninstall()


def test_oracle_spatial_bn_matches_torch_batch_norm():
    g = torch.Generator().manual_seed(0)
    x = (torch.randn((3, 8, 4, 5, 6), generator=g, dtype=torch.float64) * 2 + 0.7).requires_grad_(True)
    s = (torch.rand(8, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    b = torch.randn(8, generator=g, dtype=torch.float64).requires_grad_(True)
    rm, rv = torch.randn(8, generator=g, dtype=torch.float64), torch.rand(8, generator=g, dtype=torch.float64) + 0.5
    w = torch.randn(x.shape, generator=g, dtype=torch.float64)
    y, nrm, nrv, sm, siv = OPS.spatial_bn(x, s, b, rm, rv, eps=1e-5, momentum=0.9, is_test=False)
    (y * w).sum().backward()
    got = [t.grad.clone() for t in (x, s, b)]
    for t in (x, s, b):
        t.grad = None
    trm, trv = rm.clone(), rv.clone()
    ty = F.batch_norm(x, trm, trv, s, b, training=True, momentum=0.1, eps=1e-5)     # torch momentum = 1 - Caffe2 momentum
    (ty * w).sum().backward()
    assert (y - ty).abs().max() < 1e-12
    assert (nrm - trm).abs().max() < 1e-12 and (nrv - trv).abs().max() < 1e-12
    for a, t in zip(got, (x, s, b)):
        assert (a - t.grad).abs().max() < 1e-10
    xd = x.detach()
    assert (sm - xd.mean(dim=(0, 2, 3, 4))).abs().max() < 1e-12
    assert (siv - 1 / torch.sqrt(xd.var(dim=(0, 2, 3, 4), unbiased=False) + 1e-5)).abs().max() < 1e-12
    yt = OPS.spatial_bn(xd, s, b, rm, rv, eps=1e-5, is_test=True)[0]
    assert (yt - F.batch_norm(xd, rm, rv, s, b, training=False, eps=1e-5)).abs().max() < 1e-12


BLOBS = ['res_conv1_bn', 'pool1', 'res2_0_branch2c_bn', 'res2_2_branch2c_bn', 'nonlocal_conv3_1_bn', 'nonlocal_conv3_1_sum',
         'res5_2_branch2c_bn', 'box_pooled', 'pred', 'prob']
STATS = ['res_conv1_bn', 'res2_0_branch1_bn', 'res3_1_branch2b_bn', 'nonlocal_conv4_1_bn', 'res5_2_branch2c_bn']


def test_bn_net_matches_oracle_on_the_cpu_engine(fake):
    from oracle import model as OM
    from vlfb import workspace
    H.setup_cfg('ava_r50_lfb_nl.yaml', BN)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', BN)
    params = OM.make_params(ocfg, seed=2)
    assert 'res2_0_branch2a_bn_riv' in params and 'nonlocal_conv3_1_bn_rm' in params
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    model, sfx = H.build('train', True)
    assert sorted(model.GetComputedParams()) == sorted(k for k in params if k.endswith('_bn_rm') or k.endswith('_bn_riv'))
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    p64 = dict((k, v.double().requires_grad_(True)) for k, v in params.items())
    i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items())
    blobs, prob, loss = OM.forward(ocfg, p64, i64, 'train')
    loss.backward()
    net = workspace.current().nets[model.net.Proto().name]
    upd, net.update_ops = net.update_ops, []           # forward + backward, no SGD: compare gradients
    workspace.RunNet(model.net.Proto().name)
    net.update_ops = upd
    assert H.rel(workspace.FetchBlob('gpu_0/loss'), loss.item()) < 1e-9
    for b in BLOBS:
        assert H.rel(workspace.FetchBlob('gpu_0/' + b), blobs[b].detach().numpy()) < 1e-9, b
    for layer in STATS:                  # batch statistics (precise-BN inputs) and the momentum-updated running statistics
        for sfx_ in ('_sm', '_siv', '_rm', '_riv'):
            assert H.rel(workspace.FetchBlob('gpu_0/' + layer + sfx_), blobs[layer + sfx_].numpy()) < 1e-9, layer + sfx_
    trainable = model.TrainableParams()
    # BN scale / bias ARE trained (with SOLVER.WEIGHT_DECAY_BN); the running statistics are not
    assert sorted(trainable) == sorted(k for k in params if not (k.endswith('_bn_rm') or k.endswith('_bn_riv')))
    for name in trainable:
        g = workspace.FetchBlob('gpu_0/' + name + '_grad')
        ref = p64[name].grad.numpy()
        e = float(np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-5))
        assert e < 1e-7, (name, e)
    # a full step moves the BN parameters with the '_bn' weight-decay rule (model_builder_video.py:373)
    before = workspace.FetchBlob('gpu_0/res3_0_branch2a_bn_s').copy()
    workspace.RunNet(model.net.Proto().name)
    assert np.abs(workspace.FetchBlob('gpu_0/res3_0_branch2a_bn_s') - before).max() > 0


def test_bn_test_net_uses_running_statistics(fake):
    from oracle import model as OM
    from vlfb import workspace
    H.setup_cfg('ava_r50_lfb_nl.yaml', BN)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', BN)
    params = OM.make_params(ocfg, seed=2, split='val')
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    model, sfx = H.build('val', False)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    p64 = dict((k, v.double()) for k, v in params.items())
    i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items())
    blobs, prob, _ = OM.forward(ocfg, p64, i64, 'val')
    workspace.RunNet(model.net.Proto().name)
    for b in ['res_conv1_bn', 'res3_3_branch2c_bn', 'nonlocal_conv4_1_sum', 'box_pooled', 'pred']:
        assert H.rel(workspace.FetchBlob('gpu_0/' + b), blobs[b].detach().numpy()) < 1e-9, b
    assert not workspace.HasBlob('gpu_0/res_conv1_bn_sm')             # only training-mode BN emits sm / siv
    assert np.array_equal(workspace.FetchBlob('gpu_0/res_conv1_bn_rm'), params['res_conv1_bn_rm'].numpy().astype(np.float64))


def test_precise_bn_helper_installs_population_statistics(fake):
    """bn_helper.BatchNormHelper over 3 batches == mean / variance of the concatenated activations of those batches
    (with the parameters frozen, the aux net is forward only)."""
    from core.config import config as cfg
    from oracle import model as OM
    from utils import bn_helper
    from vlfb import workspace
    H.setup_cfg('ava_r50_lfb_nl.yaml', BN)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', BN)
    params = OM.make_params(ocfg, seed=2)
    model, sfx = H.build('train', True)
    H.feed_params(params)
    batches = [OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8, seed=10 + i) for i in range(3)]
    helper = bn_helper.BatchNormHelper()
    helper.create_bn_aux_model(node_id=0, suffix=sfx)
    assert 'res_conv1' in helper._bn_layers and 'nonlocal_conv3_1' in helper._bn_layers
    assert len(helper._bn_layers) == len([k for k in params if k.endswith('_bn_riv')])
    w_before = workspace.FetchBlob('gpu_0/res2_0_branch2a_w').copy()
    helper.compute_and_update_bn_stats(curr_iter=7, feed_fn=lambda i: H.feed_inputs(batches[i], sfx))
    assert np.array_equal(workspace.FetchBlob('gpu_0/res2_0_branch2a_w'), w_before)        # no update ran
    # expected: E[x], E[x^2] averaged over the batches, from the oracle's conv outputs feeding each BN
    p64 = dict((k, v.double()) for k, v in params.items())
    ex, ex2 = {}, {}
    for b in batches:
        i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in b.items())
        blobs, _, _ = OM.forward(ocfg, p64, i64, 'train')
        for layer in ('res_conv1', 'res4_2_branch2b', 'nonlocal_conv4_1'):
            m = blobs[layer + '_bn_sm'].numpy()
            v = (1. / blobs[layer + '_bn_siv'].numpy()) ** 2 - cfg.MODEL.BN_EPSILON
            ex[layer] = ex.get(layer, 0) + m / 3
            ex2[layer] = ex2.get(layer, 0) + (v + m ** 2) / 3
    for layer in ex:
        # the helper feeds np.float32 arrays, as the reference does (bn_helper.py:206-220): fp32 rounding of the values
        assert H.rel(workspace.FetchBlob('gpu_0/%s_bn_rm' % layer), ex[layer]) < 2e-7, layer
        assert H.rel(workspace.FetchBlob('gpu_0/%s_bn_riv' % layer), ex2[layer] - ex[layer] ** 2) < 2e-7, layer
    # same iteration again: cached statistics are re-installed without running the net (bn_helper.py:130-136)
    workspace.FeedBlob('gpu_0/res_conv1_bn_rm', np.zeros(64, dtype=np.float32))
    helper.compute_and_update_bn_stats(curr_iter=7, feed_fn=lambda i: 1 / 0)
    assert H.rel(workspace.FetchBlob('gpu_0/res_conv1_bn_rm'), ex['res_conv1']) < 2e-7


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize('rows,C', [(2 * 16 * 56 * 56, 64), (2 * 16 * 14 * 14, 1024), (37, 8), (1, 4), (6272, 2048)])
def test_spatial_bn_kernels_match_oracle_op(rows, C):
    from vlfb import kernels as K
    g = torch.Generator().manual_seed(rows + C)
    x = torch.randn((rows, C), generator=g) * 1.7 + torch.randn((C,), generator=g) * 3.0     # |mean| up to ~3 sigma
    s, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    dy = torch.randn((rows, C), generator=g)
    x64 = x.double().t().reshape(1, C, rows).requires_grad_(True)
    s64, b64 = s.double().requires_grad_(True), b.double().requires_grad_(True)
    y, nrm, nrv, sm, siv = OPS.spatial_bn(x64, s64, b64, rm.double(), rv.double(), 1e-5, 0.9, False)
    (y * dy.double().t().reshape(1, C, rows)).sum().backward()
    d = 'cuda'
    xg, sg, bg, rmg, rvg = x.to(d), s.to(d), b.to(d), rm.to(d), rv.to(d)
    smg, sivg, yg = torch.empty(C, device=d), torch.empty(C, device=d), torch.empty((rows, C), device=d)
    K.spatial_bn_fwd(xg, sg, bg, rmg, rvg, smg, sivg, yg, 1e-5, 0.9)

    def close(a, ref, tol):
        ref = ref.detach()
        return float((a.double().cpu() - ref).abs().max() / max(float(ref.abs().max()), 1e-6)) < tol
    assert close(yg.t().reshape(1, C, rows), y, 3e-6)
    assert close(smg, sm, 1e-6) and close(sivg, siv, 1e-5)
    assert close(rmg, nrm, 1e-6) and close(rvg, nrv, 1e-5)
    if rows > 1:
        dx, ds, db = torch.empty((rows, C), device=d), torch.full((C,), 0.5, device=d), torch.full((C,), -0.25, device=d)
        K.spatial_bn_bwd(dy.to(d), xg, sg, smg, sivg, dx, ds, db)
        assert close(dx.t().reshape(1, C, rows), x64.grad, 2e-5)
        assert close(ds - 0.5, s64.grad, 2e-5) and close(db + 0.25, b64.grad, 2e-5)         # accumulated into ds / db
    yt = torch.empty((rows, C), device=d)
    K.spatial_bn_infer(xg, sg, bg, rm.to(d), rv.to(d), yt, 1e-5)
    ref = OPS.spatial_bn(x64.detach(), s.double(), b.double(), rm.double(), rv.double(), 1e-5, is_test=True)[0]
    assert close(yt.t().reshape(1, C, rows), ref, 3e-6)


@pytest.mark.gpu
def test_tiny_bn_train_step_on_gpu():
    """MODEL.USE_AFFINE False: conv (tcgen05) -> SpatialBN -> ReLU ...; forward blobs, batch statistics and gradients
    against the fp64 oracle with the tolerances of the Affine path's tiny-model test."""
    from oracle import model as OM
    from vlfb import kernels, workspace
    kernels.set_gemm_backend('tcgen05')
    workspace.ResetWorkspace()
    H.setup_cfg('ava_r50_lfb_nl.yaml', BN)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', BN)
    params = OM.make_params(ocfg, seed=2)
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, crop=64, frames=8)
    model, sfx = H.build('train', True)
    H.feed_params(params)
    H.feed_inputs(inputs, sfx)
    p64 = dict((k, v.double().requires_grad_(True)) for k, v in params.items())
    i64 = dict((k, (v.double() if v.dtype == torch.float32 else v)) for k, v in inputs.items())
    blobs, prob, loss = OM.forward(ocfg, p64, i64, 'train')
    loss.backward()
    # second oracle: operands rounded to TF32 at the graph points where the engine rounds them.
    # Tolerances: a batch-normalised random net AMPLIFIES perturbations where the Affine net of the other tests damps
    # them -- measured on the fp64 oracle of exactly this model, a 1e-6 relative perturbation of the clip arrives at
    # res5_2_branch2c_bn as 1.5e-5 (x15) with SpatialBN and as 1.6e-7 (x0.16) with AffineNd, a factor ~100 -- so the
    # TF32-level noise every layer injects (rounding-boundary flips, fp32 accumulation order), which the Affine net keeps
    # below 1e-3 at the last blobs, reaches ~1e-2 of their range here (call K: 2.9e-4 after conv1, 5.4e-4 after res2,
    # 1.3e-3 after NL3, 8.3e-3 at res5).  The bounds below follow that growth; the kernels themselves are held to 3e-6 /
    # 2e-5 by test_spatial_bn_kernels_match_oracle_op and the lowering to 1e-9 by the fp64 CPU-engine test above.
    e64 = dict((k, v.double()) for k, v in params.items())
    eblobs, _, eloss = OM.forward(ocfg, e64, i64, 'train', emulate_tf32=True)
    net = workspace.current().nets[model.net.Proto().name]
    upd, net.update_ops = net.update_ops, []
    workspace.RunNet(model.net.Proto().name)
    net.update_ops = upd
    torch.cuda.synchronize()
    names = ['res_conv1_bn', 'res2_2_branch2c_bn', 'nonlocal_conv3_1_sum', 'res5_2_branch2c_bn', 'box_pooled', 'pred']
    rep = dict((b, (H.rel(workspace.FetchBlob('gpu_0/' + b), eblobs[b].detach().numpy()),
                    H.rel(workspace.FetchBlob('gpu_0/' + b), blobs[b].detach().numpy()))) for b in names)
    print('\nBN tiny model, rel err vs (tf32-emulating oracle, fp64 oracle):', ' '.join('%s=%.1e/%.1e' % (k, a, b) for k, (a, b) in rep.items()))
    assert H.rel(workspace.FetchBlob('gpu_0/loss'), eloss.item()) < 5e-3
    assert H.rel(workspace.FetchBlob('gpu_0/loss'), loss.item()) < 5e-3
    bound = {'res_conv1_bn': 1e-3, 'res2_2_branch2c_bn': 2e-3, 'nonlocal_conv3_1_sum': 5e-3}
    for b, (e_emul, e_exact) in rep.items():
        assert e_emul < bound.get(b, 3e-2), (b, e_emul)
        assert e_exact < 2 * bound.get(b, 3e-2), (b, e_exact)
    for layer in STATS:
        tol = 2e-3 if layer in ('res_conv1_bn', 'res2_0_branch1_bn') else 3e-2
        for sfx_ in ('_sm', '_siv', '_rm', '_riv'):
            assert H.rel(workspace.FetchBlob('gpu_0/' + layer + sfx_), eblobs[layer + sfx_].numpy()) < tol, layer + sfx_
    cos = []
    for name in model.TrainableParams():
        g = workspace.FetchBlob('gpu_0/' + name + '_grad').astype(np.float64).ravel()
        ref = p64[name].grad.numpy().ravel()
        if np.abs(ref).max() < 1e-12:
            continue
        cos.append(float(g @ ref / (np.linalg.norm(g) * np.linalg.norm(ref) + 1e-300)))
    print('gradient cosine vs fp64 oracle: min %.4f median %.5f' % (min(cos), float(np.median(cos))))
    # call L: min 0.965, median 0.982 (the forward noise above, once more through the batch-norm backward)
    assert min(cos) > 0.9 and float(np.median(cos)) > 0.95, (min(cos), float(np.median(cos)))
    workspace.ResetWorkspace()
