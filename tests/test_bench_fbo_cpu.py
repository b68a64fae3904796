"""bench_fbo.py's FBO-only net (BASELINE.json configs[4]) builds and runs through the engine on the CPU stand-in
kernels: folded and as-written inference agree, the train-mode net produces a loss and gradients."""
import numpy as np
import pytest


@pytest.fixture
def fake():
    import fake_kernels
    from vlfb import workspace
    fake_kernels.install()
    workspace.ResetWorkspace()
    yield fake_kernels
    workspace.ResetWorkspace()
    fake_kernels.uninstall()


def test_fbo_only_net_modes_agree(fake):
    import bench_fbo as B
    from vlfb import workspace
    from vlfb import executor as X
    out = {}
    for mode in ('infer_fold', 'infer'):
        model, name = B.build_case(mode, 3, 20, 3)
        net = workspace.current().nets[name]
        assert sum(isinstance(s, X.FboFoldStep) for s in net.steps) == (3 if mode == 'infer_fold' else 0)
        workspace.RunNet(name)
        out[mode] = workspace.FetchBlob('gpu_0/prob').copy()
    assert np.abs(out['infer_fold'] - out['infer']).max() < 1e-10
    assert out['infer'].std() > 0
    # the bf16-bank mode of the microbenchmark: same net, the scan reads the bf16 copy made when the bank was fed
    model, name = B.build_case('infer_fold_bf16', 3, 20, 3)
    import torch
    assert workspace.current().blobs['lfb_test@bf16'].dtype == torch.bfloat16
    workspace.RunNet(name)
    p16 = workspace.FetchBlob('gpu_0/prob')
    assert 0 < np.abs(p16 - out['infer_fold']).max() < 2e-2
    assert B.fbo_bytes(4, 300, 3, bank_s=2) == B.fbo_bytes(4, 300, 3) - 2 * 4 * 300 * 2048
    model, name = B.build_case('train', 3, 20, 2)
    workspace.RunNet(name)
    assert np.isfinite(workspace.FetchBlob('gpu_0/loss'))
    assert np.abs(workspace.FetchBlob('gpu_0/lfb_1x1_w_grad')).max() > 0
    # the byte / FLOP formulas of SURVEY 8(d): fp32, R=4, L=300, 3 layers -> 9.87 MB + 20.99 MB; 1.58 GFLOP per RoI
    assert abs(B.fbo_bytes(4, 300, 3) - 30.9e6) < 0.1e6
    assert abs(B.fbo_flops(1, 300, 3) - 1.58e9) < 0.01e9
