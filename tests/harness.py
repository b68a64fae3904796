"""Shared test harness: build a (tiny or full) model through the reference-style builder API,
load oracle parameters, run it through vlfb.workspace and compare with the oracle."""
import os

import numpy as np
import torch
import torch

from core import config as C
from core.config import config as cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG_DIR = os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'configs')


def setup_cfg(yaml_name, overrides):
    C.reset_cfg()
    C.cfg_from_file(os.path.join(CFG_DIR, yaml_name))
    C.cfg_from_list(overrides)
    C.assert_and_infer_cfg()


def oracle_cfg(yaml_name, overrides):
    """The oracle's own view of the same configuration."""
    from oracle import refcfg
    o = {}
    for k, v in zip(overrides[0::2], overrides[1::2]):
        d = o
        parts = k.split('.')
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return refcfg.load(os.path.join(CFG_DIR, yaml_name), o)


def build(split='train', train=True, suffix=None, lfb_infer_only=False):
    from models import model_builder_video
    from vlfb import workspace
    suffix = ('_train' if split == 'train' else '_test') if suffix is None else suffix
    model = model_builder_video.ModelBuilder(train=train, use_cudnn=True, cudnn_exhaustive_search=True,
                                             ws_nbytes_limit=256 * 1024 * 1024, split=split,
                                             name='{}_net'.format(split))
    model.build_model(suffix=suffix, lfb=None, lfb_infer_only=lfb_infer_only)
    workspace.RunNetOnce(model.param_init_net)
    workspace.CreateNet(model.net)
    return model, suffix


def feed_params(params):
    from vlfb import workspace
    for name, t in params.items():
        workspace.FeedBlob('gpu_0/' + name, t.detach().double().numpy())


def feed_inputs(inputs, suffix):
    from vlfb import workspace
    for name, t in inputs.items():
        workspace.FeedBlob('gpu_0/{}{}'.format(name, suffix), t.double().numpy() if t.dtype == torch.float32 else t.numpy())


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
