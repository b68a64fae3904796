"""Minimal config reader for the oracle (test infrastructure).

Independent of the product's `core.config`: it only knows the keys the model
graph reads (reference lib/core/config.py:52-364 defaults) and merges a YAML on
top without type checking.
"""
import copy

import yaml


class ND(dict):
    """dict with attribute access (reference lib/utils/collections.py AttrDict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


# Defaults of the keys the graph reads; values from reference lib/core/config.py.
_DEFAULTS = {
    'DATASET': '',
    'NUM_GPUS': 8,
    'TRAIN': {'BATCH_SIZE': 64, 'CROP_SIZE': 224, 'VIDEO_LENGTH': 32,       # :102,:112,:131
              'DROPOUT_RATE': 0.0},                                          # :137
    'TEST': {'BATCH_SIZE': 64, 'CROP_SIZE': 256, 'VIDEO_LENGTH': 32},       # :196-207
    'MODEL': {'NUM_CLASSES': -1, 'VIDEO_ARC_CHOICE': 2, 'DEPTH': 50,         # :146-152
              'FC_INIT_STD': 0.01, 'USE_AFFINE': False, 'MULTI_LABEL': True,
              'DILATIONS_AFTER_CONV5': True, 'FREEZE_BACKBONE': False,
              'BN_EPSILON': 1.0000001e-5, 'BN_MOMENTUM': 0.9},                # :159-160
    'RESNETS': {'NUM_GROUPS': 1, 'WIDTH_PER_GROUP': 64},
    'NONLOCAL': {'CONV_INIT_STD': 0.01, 'NO_BIAS': 0, 'USE_MAXPOOL': True,   # :243-262
                 'USE_SOFTMAX': True, 'USE_ZERO_INIT_CONV': False, 'USE_BN': True,
                 'USE_SCALE': True, 'USE_AFFINE': False, 'LAYER_MOD': 2,
                 'CONV3_NONLOCAL': True, 'CONV4_NONLOCAL': True,
                 'BN_MOMENTUM': 0.9, 'BN_EPSILON': 1.0000001e-5},
    'AVA': {'LFB_MAX_NUM_FEAT_PER_STEP': 5},                                 # :311
    'ROI': {'SCALE_FACTOR': 16, 'XFORM_RESOLUTION': 7},                      # :339-340
    'LFB': {'ENABLED': False, 'LFB_DIM': 2048, 'WINDOW_SIZE': 100,           # :345-352
            'FBO_TYPE': 'nl'},
    'FBO_NL': {'NUM_LAYERS': 2, 'PRE_ACT': True, 'PRE_ACT_LN': True,         # :350-360
               'SCALE': True, 'LATENT_DIM': 512, 'INPUT_REDUCE_DIM': True,
               'DROPOUT_RATE': 0.2, 'INPUT_DROPOUT_ON': True,
               'LFB_DROPOUT_ON': True},
    'SOLVER': {'NESTEROV': True, 'WEIGHT_DECAY': 0.0001, 'WEIGHT_DECAY_BN': 0.0001,
               'MOMENTUM': 0.9, 'BASE_LR': 0.1},
}


def _to_nd(d):
    return ND({k: _to_nd(v) if isinstance(v, dict) else v for k, v in d.items()})


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v


def load(yaml_path=None, overrides=None):
    """Return an ND config: defaults <- yaml <- overrides (nested dict)."""
    cfg = copy.deepcopy(_DEFAULTS)
    if yaml_path is not None:
        with open(yaml_path, 'r') as f:
            _merge(cfg, yaml.safe_load(f))
    if overrides:
        _merge(cfg, overrides)
    return _to_nd(cfg)
