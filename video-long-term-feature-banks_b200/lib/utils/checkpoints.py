"""Weight interchange with the reference's pickle checkpoints (reference
lib/utils/checkpoints.py:88-177 conversion, :271-383 loading, :386-407 broadcast, :421-459 saving).

File format (unchanged): pickle of {'blobs': {unscoped_name: ndarray, 'lr': float, 'model_iter': int, ...}} or the
bare blobs dict; python-2 pickles are read with encoding='latin1'.  Names are the reference's blob names, shapes
its NCTHW conv layout (Cout, Cin, kT, kH, kW) -- vlfb.workspace.FeedBlob re-lays weights out for the kernels
([Cout][kT][kH][kW][Cin], stem padded to 8 px x 4 ch) and refreshes the TF32 operand copy, so nothing here touches
device layouts.  What this module restates is the reference's loading RULES: BN -> Affine fold (eps 1e-5), removal of
momentum / bookkeeping fields, 2D -> 3D inflation (stack kT copies / kT), `pred*` reshaped or skipped when the class
count differs, `lr` / `model_iter` carried over, momentum restored for training nets.
"""
import logging
import os
import pickle
from collections import OrderedDict

import numpy as np

from core.config import config as cfg
from vlfb import workspace

logger = logging.getLogger(__name__)


def _unscope(name):
    return str(name).split('/')[-1]


def _read_pickle(path):
    with open(path, 'rb') as f:
        return pickle.load(f, encoding='latin1')


# ---- conversion of a Caffe2 classification checkpoint (trainable BN) into the Affine form ---------------------------
def remove_spatial_bn_layers(c2cls_weights):
    """Fold every '<layer>_bn_{s,b,rm,riv}' quadruple into scale/bias: s' = s/sqrt(var+1e-5), b' = b - mean*s'
    (reference :88-116; riv holds the running VARIANCE)."""
    blobs = c2cls_weights['blobs']
    done = set()
    for name in sorted(blobs.keys()):
        cut = name.find('_bn_')
        if cut < 0 or name[:cut] in done:
            continue
        layer = name[:cut]
        done.add(layer)
        if layer + '_bn_rm' not in blobs:
            continue                                   # already an affine layer
        std = np.sqrt(blobs[layer + '_bn_riv'] + 1e-5)
        scale = blobs[layer + '_bn_s'] / std
        blobs[layer + '_bn_b'] = blobs[layer + '_bn_b'] - blobs[layer + '_bn_rm'] * scale
        blobs[layer + '_bn_s'] = scale
        del blobs[layer + '_bn_rm'], blobs[layer + '_bn_riv']


def remove_non_param_fields(c2cls_weights):
    for field in ('epoch', 'model_iter', 'lr'):
        c2cls_weights['blobs'].pop(field, None)


def remove_momentum(c2cls_weights):
    for k in [k for k in c2cls_weights['blobs'] if k.endswith('_momentum')]:
        del c2cls_weights['blobs'][k]


def load_and_convert_caffe2_cls_model(model_file_name):
    weights = _read_pickle(model_file_name)
    if 'blobs' not in weights:
        weights = {'blobs': weights}
    remove_non_param_fields(weights)
    remove_momentum(weights)
    remove_spatial_bn_layers(weights)
    return weights


def convert_model(model_path, out_dir=None):
    """Kinetics-pretrained classification model -> initialisation file: classifier and momentum dropped,
    lr = 0.00125 recorded (reference :145-177).  Returns the path of 'converted_model.pkl'."""
    out_dir = out_dir or get_checkpoint_directory()
    blobs = load_and_convert_caffe2_cls_model(model_path)['blobs']
    for name in list(blobs.keys()):
        if 'pred' in name or 'momentum' in name:
            del blobs[name]
    blobs['lr'] = 0.00125
    out = os.path.join(out_dir, 'converted_model.pkl')
    with open(out, 'wb') as f:
        pickle.dump(blobs, f, 2)
    return out


def get_checkpoint_directory():
    d = cfg.CHECKPOINT.DIR if 'CHECKPOINT' in cfg and cfg.CHECKPOINT.get('DIR') else '.'
    os.makedirs(d, exist_ok=True)
    return d


# ---- loading --------------------------------------------------------------------------------------------------------
def _fit_to_workspace(name, value, ws_shape):
    """The reference's per-blob rules (:316-372).  Returns the array to feed or None (= skip)."""
    value = np.asarray(value)
    ws_shape = tuple(int(s) for s in ws_shape)
    if 'pred' in name:
        if int(np.prod(ws_shape)) != int(np.prod(value.shape)):
            logger.info('%s (classifier) found but unmatching (not loaded): %s ---> %s', name, value.shape, ws_shape)
            return None
        value = value.reshape(ws_shape)
    if len(ws_shape) != value.ndim:
        assert ws_shape[:2] == value.shape[:2] and ws_shape[-2:] == value.shape[-2:], \
            'Workspace blob {} with shape {} does not match weights file shape {}'.format(name, ws_shape, value.shape)
        kt = ws_shape[2]
        logger.info('%s inflated %s ---> %s', name, value.shape, ws_shape)
        value = np.stack([value] * kt, axis=2) / float(kt)
    assert ws_shape == tuple(value.shape), \
        'Workspace blob {} with shape {} does not match weights file shape {}'.format(name, ws_shape, value.shape)
    return value.astype(np.float32, copy=False)


def initialize_master_gpu_model_params(model, weights_file, load_momentum=True):
    """Feed every parameter (and, for training nets, momentum) found in `weights_file`; returns (model_iter, lr)."""
    blobs = _read_pickle(weights_file)
    if 'blobs' in blobs:
        blobs = blobs['blobs']
    model_iter = blobs.get('model_iter', 0)
    if 'lr' in blobs:
        prev_lr = float(blobs['lr'])
    elif cfg.TRAIN.RESET_START_ITER:
        prev_lr = 1.
    else:
        raise Exception('No lr blob found.')
    wanted = OrderedDict()
    if 'test' not in model.net.Name() and load_momentum:
        trainable = set(model.TrainableParams())
        for p in model.params:
            if p in trainable:
                wanted[_unscope(p) + '_momentum'] = True
    for p in model.GetAllParams():
        wanted[_unscope(p)] = True
    store = workspace.current().params
    root = 'gpu_{}/'.format(cfg.ROOT_GPU_ID)
    for name in wanted:
        if name not in blobs:
            logger.info('%s not found', name)
            continue
        base = name[:-len('_momentum')] if name.endswith('_momentum') else name
        if not store.has(base):
            continue
        value = _fit_to_workspace(name, blobs[name], store.logical_shape(base))
        if value is not None:
            workspace.FeedBlob(root + name, value)
    workspace.FeedBlob(root + 'lr', np.array(prev_lr, dtype=np.float32))
    return model_iter, prev_lr


def broadcast_parameters(model):
    """One process per GPU: rank 0's parameters (and momentum) go to every rank over NCCL (reference :386-407 copies
    them between the in-process replicas)."""
    from vlfb import dist as vdist
    if vdist.world_size() > 1:
        vdist.broadcast_params(workspace.current().params)


def initialize_params_from_file(model, weights_file, load_momentum=True):
    model_iter, prev_lr = initialize_master_gpu_model_params(model, weights_file, load_momentum)
    broadcast_parameters(model)
    return model_iter, prev_lr


def load_model_from_params_file_for_test(model, weights_file):
    initialize_params_from_file(model=model, weights_file=weights_file)


def load_model_from_params_file(model):
    """Resume / initialise as tools/train_net.py does (reference :180-214): returns (start_model_iter, prev_lr)."""
    params_file = cfg.TRAIN.PARAMS_FILE if cfg.TRAIN.get('PARAMS_FILE') else ''
    if not params_file:
        return 0, None
    model_iter, prev_lr = initialize_params_from_file(model, params_file, load_momentum=not cfg.TRAIN.RESET_START_ITER)
    return (0 if cfg.TRAIN.RESET_START_ITER else model_iter), prev_lr


# ---- saving ---------------------------------------------------------------------------------------------------------
def save_model_params(model, params_file, model_iter):
    """{'blobs': {model_iter, lr, <param>_momentum..., <param>...}} with reference names and NCTHW shapes (:421-459)."""
    root = 'gpu_{}/'.format(cfg.ROOT_GPU_ID)
    out = OrderedDict()
    out['model_iter'] = model_iter + 1
    out['lr'] = workspace.FetchBlob(root + 'lr')
    trainable = set(model.TrainableParams())
    for p in model.GetParams():
        if p in trainable:
            out.setdefault(_unscope(p) + '_momentum', workspace.FetchBlob(root + _unscope(p) + '_momentum'))
    for p in model.GetAllParams():
        out.setdefault(_unscope(p), workspace.FetchBlob(root + _unscope(p)))
    with open(params_file, 'wb') as f:
        pickle.dump(dict(blobs=out), f, 2)
