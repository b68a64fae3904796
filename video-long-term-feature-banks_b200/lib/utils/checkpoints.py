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
    out_dir = out_dir or create_and_get_checkpoint_directory()
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
    """<CHECKPOINT.DIR>/checkpoints (reference :233-238)."""
    if cfg.CHECKPOINT.DIR:
        return os.path.abspath(os.path.join(cfg.CHECKPOINT.DIR, 'checkpoints'))
    raise Exception('No cfg.CHECKPOINT.DIR specified.')


def create_and_get_checkpoint_directory():
    d = get_checkpoint_directory()
    os.makedirs(d, exist_ok=True)
    return d


def _checkpoint_iters():
    d = get_checkpoint_directory()
    if not os.path.isdir(d):
        return []
    its = []
    for f in os.listdir(d):
        if f.endswith('.pkl') and f.startswith('c2_model_iter'):
            its.append(int(f[len('c2_model_iter'):-len('.pkl')]))
    return sorted(its)


def find_checkpoint():
    """True when <dir>/c2_model_iter*.pkl exists (reference :73-82)."""
    return bool(_checkpoint_iters())


def get_checkpoint_resume_file():
    """The latest c2_model_iter{N}.pkl, or None (reference :50-70)."""
    its = _checkpoint_iters()
    return os.path.join(get_checkpoint_directory(), 'c2_model_iter{}.pkl'.format(its[-1])) if its else None


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
    loaded, missing = 0, []
    for name in wanted:
        if name not in blobs:
            logger.info('%s not found', name)
            if not name.endswith('_momentum') and 'pred' not in name:
                missing.append(name)
            continue
        base = name[:-len('_momentum')] if name.endswith('_momentum') else name
        if store.has(base):
            value = _fit_to_workspace(name, blobs[name], store.logical_shape(base))
            if value is None:
                continue
        else:
            value = np.asarray(blobs[name]).astype(np.float32, copy=False)     # the reference feeds it anyway
        workspace.FeedBlob(root + name, value)
        loaded += 1
    if missing:
        # a mis-built net (renamed layers) would otherwise "load" a checkpoint and silently train from scratch
        logger.warning('%d of %d model parameters are absent from %s (first: %s)', len(missing), len(wanted),
                       weights_file, ', '.join(missing[:5]))
    if wanted and loaded == 0:
        raise Exception('none of the model parameters was found in {}'.format(weights_file))
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


def resume_from(start_model_iter):
    """Iteration count of a run trained with another batch size (reference lib/utils/misc.py resume_from)."""
    assert cfg.TRAIN.RESUME_FROM_BATCH_SIZE > 0
    return int(start_model_iter * cfg.TRAIN.RESUME_FROM_BATCH_SIZE / cfg.TRAIN.BATCH_SIZE)


def load_model_from_params_file(model):
    """Resume / initialise as tools/train_net.py does (reference :180-230).  Returns start_model_iter (int).
      case 0  CHECKPOINT.CONVERT_MODEL: convert TRAIN.PARAMS_FILE (BN -> Affine, classifier dropped) first
      case 1  RESUME False, PARAMS_FILE set: load it (never its momentum)
      case 2  RESUME True,  PARAMS_FILE set: latest c2_model_iter*.pkl if one exists, else PARAMS_FILE
      case 3  RESUME True,  no PARAMS_FILE : latest checkpoint if one exists, else start from scratch (0)"""
    use_checkpoint = bool(cfg.CHECKPOINT.RESUME) and find_checkpoint()
    if cfg.TRAIN.PARAMS_FILE and cfg.CHECKPOINT.CONVERT_MODEL:
        assert cfg.MODEL.USE_AFFINE, 'a converted model uses affine layers'
        cfg.TRAIN.PARAMS_FILE = convert_model(cfg.TRAIN.PARAMS_FILE)
    if cfg.TRAIN.PARAMS_FILE and not use_checkpoint:
        start_model_iter, prev_lr = initialize_params_from_file(model=model, weights_file=cfg.TRAIN.PARAMS_FILE,
                                                                load_momentum=False)
        model.current_lr = prev_lr
        if cfg.TRAIN.RESUME_FROM_BATCH_SIZE > 0:
            start_model_iter = resume_from(start_model_iter)
        if cfg.TRAIN.RESET_START_ITER:
            start_model_iter = 0
    elif use_checkpoint:
        start_model_iter, prev_lr = initialize_params_from_file(model=model, weights_file=get_checkpoint_resume_file())
        model.current_lr = prev_lr
    else:
        start_model_iter = 0
        logger.info('No checkpoint found; training from scratch...')
    return int(start_model_iter)


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
