"""One-process-per-GPU data parallelism: NCCL all-reduce (sum) of the flat gradient buffers
and the initial parameter broadcast.  Replaces Caffe2's per-parameter NCCLAllreduce ops
 and checkpoints.broadcast_parameters (lib/utils/checkpoints.py:386-407).

torch.distributed is plumbing only (NCCL communicator + stream handling); on CPU test
runs the same code path uses the gloo backend.
"""
import os

import torch
import torch.distributed as dist

BUCKET_ELEMS = 8 * 1024 * 1024        # 32 MB fp32 buckets: large enough to saturate NVLink 5


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment.  Returns (rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            local = int(os.environ.get('LOCAL_RANK', '0'))
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend=backend)
    return rank, world


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_grads(store):
    """Sum the gradient ranges across ranks, bucket by bucket (async, then one wait)."""
    if world_size() == 1:
        return 0
    handles = []
    for flat in store.grad_ranges():
        for s in range(0, flat.numel(), BUCKET_ELEMS):
            handles.append(dist.all_reduce(flat[s:s + BUCKET_ELEMS], op=dist.ReduceOp.SUM, async_op=True))
    for h in handles:
        h.wait()
    return len(handles)


def broadcast_params(store, src=0):
    """Rank `src`'s parameters (and momentum, when allocated) become everyone's."""
    if world_size() == 1:
        return
    for ch in store.chunks:
        dist.broadcast(ch['P'], src)
        dist.broadcast(ch['Pt'], src)
        if ch['Mo'] is not None:
            dist.broadcast(ch['Mo'], src)


def install(ws):
    """Hook the gradient all-reduce into workspace.RunNet for training nets."""
    ws.allreduce = allreduce_grads if world_size() > 1 else None
