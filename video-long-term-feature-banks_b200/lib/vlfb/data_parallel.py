"""Data-parallel wiring: stands in for caffe2.python.data_parallel_model.Parallelize_GPU
(called at reference lib/models/model_builder_video.py:147-157).

The reference replicates the net over NUM_GPUS devices inside ONE process and chains one
NCCLAllreduce per parameter.  Here ONE process drives ONE GPU (torchrun), so the builder
functions run once; gradients of all trainable parameters live in a flat buffer that
vlfb.dist all-reduces (sum -- the loss is pre-scaled by 1/NUM_GPUS, resnet_video.py:333).
"""
from . import executor


def Parallelize_GPU(model, input_builder_fun, forward_pass_builder_fun, param_update_builder_fun=None,
                    devices=None, rendezvous=None, broadcast_computed_params=False,
                    optimize_gradient_memory=False, use_nccl=True, **kwargs):
    devices = list(devices) if devices is not None else [0]
    input_builder_fun(model)
    losses = forward_pass_builder_fun(model, 1.0 / max(len(devices), 1))
    model._losses = [str(l) for l in (losses or []) if l is not None]
    model._devices = devices
    if param_update_builder_fun is not None and model._losses:
        model._want_grads = True
        # "AddGradientOperators": static analysis that fills model.param_to_grad
        executor.CompiledNet(model, None)
        param_update_builder_fun(model)
    return model
