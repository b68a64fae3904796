"""Learning-rate schedule (host arithmetic; reference lib/utils/lr_policy.py:41-157)."""
import numpy as np

from core.config import config as cfg


def get_step_index(cur_iter):
    """Which LR step `cur_iter` falls in (reference :138-147)."""
    assert cfg.SOLVER.STEPS[0] == 0, 'The first step should always start at 0.'
    steps = list(cfg.SOLVER.STEPS) + [cfg.SOLVER.MAX_ITER]
    ind = 0
    for ind, step in enumerate(steps):
        if cur_iter < step:
            break
    return ind - 1


def _base_lr(it):
    policy = cfg.SOLVER.LR_POLICY
    if policy == 'steps_with_lrs':
        return cfg.SOLVER.LRS[get_step_index(it)]
    if policy == 'steps_with_relative_lrs':
        return cfg.SOLVER.LRS[get_step_index(it)] * cfg.SOLVER.BASE_LR
    if policy == 'steps_with_decay':
        return cfg.SOLVER.BASE_LR * cfg.SOLVER.GAMMA ** get_step_index(it)
    if policy == 'step':
        return cfg.SOLVER.BASE_LR * cfg.SOLVER.GAMMA ** (it // cfg.SOLVER.STEP_SIZE)
    raise NotImplementedError('Unknown LR policy: {}'.format(policy))


def get_lr_at_iter(it):
    """LR at iteration `it` incl. the linear warm-up (reference :41-65)."""
    lr = np.float32(_base_lr(it))
    last_it = cfg.SOLVER.WARMUP.WARMUP_END_ITER
    if cfg.SOLVER.WARMUP.WARMUP_ON and it < last_it:
        lr_start = np.float32(cfg.SOLVER.WARMUP.WARMUP_START_LR)
        lr_end = np.float32(_base_lr(last_it))
        lr = it * (lr_end - lr_start) / (last_it - 1) + lr_start
    return np.float32(lr)
