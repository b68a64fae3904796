"""Small helpers that are part of the builder boundary (reference lib/utils/misc.py:68-79)."""
from core.config import config as cfg


def get_batch_size(split):
    """Per-GPU batch size (reference lib/utils/misc.py:68-72)."""
    if split in ['test', 'val']:
        return int(cfg.TEST.BATCH_SIZE / cfg.NUM_GPUS)
    elif split == 'train':
        return int(cfg.TRAIN.BATCH_SIZE / cfg.NUM_GPUS)


def get_crop_size(split):
    """reference lib/utils/misc.py:75-79."""
    if split in ['test', 'val']:
        return cfg.TEST.CROP_SIZE
    elif split == 'train':
        return cfg.TRAIN.CROP_SIZE
