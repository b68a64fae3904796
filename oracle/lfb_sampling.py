"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's long-term-feature-bank construction and
window sampling -- the host code that runs immediately before the FBO.  Parity unpinned (the
reference ships no fixtures); each function follows the cited reference lines statement by statement, including the
ORDER of the np.random calls, so that with the same numpy seed the product's index tables select the same rows.

Only tests/ may import this module.
"""
import numpy as np

FPS = 24          # lib/datasets/charades.py:40, epic.py:41


def construct_ava_lfb(all_features, all_metadata):
    """tools/lfb_loader.py:82-113: {video_id: {sec: [feature, ...]}} in arrival order."""
    lfb = {}
    for iter_features, iter_metadata in zip(all_features, all_metadata):
        for gpu_features, gpu_metadata in zip(iter_features, iter_metadata):
            assert gpu_features.shape[0] == gpu_metadata.shape[0]
            for i in range(gpu_features.shape[0]):
                video_id, sec, _, _ = gpu_metadata[i].tolist()
                video_id = int(np.round(video_id))
                sec = int(np.round(sec))
                lfb.setdefault(video_id, {}).setdefault(sec, []).append(np.squeeze(gpu_features[i]))
    return lfb


def construct_frame_level_lfb(all_features, all_metadata):
    """tools/lfb_loader.py:51-79 (charades flavour: metadata = (video_id, frame_id))."""
    lfb = {}
    global_idx = 0
    for iter_features in all_features:
        for gpu_features in iter_features:
            for i in range(gpu_features.shape[0]):
                if global_idx >= len(all_metadata):
                    break
                video_id, frame_id = all_metadata[global_idx][-2:] if len(all_metadata[global_idx]) == 2 \
                    else (all_metadata[global_idx][1], all_metadata[global_idx][2])
                global_idx += 1
                lfb.setdefault(video_id, {})[frame_id] = np.squeeze(gpu_features[i])
    return lfb


def sample_lfb_ava(in_video_lfb, sec, window_size, K, lfb_dim):
    """lib/datasets/ava.py:300-323."""
    lower = sec - (window_size // 2)
    video_lfb = np.zeros((window_size * K, lfb_dim))
    for j, si in enumerate(range(lower, lower + window_size)):
        if si in in_video_lfb:
            num_feat = len(in_video_lfb[si])
            num_feat_used = min(num_feat, K)
            random_lfb_indices = np.random.choice(range(num_feat), num_feat_used, replace=False)
            for k, rand_idx in enumerate(random_lfb_indices):
                video_lfb[j * K + k] = in_video_lfb[si][rand_idx]
    return video_lfb


def sample_lfb_charades(video_lfb, center_idx, window_size, clips_per_second, lfb_dim):
    """lib/datasets/charades.py:251-276."""
    secs = window_size // clips_per_second
    begin = int(np.round(center_idx - (float(secs) / 2.0 * FPS)))
    end = begin + secs * FPS
    out_lfb = []
    for frame_idx in range(begin, end + 1):
        if frame_idx in video_lfb:
            if len(out_lfb) < window_size:
                out_lfb.append(video_lfb[frame_idx])
    out = np.zeros((window_size, lfb_dim))
    if len(out_lfb) > 0:
        out[:len(out_lfb)] = np.array(out_lfb)
    return out


def sample_verb_lfb_epic(center_idx, video_lfb, window_size, lfb_dim):
    """lib/datasets/epic.py:310-331."""
    half_len = (window_size * FPS) // 2
    lower, upper = center_idx - half_len, center_idx + half_len
    out_lfb = []
    for frame_idx in range(lower, upper + 1):
        if frame_idx in video_lfb.keys():
            if len(out_lfb) < window_size:
                out_lfb.append(video_lfb[frame_idx])
    out_lfb = np.array(out_lfb)
    if out_lfb.shape[0] < window_size:
        new = np.zeros((window_size, lfb_dim))
        if out_lfb.shape[0] > 0:
            new[:out_lfb.shape[0]] = out_lfb
        out_lfb = new
    return out_lfb.astype(np.float32)


def sample_noun_lfb_epic(center_idx, video_lfb, window_size, max_num_feat_per_frame, frames_per_second, lfb_dim):
    """lib/datasets/epic.py:338-374 (several detections per frame; the first `max_num_feat_per_frame` of each frame
    are taken in frame order until `window_size` rows are collected)."""
    secs = float(window_size) / (max_num_feat_per_frame * frames_per_second)
    lower = int(center_idx - (secs / 2) * FPS)
    upper = int(lower + secs * FPS)
    out_lfb = []
    num_feat = 0
    for frame_idx in range(lower, upper + 1):
        if frame_idx in video_lfb:
            frame_lfb = video_lfb[frame_idx]
            if not (isinstance(frame_lfb, list) and len(frame_lfb) == 0):
                curr_num = min(max_num_feat_per_frame, frame_lfb.shape[0])
                num_feat += curr_num
                out_lfb.append(frame_lfb[:curr_num])
                if num_feat >= window_size:
                    break
    if len(out_lfb) == 0:
        return np.zeros((window_size, lfb_dim))
    out_lfb = np.vstack(out_lfb)[:window_size].astype(np.float32)
    if out_lfb.shape[0] < window_size:
        new = np.zeros((window_size, lfb_dim))
        new[:out_lfb.shape[0]] = out_lfb
        out_lfb = new
    return out_lfb
