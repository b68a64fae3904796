"""Where does the end-to-end loop of bench.py lose its ~1.3 ms per step over the device-resident step?  Times (CUDA events /
wall clock): the pinned host->device copies of one batch alone, the dequeue conversions alone, the captured step alone, and
the e2e loop with the copies (a) as in bench.py, (b) skipped (inputs stay resident, RunNet + loss read-back only)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'video-long-term-feature-banks_b200', 'lib'), ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import harness as H
from core.config import config as cfg
from oracle import model as OM
from vlfb import workspace


def main():
    ov = ['NUM_GPUS', 1, 'TRAIN.BATCH_SIZE', 2]
    H.setup_cfg('ava_r50_lfb_nl.yaml', ov)
    cfg.RNG_SEED = 2
    workspace.ResetWorkspace()
    model, sfx = H.build('train', True)
    ocfg = H.oracle_cfg('ava_r50_lfb_nl.yaml', ov)
    H.feed_params(OM.make_params(ocfg, seed=2))
    workspace.FeedBlob('gpu_0/lr', np.array(1e-4, dtype=np.float32))
    inputs = OM.make_inputs(ocfg, n_clips=2, rois_per_clip=2, seed=100)
    host = dict((k, v.contiguous().pin_memory()) for k, v in inputs.items())
    name = model.net.Proto().name
    feed = dict(('gpu_0/%s%s' % (k, sfx), v) for k, v in host.items())
    for k, v in feed.items():
        workspace.FeedBlob(k, v)
    for _ in range(4):
        workspace.RunNet(name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

    nbytes = sum(v.numel() * v.element_size() for v in host.values())
    dev = dict((k, torch.empty_like(v, device='cuda')) for k, v in host.items())

    def copies():
        for k, v in host.items():
            dev[k].copy_(v, non_blocking=True)
    ms, wall = timed(copies)
    print('H2D copies of one batch (%.1f MB, pinned): %.3f ms device, %.3f ms wall = %.1f GB/s' % (nbytes / 1e6, ms, wall, nbytes / 1e6 / ms))
    print('captured step alone: %.3f ms device, %.3f ms wall' % timed(lambda: workspace.RunNet(name)))

    def loop_full():
        workspace.RunNet(name)
        workspace.EnqueueBlobs(feed)
        float(workspace.FetchBlob('gpu_0/loss'))
    workspace.EnqueueBlobs(feed)
    print('e2e loop (RunNet, EnqueueBlobs, blocking FetchBlob): %.3f ms device, %.3f ms wall' % timed(loop_full))
    workspace.RunNet(name)

    def loop_nofetch():
        workspace.RunNet(name)
        workspace.EnqueueBlobs(feed)
    workspace.EnqueueBlobs(feed)
    print('e2e loop without the loss read-back: %.3f ms device, %.3f ms wall' % timed(loop_nofetch))
    workspace.RunNet(name)

    def loop_nocopy():
        workspace.RunNet(name)
        float(workspace.FetchBlob('gpu_0/loss'))
    print('RunNet + blocking FetchBlob, inputs resident: %.3f ms device, %.3f ms wall' % timed(loop_nocopy))

    def dequeue_only():
        workspace.EnqueueBlobs(feed)
        workspace._dequeue_blobs()
    print('EnqueueBlobs + dequeue (copies + conversions, no step): %.3f ms device, %.3f ms wall' % timed(dequeue_only))

    # does sampling the clocks while the loop runs perturb it?  (bench.py reported 14.1-14.2 ms for the loop measured at
    # 12.9 ms above)
    import subprocess
    import threading
    workspace.EnqueueBlobs(feed)
    q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap'
    for period in (100, 500):
        proc = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', str(period)],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        time.sleep(0.5)
        print('e2e loop with `nvidia-smi -lms %d` running: %.3f ms device, %.3f ms wall' % ((period,) + timed(loop_full, 20)))
        proc.terminate()
        proc.wait()
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        stop = []
        got = []

        def poll():
            while not stop:
                got.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
                time.sleep(0.05)
        th = threading.Thread(target=poll, daemon=True)
        th.start()
        time.sleep(0.2)
        print('e2e loop with an NVML polling thread (50 ms): %.3f ms device, %.3f ms wall' % timed(loop_full, 20), len(got), got[-1])
        stop.append(1)
        th.join()
    except Exception as exc:
        print('pynvml unavailable: %r' % (exc,))
    print('e2e loop, nothing sampling: %.3f ms device, %.3f ms wall' % timed(loop_full, 20))
    workspace.RunNet(name)


if __name__ == '__main__':
    main()
