"""Developer timing of K9 (window_resample) for the two shapes of the named configs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200 import ops
from impulcifer_pip313_b200.impulse_response import resample_taps


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), r


for rows, n, fs_new, fs_old in ((14, 230400, 48000, 192000), (14, 57600, 44100, 48000), (896, 96000, 44100, 48000)):
    up, down, taps = resample_taps(fs_new, fs_old)
    x = torch.randn(rows, n, dtype=torch.float64, device='cuda')
    ms, y = timed(lambda: ops.window_resample(x, taps, up, down))
    macs = y.shape[1] * rows * (len(taps) / up)
    print(f'{rows} x {n} {fs_old}->{fs_new} (up {up}, down {down}): {ms:.3f} ms, out {tuple(y.shape)}, '
          f'{macs / ms / 1e6:.1f} GMAC/s')
