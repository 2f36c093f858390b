"""Developer timing of BatchDispatcher.postprocess stages (not the contract bench)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200 import ops
from impulcifer_pip313_b200.frequency_response import FrequencyResponse, minimum_phase_batch, smoothen_batch
from impulcifer_pip313_b200.impulse_response import frequency_response_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 896
n = int(sys.argv[2]) if len(sys.argv) > 2 else 96000
fs = 48000
x = torch.randn(B, n, dtype=torch.float64, device='cuda') * torch.exp(-torch.arange(n, device='cuda') / 4000.0)[None, :]
grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=fs / 2)


def timed(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append((a.elapsed_time(b), (time.perf_counter() - t0) * 1e3))
    print(f'{name:28s} device {min(t[0] for t in ts):8.3f} ms   wall {min(t[1] for t in ts):8.3f} ms')
    return r


mag = timed('magnitude_db', lambda: ops.magnitude_db(x))
raw = timed('frequency_response_batch', lambda: frequency_response_batch(x, fs))
sm = timed('smoothen_batch', lambda: smoothen_batch(grid, raw, 1 / 3, 1, 1 / 3, 1, 100.0, 10000.0))
eq = -(sm - sm.mean(dim=1, keepdim=True))
fir = timed('minimum_phase_batch f_res=5', lambda: minimum_phase_batch(grid, eq, fs=fs, f_res=5, normalize=False))
fir = timed('minimum_phase_batch f_res=10', lambda: minimum_phase_batch(grid, eq, fs=fs, f_res=10, normalize=False))
