"""CUDA-event timing of the stages of the per-IR frequency-response post-processing (bench.py's `fr_postprocess`):
frequency_response (K8 + K10), fractional-octave smoothing (K11), minimum-phase FIR (K12 + K13), per tail length.
usage: python tools/fr_post_phases.py [rows_per_tail]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200.frequency_response import (FrequencyResponse, equalize_batch, minimum_phase_batch,
                                                       smoothen_batch)
from impulcifer_pip313_b200.dispatcher import equalization_worker_batch
from impulcifer_pip313_b200.impulse_response import frequency_response_batch
from impulcifer_pip313_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 896
fs = 48000
grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=fs / 2)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


for tail in (84375, 96000, 57600, 21600):
    gen = torch.Generator(device='cuda').manual_seed(tail)
    x = torch.randn((rows, tail), generator=gen, device='cuda', dtype=torch.float64)
    x *= torch.exp(-torch.arange(tail, device='cuda', dtype=torch.float64) / (tail / 8.0))[None, :]
    t_full, _ = timed(lambda: ops.magnitude_db(x))
    t_fr, raw = timed(lambda: frequency_response_batch(x, fs))
    t_sm, sm = timed(lambda: smoothen_batch(grid, raw, 1 / 3, 1, 1 / 3, 1, 100.0, 10000.0))
    eq = -(sm - sm.mean(dim=1, keepdim=True))
    t_mp, _ = timed(lambda: minimum_phase_batch(grid, eq, fs=fs, f_res=5, normalize=False))
    print(f"tail {tail:6d} rows {rows}: magnitude_db (all bins) {t_full:.3f} ms | frequency_response {t_fr:.3f} ms | "
          f"smoothing {t_sm:.3f} ms | minimum phase {t_mp:.3f} ms", flush=True)


# the EQ-worker arithmetic (core/parallel_workers.py:98-131) on the last raw curves
err = raw - raw.mean(dim=1, keepdim=True)
t_hl, err_s = timed(lambda: smoothen_batch(grid, torch.maximum(smoothen_batch(grid, err, 1 / 6, 1, 1 / 3, 1, 100, 10000),
                                                                smoothen_batch(grid, err, 1 / 3, 1, 1.3, 1, 1000, 6000)),
                                           1 / 3, 1, 1 / 3, 1, 100, 10000))
t_eq, _ = timed(lambda: equalize_batch(grid, err_s, max_gain=40, treble_f_lower=10000, treble_f_upper=fs / 2))
t_all, _ = timed(lambda: equalization_worker_batch(grid, raw, fs, f_res=5))
print(f"EQ worker rows {rows}: smoothen_heavy_light {t_hl:.3f} ms | equalize {t_eq:.3f} ms | whole worker {t_all:.3f} ms", flush=True)
