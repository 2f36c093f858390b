"""Latency of the drop-in class API on ONE recording set (what a user of the reference pipeline runs), developer tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200 import HRIR, ImpulseResponseEstimator, synth

est = ImpulseResponseEstimator(5.0, 48000)
rec = synth.recording_set(est.test_signal, 48000, synth.LAYOUT_71, seed=1234)


def chain():
    t = [time.perf_counter()]
    h = HRIR(est)
    h.ingest(rec, 48000, synth.LAYOUT_71)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    h.crop_heads()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    h.align_ipsilateral_all()
    h.align_onset_groups_peak_leftref()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    h.crop_tails()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    h.normalize()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    out = {s: {e: ir.data for e, ir in p.items()} for s, p in h.irs.items()}
    t.append(time.perf_counter())
    return np.diff(t) * 1e3


for _ in range(3):
    chain()
runs = np.array([chain() for _ in range(10)])
names = ['ingest (H2D + K1)', 'crop_heads', 'align x2', 'crop_tails', 'normalize', '.data (D2H)']
best = runs.min(axis=0)
print('one 7.1 set through the class API (ms, best of 10):', dict(zip(names, np.round(best, 2))), 'total', round(best.sum(), 2))
