"""Where does BatchDispatcher.process spend host time?  (developer tool)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200 import BatchDispatcher, ImpulseResponseEstimator, synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
est = ImpulseResponseEstimator(5.0, 48000)
base = synth.recording_set(est.test_signal, 48000, synth.LAYOUT_71, seed=1234)
batch = torch.from_numpy(base).cuda()[None].repeat(S, 1, 1).contiguous()
disp = BatchDispatcher(est, synth.LAYOUT_71)
for _ in range(3):
    disp.process(batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    disp.process(batch)
torch.cuda.synchronize()
print('wall per step ms', (time.perf_counter() - t0) / 5 * 1e3)
disp.phase_timing = True
acc, hacc = {}, {}
for _ in range(5):
    disp.process(batch)
    for name, ms in disp.phase_ms:
        acc[name] = acc.get(name, 0.0) + ms / 5
    for name, ms in disp.phase_host_ms:
        hacc[name] = hacc.get(name, 0.0) + ms / 5
disp.phase_timing = False
print('device time per phase (ms):', {k: round(v, 3) for k, v in acc.items()}, 'sum', round(sum(acc.values()), 3))
print('host time per phase (ms):', {k: round(v, 3) for k, v in hacc.items()}, 'sum', round(sum(hacc.values()), 3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    disp.process(batch)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
