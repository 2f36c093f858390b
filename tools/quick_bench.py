"""Developer timing loop for K1 (not the contract bench: see bench.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 224
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
fs = int(sys.argv[4]) if len(sys.argv) > 4 else 48000
est = ImpulseResponseEstimator(5.0, fs)
L = 2 * est.fs + len(est)
x = torch.randn(batch, L, dtype=torch.float64, device='cuda') * 0.01
if prec == 32:
    x = x.float()
plan = est.plan(L, precision=prec)
print('plan', plan.nfft, plan.n1, plan.n2, 'launches', plan.launches(batch), 'ws MB', plan.workspace(batch).numel() / 2**20)
out = torch.empty(batch, L, dtype=torch.float64 if prec == 64 else torch.float32, device='cuda')
track = sys.argv[5] if len(sys.argv) > 5 else ''
kw = {}
if 'a' in track:
    kw['absmax'] = torch.empty(batch, dtype=torch.float64, device='cuda')
if 'g' in track:
    kw['argmax'] = torch.empty(batch, dtype=torch.int64, device='cuda')
if 'b' in track and plan.blockmax_info()[0]:
    kw['blockmax'] = torch.empty((batch, plan.blockmax_info()[0]), dtype=torch.float64, device='cuda')
for _ in range(3):
    plan.execute(x, out=out, **kw)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for i in range(iters):
    plan.execute(x, out=out, **kw)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
t = min(ts)
bytes_alg = batch * L * (16 if prec == 64 else 8)
print(f'batch {batch} prec {prec}: {t:.3f} ms best, {np.mean(ts):.3f} mean -> {batch / t * 1e3:.0f} IR/s, '
      f'{bytes_alg / t / 1e6:.1f} GB/s algorithmic')
