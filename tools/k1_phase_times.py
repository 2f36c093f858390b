"""Per-kernel time of the K1 launch triple via CUDA events around single launches (developer tool).
usage: [IMPB_LIB=variant.so] python tools/k1_phase_times.py [batch] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1792
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 64
est = ImpulseResponseEstimator(5.0, 48000)
L = 2 * est.fs + len(est)
x = torch.randn(batch, L, dtype=torch.float64, device='cuda') * 0.01
if prec == 32:
    x = x.float()
plan = est.plan(L, precision=prec)
out = torch.empty(batch, L, dtype=x.dtype, device='cuda')
for _ in range(3):
    plan.execute(x, out=out)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        plan.execute(x, out=out)
    torch.cuda.synchronize()
tot = {}
for e in prof.key_averages():
    if 'k_fx' in e.key:
        name = e.key.split('<')[0].replace('void impb::', '')
        tot[name] = tot.get(name, 0.0) + e.device_time_total / 3 / 1e3
print(os.environ.get('IMPB_LIB', 'default'), {k: round(v, 3) for k, v in tot.items()}, 'sum', round(sum(tot.values()), 3))
