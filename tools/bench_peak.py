"""Developer timing of impb_peak_index on K1-shaped rows (not the contract bench)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 896
L = 391270
g = torch.Generator(device='cuda').manual_seed(5)
x = torch.randn(B, L, dtype=torch.float64, device='cuda', generator=g) * 1e-5
onset = torch.randint(100, 24000, (B,), device='cuda', generator=g)
t = torch.arange(L, device='cuda')[None, :] - onset[:, None]
env = torch.where(t >= 0, torch.exp(-t.clamp(min=0).double() / 9000.0), torch.zeros((), dtype=torch.float64, device='cuda'))
x += torch.randn(B, L, dtype=torch.float64, device='cuda', generator=g) * 0.05 * env
x[torch.arange(B, device='cuda'), onset] = 1.0
flat = x.reshape(-1)
offs = torch.arange(B, device='cuda', dtype=torch.int64) * L
lens = torch.full((B,), L, device='cuda', dtype=torch.int64)
known = x.abs().max(dim=1).values
for use_known in (True, False):
    for _ in range(3):
        pk = ops.peak_index(flat, offs, lens, max_len=L, known_max=known if use_known else None)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pk = ops.peak_index(flat, offs, lens, max_len=L, known_max=known if use_known else None)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    assert torch.equal(pk, onset)
    print(f'peak_index B={B} known_max={use_known}: {min(ts) * 1e3:.1f} us best, {np.mean(ts) * 1e3:.1f} mean')
