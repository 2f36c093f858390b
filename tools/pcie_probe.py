"""PCIe ceiling of the box (developer tool): pinned H2D / D2H alone and concurrently."""
import time
import torch

n_in, n_out = 1451463680, 646464000
hi = torch.empty(n_in, dtype=torch.uint8).pin_memory()
ho = torch.empty(n_out, dtype=torch.uint8).pin_memory()
di = torch.empty(n_in, dtype=torch.uint8, device='cuda')
do = torch.empty(n_out, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                di.copy_(hi, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


t = run(True, False)
print(f'H2D alone   {n_in / t / 1e9:.1f} GB/s')
t = run(False, True)
print(f'D2H alone   {n_out / t / 1e9:.1f} GB/s')
t = run(True, True)
print(f'both        H2D {n_in / t / 1e9:.1f} GB/s + D2H {n_out / t / 1e9:.1f} GB/s in {t * 1e3:.1f} ms -> '
      f'{896 / t:.0f} IR/s ceiling for the e2e step')
