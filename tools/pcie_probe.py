"""Host <-> device copy ceiling of the box for the e2e step's copy pattern (developer tool).

    python tools/pcie_probe.py                                                    # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py   # N GPUs at once

Every rank copies what one bench step moves per GPU (1.45 GB of PCM_32 recordings in, 0.65 GB of cropped responses
out; page-locked buffers) alone and concurrently; with N ranks the copies of all GPUs run at the same time (barrier
before each timed repetition), so the aggregate figures are the host-side ceiling the N-GPU `e2e` number can reach.
"""
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n_in, n_out = 1451463680, 646464000
hi = torch.empty(n_in, dtype=torch.uint8).pin_memory()
ho = torch.empty(n_out, dtype=torch.uint8).pin_memory()
di = torch.empty(n_in, dtype=torch.uint8, device='cuda')
do = torch.empty(n_out, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        sync_all()
        t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                di.copy_(hi, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


t = run(True, False)
if rank == 0:
    print(f'{world} GPU(s) at once: H2D alone   {world * n_in / t / 1e9:.1f} GB/s aggregate ({n_in / t / 1e9:.1f} per GPU)')
t = run(False, True)
if rank == 0:
    print(f'{world} GPU(s) at once: D2H alone   {world * n_out / t / 1e9:.1f} GB/s aggregate ({n_out / t / 1e9:.1f} per GPU)')
t = run(True, True)
if rank == 0:
    print(f'{world} GPU(s) at once: both        H2D {world * n_in / t / 1e9:.1f} + D2H {world * n_out / t / 1e9:.1f} GB/s aggregate '
          f'in {t * 1e3:.1f} ms -> {world * 896 / t:.0f} IR/s ceiling for the e2e step '
          f'(64 sets per GPU per step; the bench moves 128 sets = twice these bytes per step)')
if world > 1:
    dist.destroy_process_group()
