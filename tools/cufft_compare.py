"""Yardstick, not product code: the same batched 'same' convolution written with the cuFFT library through
torch.fft (rfft -> x H -> irfft -> slice), timed next to K1 on the same device and batch.

Usage: python tools/cufft_compare.py [batch] [iters]
Prints one line per variant; `profiles/` keeps the output of the last run.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from impulcifer_pip313_b200.estimator import ImpulseResponseEstimator

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 224
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
est = ImpulseResponseEstimator(5.0, 48000)
M = len(est)
L = 2 * est.fs + M
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(3)
x = torch.randn(batch, L, dtype=torch.float64, device=dev, generator=g) * 0.01
inv = torch.from_numpy(np.ascontiguousarray(est.inverse_filter)).to(dev)
start = (M - 1) // 2


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


plan = est.plan(L, precision=64)
out = torch.empty(batch, L, dtype=torch.float64, device=dev)
t_k1 = timed(lambda: plan.execute(x, out=out))
print(f'K1 (this library, N={plan.nfft}): {t_k1:.3f} ms for {batch} IRs -> {batch / t_k1 * 1e3:.0f} IR/s')

ref = None
for n in (589824, 540000, 691200):
    # n >= L + ceil((M-1)/2) is enough for the slice that is kept (wrap-around lands outside it);
    # 691200 = scipy's next_fast_len(L + M - 1), the length the reference itself transforms at
    H = torch.fft.rfft(inv, n=n)
    chunk = max(1, min(batch, (1 << 30) // (n * 8)))  # keep cuFFT's scratch + spectra around 1 GiB like K1

    def run():
        res = torch.empty(batch, L, dtype=torch.float64, device=dev)
        for lo in range(0, batch, chunk):
            X = torch.fft.rfft(x[lo:lo + chunk], n=n)
            X *= H
            y = torch.fft.irfft(X, n=n)
            res[lo:lo + chunk] = y[:, start:start + L] if n >= start + L else torch.cat(
                [y[:, start:], y[:, :start + L - n]], dim=1)
        return res

    t = timed(run)
    y = run()
    err = float((y - out).abs().max() / out.abs().max())
    print(f'cuFFT rfft/irfft n={n} (chunk {chunk}): {t:.3f} ms -> {batch / t * 1e3:.0f} IR/s; '
          f'K1 speed-up {t / t_k1:.2f}x; max rel diff vs K1 {err:.2e}')
    del H, y
