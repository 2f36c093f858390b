import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from impulcifer_pip313_b200.frequency_response import FrequencyResponse, minimum_phase_batch
B=1792
fs=48000
grid = FrequencyResponse.generate_frequencies(f_step=1.01, f_min=10, f_max=fs / 2)
eq = torch.randn(B, len(grid), dtype=torch.float64, device='cuda').cumsum(1) * 0.05
for _ in range(3):
    fir = minimum_phase_batch(grid, eq, fs=fs, f_res=5, normalize=False)
torch.cuda.synchronize()
print(fir.shape)
