"""ncu target: K8 (magnitude_db) and the minimum-phase chain alone on cropped-IR-shaped rows (developer tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from impulcifer_pip313_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 448
n = int(sys.argv[2]) if len(sys.argv) > 2 else 96000
x = torch.randn(B, n, dtype=torch.float64, device='cuda')
for _ in range(3):
    m = ops.magnitude_db(x)
torch.cuda.synchronize()
print(m.shape)
