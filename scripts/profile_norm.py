"""Short driver for ncu: a few running-normaliser updates on the headline observation batch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.normalization import MeanStdNormalizer  # noqa: E402

xs = [torch.randn(65536, 4, 72, device="cuda") for _ in range(4)]
n = MeanStdNormalizer(shape=(4, 72), device="cuda")
y = torch.empty_like(xs[0])
for i in range(6):
    n.update(xs[i % 4])
    n.rms.normalize(xs[i % 4], 10.0, out=y)
torch.cuda.synchronize()
print("ok", float(n.rms.count))
