"""Pipeline timeline of the fused actor kernel (CTA 0): SM-clock stamps per tile and phase."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.actor import FusedActor  # noqa: E402
from marl_gym_pybullet_drones_b200.mappo import MLP  # noqa: E402

rows = 262144
mlp = MLP(72, 4, [256, 256], "tanh").cuda()
fa = FusedActor(72, 256, 4)
fa.set_weights(mlp, torch.full((4,), -0.5, device="cuda"))
obs = torch.randn(rows, 72, device="cuda")
for _ in range(3):
    fa.forward(obs)
tiles = (rows // 128 + 147) // 148
buf = torch.zeros((tiles, 16), dtype=torch.int64, device="cuda")
fa._lib.bd_actor_set_trace(fa._h, C.c_void_p(buf.data_ptr()))
fa.forward(obs)
torch.cuda.synchronize()
fa._lib.bd_actor_set_trace(fa._h, None)
t = buf.cpu().numpy().astype(np.float64)
t0 = t[0, 0]
names = ["E wait L1", "E L1 ready", "E epi1 done", "E L2 ready", "E epi2 done", "A L3 ready", "A staged X(t+2)", "A rows written",
         "M staged+W1", "M L1(t+1) issued", "M L2 issued", "M L3 issued", "Ld A issue", "Ld A here+free", "Ld B issue", "Ld B here"]
ghz = 1.965
print("stamps in us relative to the first one (CTA 0, thread 0 = E, MMA thread = M, first auxiliary thread = A)")
for k in range(min(tiles, 6)):
    order = [i for i in np.argsort(t[k, :16]) if t[k, i] != 0]
    print(f"tile {k}: " + "  ".join(f"{names[i]}={(t[k, i] - t0) / ghz / 1e3:.2f}" for i in order))
d = np.diff(t[:, 1]) / ghz / 1e3
print("tile period (us):", np.round(d, 2))
