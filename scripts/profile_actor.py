import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.actor import FusedActor
from marl_gym_pybullet_drones_b200.mappo import MLP
rows = 262144
mlp = MLP(72, 4, [256, 256], "tanh").cuda()
fa = FusedActor(72, 256, 4)
fa.set_weights(mlp, torch.full((4,), -0.5, device="cuda"))
obs = torch.randn(rows, 72, device="cuda")
for _ in range(6):
    a, l = fa.forward(obs)
torch.cuda.synchronize()
print("ok", float(a.mean()))
