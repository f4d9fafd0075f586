"""CUDA-event time per control step of one configuration (default: BASELINE configs[4]'s per-GPU shape).

    python scripts/time_step_shape.py [--envs 131072 --drones 16 --physics dyn_dw --steps 50]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--drones", type=int, default=16)
ap.add_argument("--physics", default="dyn_dw")
ap.add_argument("--task", default="multihover")
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
import numpy as np  # noqa: E402
side = int(np.ceil(np.sqrt(args.drones)))     # drones of an env on a 1 m grid at 0.5 m (the bench's MAPPO block does the same)
xyz = np.array([[float(i % side), float(i // side), 0.5] for i in range(args.drones)])
env = BatchAviary(task=args.task, num_envs=args.envs, num_drones=args.drones, initial_xyzs=xyz, physics=args.physics, pyb_freq=240,
                  ctrl_freq=30, act="rpm", precision="fp32", device="cuda:0", auto_reset=True, seed=1)
env.reset_device()
acts = [torch.rand(args.envs, args.drones, 4, device="cuda") * 2 - 1 for _ in range(4)]
for i in range(10):
    env.step_device(acts[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.steps):
    env.step_device(acts[i % 4])
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / args.steps * 1e3
drones = args.envs * args.drones
D = env.obs_dim if hasattr(env, "obs_dim") else 72
print(f"{args.task} M={args.drones} {args.physics} {args.envs} envs: {us:.1f} us per control step, "
      f"{drones * 8 / us * 1e6:.3e} drone-substeps/s")
