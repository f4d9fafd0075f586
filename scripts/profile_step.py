"""Short driver for ncu: a few control steps of the headline workload (no timing here)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary, StepResult  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--drones", type=int, default=4)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--physics", default="dyn")
ap.add_argument("--task", default="multihover")
ap.add_argument("--ctrl-freq", type=int, default=30)
ap.add_argument("--many", type=int, default=0, help="also run bd_step_many over this many steps (one launch), twice")
args = ap.parse_args()
M = args.drones
side = int(np.ceil(np.sqrt(M)))
xyz = np.array([[float(i % side), float(i // side), 0.5] for i in range(M)])
env = BatchAviary(task=args.task, num_envs=args.envs, num_drones=M, initial_xyzs=xyz, pyb_freq=240,
                  ctrl_freq=args.ctrl_freq, act="rpm", precision="fp32", physics=args.physics, auto_reset=True,
                  seed=1)
slots = 4
acts = torch.rand((slots, args.envs, M, 4), device="cuda") * 2 - 1
obs = torch.empty((slots, args.envs, M, env.OBS_DIM), device="cuda")
env.reset_device(out=obs[0])
for k in range(args.steps):
    r = env.step_device(acts[k % slots])
if args.many > 0:
    K = args.many
    a = torch.rand((K, args.envs, M, 4), device="cuda") * 2 - 1
    o = torch.empty((K, args.envs, M, env.OBS_DIM), device="cuda")
    rw = torch.empty((K, args.envs), device="cuda")
    f = torch.empty((2, K, args.envs), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        env.step_many(a, o, rw, f[0], f[1])
torch.cuda.synchronize()
print("ok", float(r.reward.mean()))
