"""End-to-end on-device MAPPO timing: env steps/s of the rollout (actor forward + env step) and of a
whole train_step (rollout + GAE + PPO update)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--drones", type=int, default=4)
ap.add_argument("--physics", default="dyn")
ap.add_argument("--rollout", type=int, default=32)
ap.add_argument("--fused", type=int, default=1)
ap.add_argument("--mb", type=int, default=32768)
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--precision", default="tf32")
args = ap.parse_args()
M = args.drones
side = int(np.ceil(np.sqrt(M)))
xyz = np.array([[float(i % side) - 0.5 * (side - 1), float(i // side) - 0.5 * (side - 1), 0.5] for i in range(M)])   # 1 m grid centred on the origin (MultiHover's |x|,|y| <= 3 m box)
env = BatchAviary(task="multihover", num_envs=args.envs, num_drones=M, initial_xyzs=xyz, physics=args.physics, seed=1,
                  track_episode_stats=True)
algo = DeviceMAPPO(env, rollout_steps=args.rollout, mini_batch_size=args.mb, opt_epochs=args.epochs,
                   fused_actor=bool(args.fused), rollout_values="zeros", matmul_precision=args.precision)
algo.collect_rollout()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    algo.collect_rollout()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
S = 8
print(f"rollout  : {dt / args.rollout * 1e6:8.1f} us per env step  -> {args.envs * M * S * args.rollout / dt:.3e} drone-substeps/s "
      f"({args.envs * args.rollout / dt:.3e} env-steps/s)  fused={bool(algo.fused)}")
algo.train_step()      # includes the one-off CUDA-graph capture of the PPO minibatch
torch.cuda.synchronize()
t0 = time.perf_counter()
res = algo.train_step()
dt = time.perf_counter() - t0
print(f"train_step: {dt:.3f} s for {args.rollout * args.envs} env-steps -> {args.rollout * args.envs / dt:.3e} env-steps/s  {res}")
for name, fn in (("collect_rollout", algo.collect_rollout), ("compute_returns", algo.compute_returns), ("update", algo.update)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print(f"  {name:16s} {1e3 * (time.perf_counter() - t0):8.1f} ms")
