"""Kernel-time table of ONE DeviceMAPPO rollout at the configs[4] per-GPU shape (torch.profiler, CUDA activities)."""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO  # noqa: E402

N, M, T = int(os.environ.get("ENVS", 131072)), int(os.environ.get("DRONES", 16)), 32
side = int(np.ceil(np.sqrt(M)))
xyz = np.array([[float(i % side) - 0.5 * (side - 1), float(i // side) - 0.5 * (side - 1), 0.5] for i in range(M)])
env = BatchAviary(task="multihover", num_envs=N, num_drones=M, initial_xyzs=xyz, physics=os.environ.get("PHYSICS", "dyn_dw"), seed=1,
                  track_episode_stats=True)
algo = DeviceMAPPO(env, rollout_steps=T, mini_batch_size=32768, opt_epochs=1)
algo.collect_rollout()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
algo.collect_rollout()
e1.record()
torch.cuda.synchronize()
print(f"rollout: {e0.elapsed_time(e1):.1f} ms for {T} steps = {e0.elapsed_time(e1) / T * 1e3:.0f} us per step; norm_obs={algo.norm_obs}")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    algo.collect_rollout()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
