"""Where a MAPPO train_step's time goes (torch.profiler over one update)."""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO  # noqa: E402

M = 4
xyz = np.array([[float(i % 2), float(i // 2), 0.5] for i in range(M)])
env = BatchAviary(task="multihover", num_envs=65536, num_drones=M, initial_xyzs=xyz, seed=1, track_episode_stats=True)
algo = DeviceMAPPO(env, rollout_steps=32, mini_batch_size=32768, opt_epochs=1, rollout_values="zeros")
algo.train_step()
algo.collect_rollout()
algo.compute_returns()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    algo.update()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
