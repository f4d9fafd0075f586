"""Where the host-buffer step (bd_step_host) spends its time: H2D of the actions, kernel, D2H of the results."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200 import BatchAviary  # noqa: E402

N, M = 65536, 4
xyz = np.array([[0, 0, .5], [1, 0, .5], [0, 1, .5], [1, 1, .5]], dtype=np.float64)
env = BatchAviary(task="multihover", num_envs=N, num_drones=M, initial_xyzs=xyz, seed=1)
env.reset_device()
act_h = torch.empty((N, M, 4), pin_memory=True).uniform_(-1, 1)
obs_h = torch.empty((N, M, 72), pin_memory=True)
act_d = torch.empty((N, M, 4), device="cuda")
obs_d = torch.empty((N, M, 72), device="cuda")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_h2d = timeit(lambda: act_d.copy_(act_h, non_blocking=True))
t_d2h = timeit(lambda: obs_h.copy_(obs_d, non_blocking=True))
t_k = timeit(lambda: env.step_device(act_d))
print(f"H2D {act_h.numel() * 4 / 1e6:.1f} MB: {t_h2d * 1e3:.0f} us ({act_h.numel() * 4 / t_h2d / 1e6:.1f} GB/s)")
print(f"D2H {obs_h.numel() * 4 / 1e6:.1f} MB: {t_d2h * 1e3:.0f} us ({obs_h.numel() * 4 / t_d2h / 1e6:.1f} GB/s)")
print(f"kernel: {t_k * 1e3:.0f} us")
# both directions at once (full duplex?)
s2 = torch.cuda.Stream()


def both():
    with torch.cuda.stream(s2):
        act_d.copy_(act_h, non_blocking=True)
    obs_h.copy_(obs_d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


print(f"H2D || D2H: {timeit(both) * 1e3:.0f} us")
# D2H in 8 chunks on the same stream (copy-engine launch overhead per chunk)
chunks = obs_d.view(8, -1), obs_h.view(8, -1)
print(f"D2H in 8 chunks: {timeit(lambda: [chunks[1][i].copy_(chunks[0][i], non_blocking=True) for i in range(8)]) * 1e3:.0f} us")
import time
a_np, = (act_h.numpy(),)
t0 = time.perf_counter()
for _ in range(20):
    env.step_host(a_np)
dt = (time.perf_counter() - t0) / 20
print(f"step_host: {dt * 1e6:.0f} us per step")
