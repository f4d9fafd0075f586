"""Where BatchVecEnv.step spends its time relative to bd_step_host (same aviary, same pinned action buffer)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200 import BatchAviary  # noqa: E402
from marl_gym_pybullet_drones_b200.vec_env import BatchVecEnv  # noqa: E402

N, M = 65536, 4
xyz = np.array([[0, 0, .5], [1, 0, .5], [0, 1, .5], [1, 1, .5]], dtype=np.float64)
env = BatchAviary(task="multihover", num_envs=N, num_drones=M, initial_xyzs=xyz, seed=1, auto_reset=True,
                  reset_mode="jitter_philox")
env.reset_device()
venv = BatchVecEnv(env)
abuf = venv.action_buffer()
rng = np.random.default_rng(0)
abuf[...] = rng.uniform(-1, 1, abuf.shape).astype(np.float32)


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"step_host                : {timeit(lambda: env.step_host(abuf, actions_pinned=True)):.3f} ms")
done = []


def compact():
    r = env.step_host(abuf, actions_pinned=True, compact_terminal_obs=True)
    done.append(len(r['done_idx']))


print(f"step_host compact        : {timeit(compact):.3f} ms   finished envs per step: {np.mean(done[5:]):.0f}")
print(f"step_host full term. obs : {timeit(lambda: env.step_host(abuf, actions_pinned=True, want_terminal_obs=True)):.3f} ms")
per = []
for _ in range(30):
    t0 = time.perf_counter()
    venv.step(abuf)
    per.append((time.perf_counter() - t0) * 1e3)
print("BatchVecEnv.step, call by call (ms):", " ".join(f"{v:.2f}" for v in per))
print(f"BatchVecEnv.step         : {timeit(lambda: venv.step(abuf)):.3f} ms")
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    venv.step(abuf)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
