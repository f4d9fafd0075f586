"""The README's round-2 snippet (open-loop tape, VecEnv protocol), executed."""
import functools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_gym_pybullet_drones_b200 import BatchAviary, MultiHoverAviary, make_vec_envs  # noqa: E402

acts = torch.rand(64, 8192, 4, 4, device="cuda") * 2 - 1
obs = torch.empty(64, 8192, 4, 72, device="cuda"); rew = torch.empty(64, 8192, device="cuda")
term = torch.empty(64, 8192, dtype=torch.bool, device="cuda"); trunc = torch.empty_like(term)
small = BatchAviary(task="multihover", num_envs=8192, num_drones=4)
small.reset_device(); small.step_many(acts, obs, rew, term, trunc)
torch.cuda.synchronize()
print("step_many ok", float(rew.mean()), int(term.sum()))

venv = make_vec_envs(functools.partial(MultiHoverAviary, num_drones=4), batch_size=65536)
a = venv.action_buffer(); a[...] = -1.0
o, info = venv.reset()
for _ in range(40):
    o, r, done, info = venv.step(a)
print("vecenv ok", o.shape, r.dtype, int(done.sum()), type(info["n"][int(done.argmax())].get("terminal_observation")))
