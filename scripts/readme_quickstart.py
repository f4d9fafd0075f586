"""The README's quick-start snippet, run as it is (smoke check of the public surface)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO, FlockAviary, MultiHoverAviary  # noqa: E402

env = BatchAviary(task="multihover", num_envs=65536, num_drones=4, track_episode_stats=True,
                  initial_xyzs=[[0, 0, .5], [1, 0, .5], [0, 1, .5], [1, 1, .5]])
obs = env.reset_device()
res = env.step_device(torch.zeros(65536, 4, 4, device="cuda"))
print(obs.shape, res.reward.mean().item(), res.terminated.sum().item(), res.truncated.sum().item())
single = MultiHoverAviary(num_drones=2)
o, info = single.reset()
o, r, term, trunc, info = single.step(single.action_space.sample())
print(o.shape, r, term, trunc, info)
flock = FlockAviary(num_drones=3)
o, info = flock.reset()
print(flock.step(flock.action_space.sample())[1])
hist = DeviceMAPPO(env, rollout_steps=64, mini_batch_size=16384, opt_epochs=2).learn(max_env_steps=3 * 64 * 65536)
print([round(h["ep_return"], 2) for h in hist], round(hist[-1]["step_time"], 3))
