"""The README's swarm / evaluation snippet (smoke check)."""
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO, FlockAviary, make_vec_envs  # noqa: E402

venv = make_vec_envs(functools.partial(FlockAviary, num_drones=4), batch_size=4096)
obs, info = venv.reset()
print(obs.shape)
venv.close()
algo = DeviceMAPPO(BatchAviary(task="flock", num_envs=4096, num_drones=4, track_episode_stats=True), norm_obs=True,
                   rollout_steps=64, mini_batch_size=16384)
hist = algo.learn(max_env_steps=2_000_000)
print(len(hist), [round(h["ep_return"], 3) for h in hist[:2] + hist[-2:]])
print(algo.run(n_episodes=16)["ep_returns"][:4])
algo.save("/tmp/model_latest.pt")
algo2 = DeviceMAPPO(BatchAviary(task="flock", num_envs=64, num_drones=4, track_episode_stats=True), norm_obs=True)
algo2.load("/tmp/model_latest.pt")
print(algo2.run(n_episodes=16)["ep_returns"][:4])
