"""Small end-to-end exercise of every kernel (a quick crash / NaN check on a GPU box; compute-sanitizer is not available on the pool): tile / generic step kernels incl.
swarm tasks and chunked host steps, reset, normaliser moments / normalise, fused actor with input normalisation."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO  # noqa: E402
from marl_gym_pybullet_drones_b200.normalization import MeanStdNormalizer  # noqa: E402

for task, M, N, kw in (("multihover", 4, 9001, {}), ("multihover", 3, 5000, dict(physics="dyn_gnd_drag_dw")),
                       ("flock", 5, 777, {}), ("meetup", 6, 333, dict(precision="fp64")), ("spiral", 3, 500, dict(act="vel", ctrl_freq=48)),
                       ("leaderfollower", 2, 100, {})):
    xyz = None if task == "spiral" else np.array([[0.6 * i, 0.3 * (i % 2), 0.5] for i in range(M)])
    env = BatchAviary(task=task, num_envs=N, num_drones=M, initial_xyzs=xyz, auto_reset=True, seed=3, **kw)
    env.reset_device()
    A = env.ACTION_DIM
    for t in range(5):
        a = (torch.rand((N, M, A), device="cuda") * 2 - 1).to(env.action_dtype)
        r = env.step_device(a, want_terminal_obs=(t == 2))
    h = env.step_host(a.cpu().numpy())
    print(task, M, N, float(r.reward.double().mean()), float(h["reward"].mean()))
    env.close()
for shape, rows in (((4, 72), 1000), ((5, 119), 333), ((3,), 1), ((2, 9), 7)):
    n = MeanStdNormalizer(shape=shape, device="cuda")
    x = torch.randn((rows,) + shape, device="cuda")
    y = n(x)
    print(shape, rows, float(y.abs().max()), n.rms.count)
env = BatchAviary(task="multihover", num_envs=300, num_drones=2, track_episode_stats=True,
                  initial_xyzs=np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5]]))
algo = DeviceMAPPO(env, rollout_steps=4, hidden_dim=64, mini_batch_size=256, opt_epochs=1, norm_obs=True, norm_reward=True,
                   graph_update=False)
print(algo.train_step()["value_loss"])
