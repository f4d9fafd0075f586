"""Short driver for ncu / timing of the PPO-update kernels at the config-4 minibatch shape (32 768 env-steps x 4 agents):
    python scripts/profile_ppo.py [--iters 3] [--critic] [--trace]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--samples", type=int, default=32768)
ap.add_argument("--agents", type=int, default=4)
ap.add_argument("--time", action="store_true")
args = ap.parse_args()

import test_gpu_ppo_kernels as t  # noqa: E402
from marl_gym_pybullet_drones_b200.ppo_native import PpoNet  # noqa: E402

M, D, A, samples = args.agents, 72, 4, args.samples
T, N = 32, max(1024, samples // 16)
obs, act, g = t._rollout(T, N, M, D, A, 7)
logp_old = torch.randn((T, N, M), device="cuda", generator=g) * 0.1 - 3
adv = torch.randn((T, N), device="cuda", generator=g)
ret = torch.randn((T, N), device="cuda", generator=g)
stats2 = torch.tensor([0.0, 1.0], device="cuda")
idx = torch.randperm(T * N, device="cuda", generator=g)[:samples].contiguous()
actor = PpoNet(D, 1, A, True, samples * M)
critic = PpoNet(D, M, 1, False, samples)
am, cm = t._mlp(D, A, 6), t._mlp(M * D, 1, 3)
actor.pack(t._flat([torch.full((A,), -0.5, device="cuda")] + list(am.parameters())))
critic.pack(t._flat(list(cm.parameters())))
ga, gc = torch.zeros(actor.param_count, device="cuda"), torch.zeros(critic.param_count, device="cuda")


def one():
    actor.grad(ga, obs, N, M, idx, samples, critic=False, act=act, logp_old=logp_old, adv=adv, adv_stats=stats2, clip=0.2,
               entropy_coef=0.005)
    critic.grad(gc, obs, N, M, idx, samples, critic=True, ret=ret)


for _ in range(args.iters):
    one()
torch.cuda.synchronize()
if args.time:
    for name, f in (("actor", lambda: actor.grad(ga, obs, N, M, idx, samples, critic=False, act=act, logp_old=logp_old, adv=adv,
                                                   adv_stats=stats2, clip=0.2, entropy_coef=0.005)),
                    ("critic", lambda: critic.grad(gc, obs, N, M, idx, samples, critic=True, ret=ret)),
                    ("actor forward", lambda: actor.forward(obs, N, M, samples * M, idx=idx))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
print("ok")
