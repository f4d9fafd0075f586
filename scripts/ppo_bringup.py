"""Bring-up of the PPO-update kernels: per-parameter gradient error vs torch autograd, and timings of the pieces at the config-4 minibatch shape."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import test_gpu_ppo_kernels as t
    from marl_gym_pybullet_drones_b200 import ppo_native
    from marl_gym_pybullet_drones_b200.ppo_native import PpoNet
    T, N, M, D, A, samples = 8, 200, 4, 72, 4, 1024
    obs, act, g = t._rollout(T, N, M, D, A, 5)
    mlp = t._mlp(D, A, 6)
    logstd = torch.nn.Parameter(torch.tensor([-0.5, -0.3, -0.7, -0.5], device="cuda"))
    with torch.no_grad():
        d0 = torch.distributions.Normal(mlp(obs[:T].reshape(-1, D)) + 0.05, logstd.exp())
        logp_old = d0.log_prob(act.reshape(-1, A)).sum(-1).reshape(T, N, M).contiguous()
    adv = torch.randn((T, N), device="cuda", generator=g)
    stats2 = torch.tensor([float(adv.mean()), 1.0 / (float(adv.std(unbiased=False)) + 1e-8)], device="cuda")
    adv_n = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
    idx = torch.randperm(T * N, device="cuda", generator=g)[:samples].contiguous()
    pl, el, kl = t._actor_reference(mlp, logstd, obs, act, logp_old, adv_n, idx, T, N, M, D, A, 0.2, 0.005)
    params = [logstd] + list(mlp.parameters())
    want = torch.autograd.grad(pl + 0.005 * el, params)
    for swap in ("0",):
        net = PpoNet(D, 1, A, True, samples * M)
        net.pack(t._flat(params))
        out = net.forward(obs, N, M, samples * M, idx=idx)
        x = obs[:T].reshape(T * N, M, D)[idx].reshape(-1, D)
        print(f"swap={swap} forward max err {float((out - mlp(x)).abs().max()):.3e}", flush=True)
        grad = torch.zeros(net.param_count, device="cuda")
        net.grad(grad, obs, N, M, idx, samples, critic=False, act=act, logp_old=logp_old, adv=adv, adv_stats=stats2,
                 clip=0.2, entropy_coef=0.005)
        torch.cuda.synchronize()
        st = net.stats
        print(f"swap={swap} loss {float(st[0]) / float(st[2]):.5f} vs {float(pl):.5f}  kl {float(st[1]) / float(st[2]):.5f} vs {float(kl):.5f}")
        off = 0
        for name, w in zip(["logstd", "W1", "b1", "W2", "b2", "W3", "b3"], want):
            got = grad[off:off + w.numel()].view_as(w)
            off += w.numel()
            print(f"  {name:6s} rel {t._rel(got, w):.3e} cos {t._cos(got, w):+.5f} |ref| {float(w.norm()):.3e}")
        net.close()
    # ---- timings at the config-4 minibatch shape: 32 768 samples x 4 agents
    T, N, samples = 32, 16384, 32768
    obs, act, g = t._rollout(T, N, M, D, A, 7)
    logp_old = torch.randn((T, N, M), device="cuda", generator=g) * 0.1 - 3
    adv = torch.randn((T, N), device="cuda", generator=g)
    ret = torch.randn((T, N), device="cuda", generator=g)
    idx = torch.randperm(T * N, device="cuda", generator=g)[:samples].contiguous()
    actor = PpoNet(D, 1, A, True, samples * M)
    critic = PpoNet(D, M, 1, False, samples)
    actor.pack(t._flat(params))
    cm = t._mlp(M * D, 1, 3)
    critic.pack(t._flat(list(cm.parameters())))
    ga, gc = torch.zeros(actor.param_count, device="cuda"), torch.zeros(critic.param_count, device="cuda")

    def timeit(f, n=20):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    ta = timeit(lambda: actor.grad(ga, obs, N, M, idx, samples, critic=False, act=act, logp_old=logp_old, adv=adv,
                                   adv_stats=stats2, clip=0.2, entropy_coef=0.005))
    tc = timeit(lambda: critic.grad(gc, obs, N, M, idx, samples, critic=True, ret=ret))
    tf = timeit(lambda: actor.forward(obs, N, M, samples * M, idx=idx))
    print(f"actor grad {ta:.1f} us, critic grad {tc:.1f} us, actor forward {tf:.1f} us per minibatch of {samples} samples x {M}")


if __name__ == "__main__":
    main()
