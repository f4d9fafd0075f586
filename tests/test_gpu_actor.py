"""Fused tcgen05 actor forward vs a plain PyTorch fp32 reference of the same op
(mean within bf16-operand tolerance; sampling identities exact)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(mlp, logstd, obs, noise):
    with torch.no_grad():
        mean = mlp(obs)
        act = mean + logstd.exp() * noise
        logp = (-0.5 * noise.pow(2) - logstd - 0.5 * math.log(2 * math.pi)).sum(-1)
    return mean, act, logp


@pytest.mark.parametrize("obs_dim,hidden,act_dim,rows", [(72, 256, 4, 128), (72, 256, 4, 100000), (27, 256, 1, 777),
                                                         (119, 128, 4, 5000), (72, 128, 4, 3000), (72, 64, 4, 3000),
                                                         # hidden 256 = the TMEM-resident kernel: loader fast path (obs_dim % 8 == 0,
                                                         # <= 80) at its limits, the float4 and scalar loader paths, a grid with
                                                         # more tiles than SMs x 2 and a ragged last tile, act_dim 2 and 3
                                                         (80, 256, 4, 1000), (64, 256, 2, 333), (8, 256, 4, 129), (76, 256, 4, 500),
                                                         (57, 256, 3, 40001),
                                                         # one observation buffer instead of two (K1 = 96), float4 loader path
                                                         (88, 256, 4, 20000), (96, 256, 4, 1000)])
def test_fused_actor_matches_torch_fp32(obs_dim, hidden, act_dim, rows):
    from marl_gym_pybullet_drones_b200.actor import FusedActor
    from marl_gym_pybullet_drones_b200.mappo import MLP
    torch.manual_seed(obs_dim + hidden + rows)
    mlp = MLP(obs_dim, act_dim, [hidden, hidden], "tanh").cuda()
    logstd = (-0.5 + 0.1 * torch.randn(act_dim)).cuda()
    obs = torch.randn(rows, obs_dim, device="cuda")
    obs[:, :3] *= 3.0
    noise = torch.randn(rows, act_dim, device="cuda")
    fa = FusedActor(obs_dim, hidden, act_dim)
    fa.set_weights(mlp, logstd)
    act, logp, mean = fa.forward(obs, noise=noise, want_mean=True)
    torch.cuda.synchronize()
    rmean, ract, rlogp = _ref(mlp, logstd, obs, noise)
    # bf16 operands (8-bit mantissa), fp32 accumulation, tanh.approx: |mean error| well below the policy noise (std 0.6)
    err = (mean - rmean).abs().max().item()
    assert err < 3e-2, err
    assert (mean - rmean).abs().mean().item() < 4e-3
    assert torch.allclose(act, mean + logstd.exp() * noise, atol=1e-5)        # sampling identity is exact
    assert torch.allclose(logp, rlogp, atol=1e-4)
    assert fa.launch_count >= 8
    fa.close()


def test_fused_actor_philox_noise_is_standard_normal_and_reproducible():
    from marl_gym_pybullet_drones_b200.actor import FusedActor
    from marl_gym_pybullet_drones_b200.mappo import MLP
    torch.manual_seed(0)
    mlp = MLP(72, 4, [256, 256], "tanh").cuda()
    logstd = torch.full((4,), -0.5, device="cuda")
    obs = torch.randn(200000, 72, device="cuda")
    a, b = FusedActor(72, 256, 4), FusedActor(72, 256, 4)
    a.set_weights(mlp, logstd); b.set_weights(mlp, logstd)
    act1, lp1, mean1 = a.forward(obs, seed=5, want_mean=True)
    act2, lp2 = b.forward(obs, seed=5)
    assert torch.equal(act1, act2) and torch.equal(lp1, lp2)
    act3, _ = a.forward(obs, seed=5)                      # next call: new stream offset
    assert not torch.equal(act1, act3)
    eps = (act1 - mean1) / logstd.exp()
    assert abs(eps.mean().item()) < 5e-3 and abs(eps.std().item() - 1.0) < 5e-3
    assert abs((eps ** 4).mean().item() - 3.0) < 0.05     # Gaussian kurtosis
    want_lp = (-0.5 * eps.pow(2) - logstd - 0.5 * math.log(2 * math.pi)).sum(-1)
    assert torch.allclose(lp1, want_lp, atol=2e-3)
    a.close(); b.close()


def test_fused_actor_rejects_shapes_that_do_not_fit_shared_memory():
    from marl_gym_pybullet_drones_b200._native import NativeError
    from marl_gym_pybullet_drones_b200.actor import FusedActor
    with pytest.raises(NativeError):      # K1 = 128: layer-1 staging (96 KB) + resident W2 (128 KB) exceed 227 KB
        FusedActor(119, 256, 4)
    with pytest.raises(NativeError):
        FusedActor(72, 100, 4)
