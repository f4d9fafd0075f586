"""Device-side running normalisers against the reference's own outputs
(tests/golden/normalizers.npz, produced by the unmodified `normalization.py` classes) and against the
numpy oracle on larger seeded batches; the fused actor's normalise-on-load; MAPPO with `norm_obs`.

Tolerances: statistics are merged in fp64 from float32 one-pass batch sums (shifted by the batch's
first row) -> mean / var within 2e-6 relative of the reference's float32-numpy moments (numpy's own
float32 pairwise sums carry ~1e-7); normalised values are float32 -> 2e-5 absolute on O(1) outputs.
"""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN_DIR
from oracle.normalization import MeanStdNormalizerOracle, RewardStdNormalizerOracle

pytestmark = pytest.mark.gpu


def _close(a, b, rtol, atol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


def test_obs_normalizer_matches_reference_golden():
    from marl_gym_pybullet_drones_b200.normalization import MeanStdNormalizer
    g = np.load(os.path.join(GOLDEN_DIR, "normalizers.npz"))
    n = MeanStdNormalizer(shape=g["x"].shape[2:], clip=10, epsilon=1e-8, device="cuda:0")
    for t in range(g["x"].shape[0]):
        y = n(torch.as_tensor(g["x"][t], device="cuda:0"))
        assert _close(n.rms.mean.cpu().numpy(), g["mean"][t], 2e-6, 2e-6), t
        assert _close(n.rms.var.cpu().numpy(), g["var"][t], 2e-6, 1e-9), t
        assert abs(n.rms.count - g["count"][t]) < 1e-9
        assert _close(y.cpu().numpy(), g["y"][t], 2e-5, 2e-5), t
    n.set_read_only()
    y = n(torch.as_tensor(g["x_eval"], device="cuda:0"))
    assert _close(y.cpu().numpy(), g["y_eval"], 2e-5, 2e-5)
    assert (np.abs(y.cpu().numpy()) == 10.0).any()                              # the clip is exercised
    assert _close(n.rms.mean.cpu().numpy(), g["mean"][-1], 2e-6, 2e-6)          # frozen
    sd = n.state_dict()                                                         # normalization.py:90-96
    assert set(sd) == {"mean", "var"} and sd["mean"].shape == g["x"].shape[2:] and sd["mean"].dtype == np.float64
    n2 = MeanStdNormalizer(shape=g["x"].shape[2:], clip=10, epsilon=1e-8, device="cuda:0")
    n2.load_state_dict(sd)
    n2.set_read_only()
    assert torch.equal(n2(torch.as_tensor(g["x_eval"], device="cuda:0")), y)


def test_reward_normalizer_matches_reference_golden():
    from marl_gym_pybullet_drones_b200.normalization import RewardStdNormalizer
    g = np.load(os.path.join(GOLDEN_DIR, "normalizers.npz"))
    n = RewardStdNormalizer(gamma=0.99, clip=10, epsilon=1e-8, device="cuda:0")
    for t in range(g["r"].shape[0]):
        y = n(torch.as_tensor(g["r"][t], device="cuda:0"), torch.as_tensor(g["d"][t], device="cuda:0"))
        assert _close(y.cpu().numpy(), g["ry"][t], 1e-12, 1e-12), t              # float64 in, float64 statistics
        assert abs(float(n.var) - g["rvar"][t]) <= 1e-12 * g["rvar"][t]


@pytest.mark.parametrize("rows,shape", [(1, (3,)), (7, (2, 9)), (4096, (4, 72)), (100003, (1, 27)), (65536, (16, 72))])
def test_running_moments_match_oracle_on_large_batches(rows, shape):
    """Ragged row counts (not a multiple of the slab / unroll), one row, wide and narrow rows."""
    from marl_gym_pybullet_drones_b200.normalization import MeanStdNormalizer
    rng = np.random.default_rng(rows)
    o = MeanStdNormalizerOracle(shape=shape, clip=5.0, epsilon=1e-8)
    n = MeanStdNormalizer(shape=shape, clip=5.0, epsilon=1e-8, device="cuda:0")
    scale, off = rng.uniform(0.01, 20, shape), rng.uniform(-100, 100, shape)     # |mean| >> std columns included
    from oracle.normalization import RunningMeanStdOracle
    tm = RunningMeanStdOracle(shape=shape)
    for t in range(3):
        x = (rng.standard_normal((rows,) + shape) * scale + off).astype(np.float32)
        yo = o(x)
        y = n(torch.as_tensor(x, device="cuda:0"))
        # The kernel follows the float64 value of the reference's formulas.  numpy reduces axis 0 of a float32
        # array by sequential float32 accumulation, so the reference's own moments drift from that value with
        # the batch size (measured at 65 536 rows: mean 2.6e-4 of |offset| + spread, variance 9 % on a column
        # whose spread is 1e-4 of its offset); up to 4096 rows the drift stays below 8 sqrt(rows) 2^-24 and
        # the oracle is compared directly.
        tm.update(x.astype(np.float64))
        mean, var = n.rms.mean.cpu().numpy(), n.rms.var.cpu().numpy()
        assert _close(mean, tm.mean, 1e-6, 1e-6)
        assert _close(var, tm.var, 2e-5, 1e-12)
        if rows > 1:
            yt = np.clip((x - tm.mean) / np.sqrt(tm.var + 1e-8), -5.0, 5.0)
            assert np.all(np.abs(y.cpu().numpy() - yt) <= 3e-5 + 3e-5 * np.abs(yt) + 2e-6 * np.abs(off / scale))
        if rows <= 4096:
            loose = 8 * np.sqrt(rows) * 2.0 ** -24
            assert np.all(np.abs(mean - o.rms.mean) <= 1e-6 + loose * (np.abs(off) + scale))
            small = np.abs(off) < 100 * scale          # the reference's variance is only meaningful there
            assert _close(var[small], o.rms.var[small], 1e-3, 1e-9)
            if rows > 1:
                ok = np.abs(y.cpu().numpy() - yo) <= 1e-3 + 1e-3 * np.abs(yo) + loose * (1 + np.abs(off / scale))
                assert ok.all()
    assert n.rms._lib.bd_rms_launch_count(n.rms._h) == 3 * 2   # one fused moments launch + one normalise launch per call


def test_unaligned_batch_takes_the_scalar_path():
    """A batch whose base pointer is not 16-byte aligned (a view into a larger buffer) cannot use 128-bit loads."""
    from marl_gym_pybullet_drones_b200.normalization import MeanStdNormalizer
    rng = np.random.default_rng(11)
    rows, shape = 1001, (5, 119)                      # Spiral's observation shape, odd width
    x = (rng.standard_normal((rows,) + shape) * 2 + 3).astype(np.float32)
    buf = torch.zeros(x.size + 8, device="cuda:0")
    for off in (0, 1, 2):
        view = buf[off:off + x.size].view((rows,) + shape)
        view.copy_(torch.as_tensor(x))
        assert view.is_contiguous() and (view.data_ptr() % 16 == 0) == (off == 0)
        n = MeanStdNormalizer(shape=shape, clip=10.0, device="cuda:0")
        y = n(view)
        o = MeanStdNormalizerOracle(shape=shape, clip=10.0)
        yo = o(x.astype(np.float64))
        assert _close(n.rms.mean.cpu().numpy(), o.rms.mean, 1e-6, 1e-6)
        assert _close(n.rms.var.cpu().numpy(), o.rms.var, 2e-5, 1e-12)
        assert _close(y.cpu().numpy(), yo, 3e-5, 3e-5)


def test_moments_of_sharded_batches_merge_like_one_batch():
    """What ranks exchange when envs are sharded (bd_rms_batch_moments -> all-gather -> bd_rms_merge_moments):
    merging the parts of a batch == updating with the whole batch (== np.mean / np.var of the concatenation)."""
    from marl_gym_pybullet_drones_b200.normalization import RunningMeanStd
    rng = np.random.default_rng(5)
    shape = (2, 72)
    whole, parts = RunningMeanStd(shape=shape, device="cuda:0"), RunningMeanStd(shape=shape, device="cuda:0")
    for t in range(3):
        x = torch.as_tensor((rng.standard_normal((600,) + shape) * 3 + rng.uniform(-9, 9, shape)).astype(np.float32),
                            device="cuda:0")
        whole.update(x)
        cut = [0, 100, 101, 350, 600]          # ragged shards, one of a single row
        m = torch.stack([parts.batch_moments(x[a:b].contiguous()) for a, b in zip(cut[:-1], cut[1:])])
        assert m[:, -1].tolist() == [100.0, 1.0, 249.0, 250.0]
        parts.merge_moments(m)
        # float32 partial sums are grouped differently in the two routes: ~2^-24 per accumulated row slab
        assert _close(parts.mean.cpu().numpy(), whole.mean.cpu().numpy(), 1e-6, 1e-6)
        assert _close(parts.var.cpu().numpy(), whole.var.cpu().numpy(), 1e-5, 1e-12)
        xt = x.double().cpu().numpy()
        assert parts.count == whole.count
    m1, r1 = parts.stats()
    m2, r2 = whole.stats()
    assert torch.allclose(m1, m2, atol=1e-6) and torch.allclose(r1, r2, rtol=1e-6)


def test_fused_actor_normalises_on_load():
    """`bd_actor_set_input_norm` == normalise in torch, then run the same kernel on the result
    (identical bf16 operands -> identical outputs), with per-agent statistics (period = M)."""
    from marl_gym_pybullet_drones_b200.actor import FusedActor
    from marl_gym_pybullet_drones_b200.mappo import MLP
    torch.manual_seed(0)
    M, D, H, A, N = 4, 72, 256, 4, 3001
    mlp = MLP(D, A, [H, H], "tanh").cuda()
    logstd = torch.full((A,), -0.5, device="cuda")
    obs = torch.randn(N, M, D, device="cuda") * 7 + 3
    mean = torch.randn(M * D, device="cuda") * 3
    rstd = 1.0 / (0.05 + 5 * torch.rand(M * D, device="cuda"))
    noise = torch.randn(N * M, A, device="cuda")
    fa = FusedActor(D, H, A)
    fa.set_weights(mlp, logstd)
    fa.set_input_norm(mean, rstd, period=M, clip=4.0)
    a1, l1, m1 = fa.forward(obs.view(N * M, D), noise=noise, want_mean=True)
    fa.set_input_norm(None)
    pre = ((obs.view(N, M * D) - mean) * rstd).clamp(-4.0, 4.0).view(N * M, D).contiguous()
    assert (pre.abs() == 4.0).any()
    a2, l2, m2 = fa.forward(pre, noise=noise, want_mean=True)
    torch.cuda.synchronize()
    # x - mean, * rstd, clamp are the same float32 ops in both paths (no fma contraction across them in torch);
    # allow one bf16 ulp on a handful of inputs that round differently
    assert (m1 - m2).abs().max().item() < 2e-2 and (m1 - m2).abs().mean().item() < 1e-4
    assert torch.equal(l1, l2)
    fa.close()


def test_mappo_with_obs_and_reward_normalisation(tmp_path):
    """norm_obs / norm_reward (mappo/config.py:7-10; the Spiral run turns both on): the rollout keeps raw
    observations, statistics follow every observed batch, log-probs stored by the fused kernel equal the
    fp32 actor on the normalised observations, training stays finite, checkpoint round-trips."""
    from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO
    grid = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5]])
    env = BatchAviary(task="multihover", num_envs=256, num_drones=2, initial_xyzs=grid, seed=4, track_episode_stats=True)
    algo = DeviceMAPPO(env, rollout_steps=24, hidden_dim=256, norm_obs=True, norm_reward=True, seed=2,
                       mini_batch_size=2048, opt_epochs=2)
    assert algo.fused is not None
    algo.collect_rollout()
    T, N, M, D, A = algo.T, algo.N, algo.M, algo.D, algo.A
    # statistics = oracle fed with the same raw batches
    o = MeanStdNormalizerOracle(shape=(M, D), clip=10, epsilon=1e-8)
    raw = algo.obs.cpu().numpy()
    for t in range(T + 1):
        o(raw[t])
        assert _close(algo.nmean[t].cpu().numpy().reshape(M, D), o.rms.mean, 1e-5, 1e-5), t
        assert _close(1.0 / algo.nrstd[t].cpu().numpy().reshape(M, D), np.sqrt(o.rms.var + 1e-8), 2e-4, 1e-6), t
    assert abs(algo.obs_normalizer.rms.count - (1e-4 + (T + 1) * N)) < 1e-6
    with torch.no_grad():
        normed = torch.stack([algo._normed(algo.obs[t], t) for t in range(T)])
        lp = algo.ac.logp(normed.reshape(-1, D), algo.act.reshape(-1, A))
    ratio = torch.exp(lp - algo.logp.reshape(-1, 1))
    assert abs(ratio.mean().item() - 1.0) < 5e-3 and (ratio - 1).abs().max().item() < 0.2
    assert algo.rew.abs().max().item() <= 10.0
    stats = algo.train_step()
    assert np.isfinite([stats["policy_loss"], stats["value_loss"], stats["approx_kl"]]).all()
    p = tmp_path / "model_latest.pt"
    algo.save(p)
    sd = torch.load(p, weights_only=False)
    assert set(sd["obs_normalizer"]) == {"mean", "var"} and sd["obs_normalizer"]["mean"].shape == (M, D)
    assert sd["obs"].shape == (N, M, D) and np.abs(sd["obs"]).max() <= 10.0
    assert "actor.logstd" in sd["agent"]["ac"] and "critic.v_net.fcs.0.weight" in sd["agent"]["ac"]
    algo2 = DeviceMAPPO(env, rollout_steps=24, hidden_dim=256, norm_obs=True, norm_reward=True, seed=9)
    algo2.load(p)
    ob = algo.obs[T]
    assert torch.allclose(algo.select_action(ob), algo2.select_action(ob), atol=1e-6)
    env.close()
