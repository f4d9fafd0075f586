"""On-device MAPPO: GAE scan equals the reference's per-sequence numpy loop, the loop
learns, checkpoints round-trip.  GPU only (the envs are CUDA kernels)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_returns(rews, vals, masks, last_val, gamma, lam):
    """`mappo/buffer.py:561-614` (_compute_single_agent_returns) with terminal_vals = 0, use_gae=True."""
    T = len(rews)
    rets, advs = np.zeros(T), np.zeros(T)
    ext = np.concatenate([vals, [last_val]])
    ret, adv = last_val, 0.0
    for i in reversed(range(T)):
        ret = rews[i] + gamma * masks[i] * ret
        td = rews[i] + gamma * masks[i] * ext[i + 1] - vals[i]
        adv = adv * lam * gamma * masks[i] + td
        rets[i], advs[i] = ret, adv
    return rets, advs


@pytest.mark.parametrize("rollout_values", ["zeros", "critic"])
def test_gae_scan_matches_reference_loop(rollout_values):
    from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO
    env = BatchAviary(task="multihover", num_envs=16, num_drones=2, track_episode_stats=True,
                      initial_xyzs=np.array([[0.0, 0.0, 0.2], [1.0, 0.0, 0.2]]))
    algo = DeviceMAPPO(env, rollout_steps=48, hidden_dim=32, rollout_values=rollout_values, seed=3)
    algo.collect_rollout()
    algo.compute_returns()
    rew, val = algo.rew.cpu().numpy().astype(np.float64), algo.val.cpu().numpy().astype(np.float64)[..., 0]
    mask = 1.0 - (algo.term | algo.trunc).cpu().numpy().astype(np.float64)
    assert (mask == 0).any()                       # crashes happened: the mask path is exercised
    if rollout_values == "zeros":
        assert np.all(val[:-1] == 0)               # agent.py:413
    for e in range(16):
        rets, advs = _ref_returns(rew[:, e], val[:-1, e], mask[:, e], val[-1, e], 0.99, 0.95)
        assert np.allclose(algo.ret.cpu().numpy()[:, e, 0], rets, rtol=2e-5, atol=2e-5)
        assert np.allclose(algo.adv.cpu().numpy()[:, e, 0], advs, rtol=2e-5, atol=2e-5)
    a = algo.adv.cpu().numpy()
    assert np.allclose(algo.adv_n.cpu().numpy(), (a - a.mean()) / (a.std() + 1e-8), atol=1e-4)      # np.std, ddof = 0 (buffer.py:666-695)
    env.close()


def test_hover_learns_and_checkpoint_roundtrip(tmp_path):
    from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO
    torch.manual_seed(0)
    env = BatchAviary(task="hover", num_envs=1024, act="one_d_rpm", seed=1, track_episode_stats=True)
    algo = DeviceMAPPO(env, rollout_steps=121, hidden_dim=64, mini_batch_size=8192, opt_epochs=4,
                       rollout_values="critic", actor_lr=1e-3, target_kl=0.05, seed=0)
    assert algo.fused is not None          # the rollout policy runs on the tcgen05 kernel
    first = algo.train_step()
    assert first["episodes"] == 0 or np.isfinite(first["ep_return"])
    hist = algo.learn(max_env_steps=algo.total_env_steps + 14 * 121 * 1024)
    rets = [h["ep_return"] for h in hist if h["episodes"] > 0]
    assert len(rets) >= 4
    # random policy hovers around z0: ~1.38/step; a learnt policy climbs towards z=1 (2/step)
    assert np.mean(rets[-2:]) > np.mean(rets[:2]) + 10.0, rets
    assert all(np.isfinite(list(h.values())).all() for h in hist)
    assert algo.fused.launch_count > 15 * 121
    p = tmp_path / "model_latest.pt"
    algo.save(p)
    algo2 = DeviceMAPPO(env, rollout_steps=121, hidden_dim=64, seed=5)
    algo2.load(p)
    o = algo.obs[0]
    assert torch.equal(algo.select_action(o), algo2.select_action(o))
    assert algo2.total_env_steps == algo.total_env_steps
    env.close()


def test_fused_and_torch_rollouts_agree_statistically():
    """Same policy, fused (bf16 tensor cores, Philox) vs torch (fp32, torch RNG) rollouts: the stored
    log-probs are self-consistent with the fp32 actor to bf16 accuracy (PPO ratio ~ 1 before any update)."""
    from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO
    grid = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5], [0.0, 1.0, 0.5], [1.0, 1.0, 0.5]])
    env = BatchAviary(task="multihover", num_envs=512, num_drones=4, initial_xyzs=grid, seed=2,
                      track_episode_stats=True)
    algo = DeviceMAPPO(env, rollout_steps=16, hidden_dim=256, seed=1)
    assert algo.fused is not None
    algo.collect_rollout()
    T, N, M, D, A = algo.T, algo.N, algo.M, algo.D, algo.A
    with torch.no_grad():
        lp = algo.ac.logp(algo.obs[:T].reshape(-1, D), algo.act.reshape(-1, A))
    ratio = torch.exp(lp - algo.logp.reshape(-1, 1))
    assert abs(ratio.mean().item() - 1.0) < 5e-3 and (ratio - 1).abs().max().item() < 0.15
    eps = (algo.act.reshape(-1, A) - algo.ac.actor(algo.obs[:T].reshape(-1, D)).detach()) / algo.ac.logstd.exp().detach()
    assert abs(eps.mean().item()) < 2e-2 and abs(eps.std().item() - 1.0) < 2e-2
    env.close()


def test_graphed_update_equals_eager_update_and_kl_gate_closes():
    """The PPO minibatch replayed as a CUDA graph == the same ops launched eagerly (same permutations), and
    the device-side KL gate skips whole optimiser steps (agent.py:731) without a host read."""
    from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO
    env = BatchAviary(task="multihover", num_envs=512, num_drones=2, seed=3, track_episode_stats=True,
                      initial_xyzs=np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5]]))
    algo = DeviceMAPPO(env, rollout_steps=16, hidden_dim=64, mini_batch_size=1024, opt_epochs=3, actor_lr=3e-3,
                       target_kl=0.004, seed=1, graph_update=True)
    algo.collect_rollout()
    algo.compute_returns()
    opts = (algo.actor_opt, algo.critic_opt)
    snap = [(o.flat.clone(), o.exp_avg.clone(), o.exp_avg_sq.clone(), o.step_t.clone()) for o in opts]
    gen = algo.gen.get_state()
    res_g = algo.update()
    assert algo._graph is not None
    out_g = [o.flat.clone() for o in opts]
    steps_g = [float(o.step_t) for o in opts]
    for o, (f, m, v, st) in zip(opts, snap):
        o.flat.copy_(f); o.exp_avg.copy_(m); o.exp_avg_sq.copy_(v); o.step_t.copy_(st)
    algo.gen.set_state(gen)
    algo.cfg["graph_update"] = False
    res_e = algo.update()
    n_mb = 3 * (16 * 512 // 1024)
    assert steps_g == [float(o.step_t) for o in opts]
    assert steps_g[1] == n_mb and 0 < steps_g[0] < n_mb, steps_g      # the gate closed on some minibatches
    for a, o in zip(out_g, opts):
        assert torch.allclose(a, o.flat, rtol=1e-5, atol=1e-6)
    for k in res_g:
        assert abs(res_g[k] - res_e[k]) <= 1e-5 * max(1.0, abs(res_e[k])), k
    env.close()


def test_run_evaluates_deterministically_like_mappo_run():
    """`DeviceMAPPO.run` (reference `MAPPO.run`, mappo.py:534-581): n episodes with the mean action, returns and
    lengths per episode; the untrained hover policy holds altitude-ish and reaches the 242-step truncation or an
    earlier bound, and two evaluations agree exactly (deterministic actions, fixed spawn)."""
    from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO
    env = BatchAviary(task="hover", num_envs=64, act="one_d_rpm", seed=1, track_episode_stats=True)
    algo = DeviceMAPPO(env, rollout_steps=8, hidden_dim=64, seed=0)
    r1 = algo.run(n_episodes=5)
    r2 = algo.run(n_episodes=5)
    assert r1["ep_returns"].shape == (5,) and r1["ep_lengths"].shape == (5,)
    assert np.array_equal(r1["ep_returns"], r2["ep_returns"]) and np.array_equal(r1["ep_lengths"], r2["ep_lengths"])
    assert (r1["ep_lengths"] >= 1).all() and (r1["ep_lengths"] <= 242).all() and np.isfinite(r1["ep_returns"]).all()
    assert len(set(r1["ep_lengths"].tolist())) == 1           # identical envs, identical deterministic episodes
    env.close()


# ------------------------------------------------------------------------------------------------------------------
# the hand-written update (csrc/bd_ppo.cu): hidden_dim = 256 selects it (update_impl="auto")
# ------------------------------------------------------------------------------------------------------------------
def _mh_env(N, M=2, seed=3):
    from marl_gym_pybullet_drones_b200 import BatchAviary
    grid = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5], [0.0, 1.0, 0.5], [1.0, 1.0, 0.5]])[:M]
    return BatchAviary(task="multihover", num_envs=N, num_drones=M, seed=seed, track_episode_stats=True, initial_xyzs=grid)


def test_native_update_matches_torch_update_on_the_same_rollout():
    """One epoch of minibatches on the same rollout, same permutation: the native kernels (bf16 tensor cores) and the
    torch autograd update (fp32) report the same losses / approx_kl (<= 2e-2 relative, floor 1e-2) and move the
    parameters in the same direction (cosine of the parameter updates >= 0.98)."""
    from marl_gym_pybullet_drones_b200 import DeviceMAPPO
    env = _mh_env(512, 2)
    kw = dict(rollout_steps=16, hidden_dim=256, mini_batch_size=2048, opt_epochs=1, seed=1, rollout_values="critic",
              target_kl=0.0, matmul_precision="fp32")
    nat = DeviceMAPPO(env, update_impl="native", **kw)
    ref = DeviceMAPPO(env, update_impl="torch", graph_update=False, **kw)
    assert nat.native and not ref.native
    nat.collect_rollout()
    nat.compute_returns()
    for name in ("obs", "act", "logp", "val", "rew", "term", "trunc", "ret", "adv"):
        getattr(ref, name).copy_(getattr(nat, name))
    ref._adv_stats.copy_(nat._adv_stats)
    ref.adv_n.copy_((ref.adv - ref._adv_stats[0]) * ref._adv_stats[1])
    ref.total_env_steps = nat.total_env_steps
    ref._reset_done = True
    before = [o.flat.clone() for o in (nat.actor_opt, nat.critic_opt)]
    for o_r, o_n in zip((ref.actor_opt, ref.critic_opt), (nat.actor_opt, nat.critic_opt)):
        assert torch.equal(o_r.flat, o_n.flat)          # same seed, same initial weights
    ref.gen.set_state(nat.gen.get_state())
    res_n = nat.update()
    res_r = ref.update()
    for k in res_r:
        assert abs(res_n[k] - res_r[k]) <= 2e-2 * max(abs(res_r[k]), 1e-2), (k, res_n[k], res_r[k])
    for b, o_r, o_n in zip(before, (ref.actor_opt, ref.critic_opt), (nat.actor_opt, nat.critic_opt)):
        dn, dr = (o_n.flat - b), (o_r.flat - b)
        cos = float(torch.dot(dn, dr) / (dn.norm() * dr.norm()))
        assert cos >= 0.98, cos
        assert float(o_n.step_t) == float(o_r.step_t) == 16 * 512 // 2048
    env.close()


def test_native_epoch_graph_equals_eager_and_kl_gate_closes():
    from marl_gym_pybullet_drones_b200 import DeviceMAPPO
    env = _mh_env(512, 2)
    algo = DeviceMAPPO(env, rollout_steps=16, hidden_dim=256, mini_batch_size=1024, opt_epochs=3, actor_lr=3e-3,
                       target_kl=0.004, seed=1, graph_update=True)
    assert algo.native
    algo.collect_rollout()
    algo.compute_returns()
    opts = (algo.actor_opt, algo.critic_opt)
    snap = [(o.flat.clone(), o.exp_avg.clone(), o.exp_avg_sq.clone(), o.step_t.clone()) for o in opts]
    gen = algo.gen.get_state()
    res_g = algo.update()
    assert algo._graph is not None
    out_g = [o.flat.clone() for o in opts]
    steps_g = [float(o.step_t) for o in opts]
    for o, (f, m, v, st) in zip(opts, snap):
        o.flat.copy_(f); o.exp_avg.copy_(m); o.exp_avg_sq.copy_(v); o.step_t.copy_(st)
    algo._pack_native()
    algo.gen.set_state(gen)
    algo.cfg["graph_update"] = False
    res_e = algo.update()
    n_mb = 3 * (16 * 512 // 1024)
    assert steps_g == [float(o.step_t) for o in opts]
    assert steps_g[1] == n_mb and 0 < steps_g[0] < n_mb, steps_g      # the gate closed on some minibatches
    assert float(algo._gates) == 2 * steps_g[0]                       # counted on the device, both runs
    for a, o in zip(out_g, opts):
        assert torch.allclose(a, o.flat, rtol=1e-4, atol=1e-6)        # atomics order in the statistics only
    for k in res_g:
        assert abs(res_g[k] - res_e[k]) <= 1e-4 * max(1.0, abs(res_e[k])), k
    env.close()


def test_native_update_learns_hover_and_resumes_reproducibly(tmp_path):
    """End to end with every hot operation on this repo's kernels: step kernel, fused actor, GAE scan, fused PPO
    update.  The policy improves; a checkpoint restores weights, optimiser, RNG streams (mappo.py:203-270)."""
    from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO
    env = BatchAviary(task="hover", num_envs=1024, act="one_d_rpm", seed=1, track_episode_stats=True)
    kw = dict(rollout_steps=121, hidden_dim=256, mini_batch_size=8192, opt_epochs=4, rollout_values="critic",
              actor_lr=1e-3, target_kl=0.05)
    algo = DeviceMAPPO(env, seed=0, **kw)
    assert algo.native and algo.fused is not None
    hist = algo.learn(max_env_steps=15 * 121 * 1024)
    rets = [h["ep_return"] for h in hist if h["episodes"] > 0]
    assert len(rets) >= 4 and np.mean(rets[-2:]) > np.mean(rets[:2]) + 10.0, rets
    assert all(np.isfinite(list(h.values())).all() for h in hist)
    p = tmp_path / "model_latest.pt"
    algo.save(p)
    sd = torch.load(p, weights_only=False)
    assert sd["random_state"] is not None and sd["env_random_state"][0]["philox"][1] > 0
    algo2 = DeviceMAPPO(env, seed=5, **kw)
    algo2.load(p)
    o = algo.obs[0].contiguous()
    assert torch.equal(algo.select_action(o), algo2.select_action(o))
    assert torch.equal(algo.gen.get_state(), algo2.gen.get_state())
    assert algo2.fused._calls == algo.fused._calls and algo2.env.get_rng_state() == algo.env.get_rng_state()
    env.close()
