"""BASELINE.json configs[1] and configs[2] run AS STATED, every env against its own oracle.

configs[1]  MultiHoverAviary, M = 2, rollout_batch_size N = 176, T = 256 control steps (`learn_mappo.py:104`),
            Physics.DYN, 240/30 Hz, driven like `SubprocVecEnv` drives it (reset-on-done, `subproc_vec_env.py:188-207`):
            default spawn layout + `MultiHoverAviary.reset`'s jitter (`MultiHoverAviary.py:83-102`, the accepted draw of
            its rejection loop injected into oracle and kernel alike), 176 DISTINCT action sequences, two action
            distributions: U(-1,1) float64 and 0.3 N(0,1) float32 (policy-like, unclipped).
configs[2]  SpiralFormationAviary, M = 5, N = 64, 240/48 Hz, RPM, ground effect + drag + downwash all on, the whole
            578-step episode horizon (580 steps), reset-on-done.  The ring start is staggered in height
            (z_i = 0.3 + 0.12 i): with all drones at z = 0.3 the reference's downwash term
            (`BaseAviary.py:798-804`, alpha ~ 1/dz^2) is singular from the second substep on — rounding-level height
            differences give forces of 1e6 N.  Plus 8 envs WITHOUT reset-on-done that fly the full 12 s so that the
            578th-step truncation is seen with the aero terms on.

Stated tolerances:
  fp64 : state (pos, quat, rpy, vel, ang_vel) and reward <= 1e-9 relative at every step of the free-running horizon,
         observations (float32 storage) <= 2.5e-7, terminated / truncated identical, for ALL envs.
  fp32 : per control step from the oracle's state (teacher-forced, every env, every step):
         plain DYN (configs[1], fast tile kernel) <= 1e-5; all aero terms (configs[2], fast tile kernel, aero flavour) <= 5e-5
         of max(|x|, 1) on env-steps where no two drones are within 3 cm of the same height (the downwash term is
         singular there, see the test), <= 2e-2 otherwise; flags identical (configs[1]: all env-steps;
         configs[2]: the well-conditioned ones).
"""
import functools

import numpy as np
import pytest
import torch

from _oracle_pool import accepted_jitter, run_oracles
from _util import rel_err

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-9
OBS_F32_TOL = 2.5e-7


def _default_xyz(M, L=0.0397):
    return np.array([[4 * L * i, 4 * L * i, 0.1125] for i in range(M)])          # BaseAviary.py:194-197


# ------------------------------------------------------------------------------------------- configs[1]
@functools.lru_cache(maxsize=None)
def _cfg1(dist):
    N, M, T = 176, 2, 256
    orig = _default_xyz(M)
    jrng = np.random.default_rng(1)
    jit = np.array([[accepted_jitter(jrng, orig) for _ in range(N)] for _ in range(T + 1)])     # (T+1,N,M,3)
    arng = np.random.default_rng(2)
    if dist == "uniform":
        actions = arng.uniform(-1, 1, (T, N, M, 4))                                             # float64
    else:
        actions = (0.3 * arng.standard_normal((T, N, M, 4))).astype(np.float32)
    kw = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    jobs = [dict(kw=kw, actions=actions[:, e], jitter=jit[:, e], auto_reset=True) for e in range(N)]
    return actions, jit, run_oracles(jobs)


def _make_cfg1_env(precision, adt):
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    return BatchAviary(task="multihover", num_envs=176, num_drones=2, pyb_freq=240, ctrl_freq=30, act="rpm",
                       precision=precision, auto_reset=True, reset_mode="jitter_buffer", action_dtype=adt,
                       keep_ang_vel=(precision == "fp64"))


@pytest.mark.parametrize("dist", ["uniform", "normal03_f32"])
def test_cfg1_fp64_all_176_envs_256_steps(dist):
    actions, jit, ref = _cfg1(dist)
    T, N = actions.shape[0], actions.shape[1]
    adt = torch.float32 if actions.dtype == np.float32 else torch.float64
    env = _make_cfg1_env("fp64", adt)
    env.set_jitter(torch.as_tensor(jit[0]))
    obs0 = env.reset_device().cpu().numpy()
    assert rel_err(obs0, ref["obs0"]) <= OBS_F32_TOL
    n_done = 0
    for t in range(T):
        env.set_jitter(torch.as_tensor(jit[t + 1]))
        r = env.step_device(torch.as_tensor(actions[t]).to("cuda", adt), want_terminal_obs=True)
        st, rates, sc = env.get_state(with_rates=True, with_step_counter=True)
        st = st.cpu().numpy()
        assert rel_err(st[..., :16], ref["states"][t][..., :16]) <= FP64_TOL, (dist, t)
        assert rel_err(rates.cpu().numpy(), ref["rates"][t]) <= FP64_TOL, (dist, t)
        assert np.array_equal(sc.cpu().numpy(), ref["stepc"][t]), (dist, t)
        assert rel_err(r.obs.cpu().numpy(), ref["obs"][t]) <= OBS_F32_TOL, (dist, t)
        assert rel_err(r.reward.cpu().numpy(), ref["reward"][t]) <= FP64_TOL, (dist, t)
        assert np.array_equal(r.terminated.cpu().numpy(), ref["terminated"][t]), (dist, t)
        assert np.array_equal(r.truncated.cpu().numpy(), ref["truncated"][t]), (dist, t)
        done = ref["terminated"][t] | ref["truncated"][t]
        if done.any():
            n_done += int(done.sum())
            assert rel_err(r.terminal_obs.cpu().numpy()[done], ref["term_obs"][t][done]) <= OBS_F32_TOL, (dist, t)
            assert rel_err(env.get_targets().cpu().numpy(), ref["targets"][t]) <= FP64_TOL, (dist, t)
    assert n_done >= 100                         # reset-on-done was exercised many times (crashes, the 242-step limit)
    env.close()


@pytest.mark.parametrize("dist", ["uniform", "normal03_f32"])
def test_cfg1_fp32_per_step_all_176_envs(dist):
    """fp32 fast tile kernel, teacher-forced from the oracle's state before every step."""
    actions, jit, ref = _cfg1(dist)
    T, N = actions.shape[0], actions.shape[1]
    env = _make_cfg1_env("fp32", torch.float32)
    env.set_jitter(torch.as_tensor(jit[0]))
    env.reset_device()
    worst = 0.0
    for t in range(T):
        if t > 0:
            s = ref["states"][t - 1]
            kin = np.concatenate([s[..., 0:7], s[..., 10:13], ref["rates"][t - 1]], axis=-1)
            env.set_state(torch.as_tensor(kin), targets=torch.as_tensor(ref["targets"][t - 1]),
                          step_counter=torch.as_tensor(ref["stepc"][t - 1], dtype=torch.int32))
        env.set_jitter(torch.as_tensor(jit[t + 1]))
        r = env.step_device(torch.as_tensor(actions[t].astype(np.float32), device="cuda"))
        st = env.get_state().cpu().numpy()
        ok = ~(ref["terminated"][t] | ref["truncated"][t])          # finished envs hold the re-spawn state (checked below)
        err = rel_err(st[ok][..., :13], ref["states"][t][ok][..., :13])
        worst = max(worst, err)
        assert err <= 1e-5, (dist, t, err)
        assert rel_err(st[~ok][..., :13], ref["states"][t][~ok][..., :13]) <= 1e-6, (dist, t)
        assert rel_err(r.reward.cpu().numpy(), ref["reward"][t]) <= 2e-5, (dist, t)
        assert np.array_equal(r.terminated.cpu().numpy(), ref["terminated"][t]), (dist, t)
        assert np.array_equal(r.truncated.cpu().numpy(), ref["truncated"][t]), (dist, t)
    print(f"cfg1 {dist}: worst fp32 per-step error {worst:.2e}")
    env.close()


# ------------------------------------------------------------------------------------------- configs[2]
def _spiral_xyz(M=5):
    return np.array([[0.4 * np.cos(2 * np.pi * i / M), 0.4 * np.sin(2 * np.pi * i / M), 0.3 + 0.12 * i] for i in range(M)])


@functools.lru_cache(maxsize=None)
def _cfg2(auto_reset):
    M = 5
    N, T = (64, 580) if auto_reset else (8, 580)
    rng = np.random.default_rng(3)
    if auto_reset:
        actions = rng.uniform(-1, 1, (T, N, M, 4))
        actions[:, N // 2:] = 0.3 * rng.standard_normal((T, N - N // 2, M, 4))
        actions = actions.astype(np.float32)
    else:
        actions = (0.02 * rng.standard_normal((T, N, M, 4))).astype(np.float32)
    kw = dict(task="spiral", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=48, act="rpm", aero=7,
              initial_xyzs=_spiral_xyz(M))
    jobs = [dict(kw=kw, actions=actions[:, e], auto_reset=auto_reset) for e in range(N)]
    return actions, run_oracles(jobs)


def _make_cfg2_env(precision, N, auto_reset):
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    return BatchAviary(task="spiral", num_envs=N, num_drones=5, pyb_freq=240, ctrl_freq=48, act="rpm",
                       initial_xyzs=_spiral_xyz(5), physics="dyn_gnd_drag_dw", precision=precision,
                       auto_reset=auto_reset, reset_mode="fixed", action_dtype=torch.float32,
                       keep_ang_vel=(precision == "fp64"))


@pytest.mark.parametrize("auto_reset", [True, False])
def test_cfg2_fp64_spiral_aero_full_horizon(auto_reset):
    actions, ref = _cfg2(auto_reset)
    T, N = actions.shape[0], actions.shape[1]
    env = _make_cfg2_env("fp64", N, auto_reset)
    obs0 = env.reset_device().cpu().numpy()
    assert rel_err(obs0, ref["obs0"]) <= OBS_F32_TOL
    n_done, n_trunc = 0, 0
    for t in range(T):
        r = env.step_device(torch.as_tensor(actions[t], device="cuda"), want_terminal_obs=auto_reset)
        st, rates, sc = env.get_state(with_rates=True, with_step_counter=True)
        st = st.cpu().numpy()
        assert rel_err(st[..., :16], ref["states"][t][..., :16]) <= FP64_TOL, (t, rel_err(st[..., :16], ref["states"][t][..., :16]))
        assert rel_err(rates.cpu().numpy(), ref["rates"][t]) <= FP64_TOL, t
        assert np.array_equal(sc.cpu().numpy(), ref["stepc"][t]), t
        assert rel_err(r.obs.cpu().numpy(), ref["obs"][t]) <= OBS_F32_TOL, t
        assert rel_err(r.reward.cpu().numpy(), ref["reward"][t]) <= FP64_TOL, t
        assert np.array_equal(r.terminated.cpu().numpy(), ref["terminated"][t]), t
        assert np.array_equal(r.truncated.cpu().numpy(), ref["truncated"][t]), t
        done = ref["terminated"][t] | ref["truncated"][t]
        n_done += int(done.sum())
        n_trunc += int(ref["truncated"][t].sum())
        if auto_reset and done.any():
            assert rel_err(r.terminal_obs.cpu().numpy()[done], ref["term_obs"][t][done]) <= OBS_F32_TOL, t
    if auto_reset:
        assert n_done >= 64
    else:
        assert ref["truncated"][577].all() and not ref["truncated"][576].any() and n_trunc >= N     # SpiralAviary.py:196
    env.close()


def test_cfg2_fp32_per_step_spiral_aero():
    """Fast tile kernel (aero flavour: ground effect + drag + downwash by warp shuffle), teacher-forced from the oracle's state.

    The downwash force on a drone is alpha exp(..), alpha = DW1 (r_prop / 4 dz)^2 (`BaseAviary.py:802`): its relative
    sensitivity to the height difference is 2/dz, so a float32 rounding of dz (6e-8 of |z| ~ 0.5) becomes a relative
    force error of ~6e-8/dz, and the force itself grows like 1/dz^2.  Per-step agreement is therefore stated by
    conditioning: env-steps whose closest pair of drones is more than 3 cm apart in height (before and after the step)
    <= 5e-5; all others (where the reference's own trajectory is dominated by the singular term) <= 2e-2, and
    nothing is asserted where the reference state has left |x| < 50 m."""
    actions, ref = _cfg2(True)
    T, N = actions.shape[0], actions.shape[1]
    env = _make_cfg2_env("fp32", N, True)
    env.reset_device()
    worst_good, worst_all, n_good, n_all = 0.0, 0.0, 0, 0

    def min_dz(st):            # (N, M, 20) -> (N,) smallest height difference between two drones of an env
        z = st[..., 2]
        d = np.abs(z[:, :, None] - z[:, None, :]) + 10.0 * np.eye(z.shape[1])[None]
        return d.min(axis=(1, 2))
    for t in range(T):
        prev = ref["states"][t - 1] if t > 0 else None
        if t > 0:
            kin = np.concatenate([prev[..., 0:7], prev[..., 10:13], ref["rates"][t - 1]], axis=-1)
            env.set_state(torch.as_tensor(kin), step_counter=torch.as_tensor(ref["stepc"][t - 1], dtype=torch.int32))
        r = env.step_device(torch.as_tensor(actions[t], device="cuda"))
        st = env.get_state().cpu().numpy()
        cur = ref["states"][t]
        sane = np.abs(cur[..., :3]).max(axis=(1, 2)) < 50.0
        ok = sane & ~(ref["terminated"][t] | ref["truncated"][t])
        good = ok & (min_dz(cur) > 0.03) & ((min_dz(prev) > 0.03) if prev is not None else True)
        err_e = np.max(np.abs(st[..., :13] - cur[..., :13]) / np.maximum(np.abs(cur[..., :13]), 1.0), axis=(1, 2))
        if good.any():
            worst_good = max(worst_good, float(err_e[good].max()))
            assert err_e[good].max() <= 5e-5, (t, float(err_e[good].max()))
        if ok.any():
            worst_all = max(worst_all, float(err_e[ok].max()))
            assert err_e[ok].max() <= 2e-2, (t, float(err_e[ok].max()))
        n_good += int(good.sum())
        n_all += int(ok.sum())
        assert np.array_equal(r.terminated.cpu().numpy()[good], ref["terminated"][t][good]), t
        assert np.array_equal(r.truncated.cpu().numpy()[good], ref["truncated"][t][good]), t
        assert rel_err(r.reward.cpu().numpy()[good], ref["reward"][t][good]) <= 5e-5, t
    print(f"cfg2 fp32 per step: worst {worst_good:.2e} over {n_good} well-conditioned env-steps, {worst_all:.2e} over all {n_all}")
    assert n_good > 0.5 * n_all
    env.close()


# ------------------------------------------------------------------------------------------- configs[4]
def test_cfg4_full_size_shape_properties_and_env_sharding():
    """BASELINE configs[4]'s per-GPU shape at FULL size — 131 072 envs x 16 drones, DYN + downwash, fp32 — through the
    properties that do not need an oracle: (1) env sharding is exact: the batch stepped as one handle equals
    the same envs stepped as two half-size handles (what `--gpus N` relies on: no data-path collective, downwash couples
    drones only inside an env); (2) the observation's action-history columns shift by one slot per step and end in the
    action just applied; (3) quaternions stay unit, everything stays finite; (4) the K-steps-in-one-launch kernel
    reproduces the per-step launches at this size."""
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    N, M, T = 131072, 16, 4
    side = 4
    xyz = np.array([[float(i % side) - 1.5, float(i // side) - 1.5, 0.5 + 0.07 * (i % 5)] for i in range(M)])
    mk = lambda n: BatchAviary(task="multihover", num_envs=n, num_drones=M, initial_xyzs=xyz, physics="dyn_dw",   # noqa: E731
                               precision="fp32", auto_reset=True, reset_mode="fixed", seed=1)
    whole, lo, hi, many = mk(N), mk(N // 2), mk(N // 2), mk(N)
    for e in (whole, lo, hi, many):
        e.reset_device()
    gen = torch.Generator(device="cuda").manual_seed(4)
    acts = (torch.rand((T, N, M, 4), generator=gen, device="cuda") * 2 - 1.2).contiguous()
    prev = None
    outs = []
    for t in range(T):
        r = whole.step_device(acts[t])
        a, b = lo.step_device(acts[t, :N // 2].contiguous()), hi.step_device(acts[t, N // 2:].contiguous())
        assert torch.equal(r.obs[:N // 2], a.obs) and torch.equal(r.obs[N // 2:], b.obs), t
        assert torch.equal(r.reward[:N // 2], a.reward) and torch.equal(r.reward[N // 2:], b.reward), t
        assert torch.equal(r.terminated[N // 2:], b.terminated) and torch.equal(r.truncated[:N // 2], a.truncated), t
        assert bool(torch.isfinite(r.obs).all()) and bool(torch.isfinite(r.reward).all())
        assert torch.equal(r.obs[..., -4:], acts[t])
        if prev is not None:
            keep = ~(r.terminated | r.truncated)          # (a reset keeps the ring too, BaseRLAviary.py:153-154; all rows hold)
            assert torch.equal(r.obs[:, :, 12:12 + 56], prev[:, :, 16:16 + 56]) and bool(keep.any())
        prev = r.obs.clone()
        outs.append((prev, r.reward.clone(), r.terminated.clone(), r.truncated.clone()))
    st = whole.get_state()
    assert float((torch.linalg.vector_norm(st[:, :, 3:7], dim=-1) - 1).abs().max()) <= 3e-7
    obs = torch.empty((T, N, M, 72), device="cuda")
    rew = torch.empty((T, N), device="cuda")
    term = torch.empty((T, N), dtype=torch.bool, device="cuda")
    trunc = torch.empty((T, N), dtype=torch.bool, device="cuda")
    l0 = many.launch_count
    many.step_many(acts, obs, rew, term, trunc)
    assert many.launch_count - l0 == 1
    for t in range(T):
        assert torch.equal(obs[t], outs[t][0]) and torch.equal(rew[t], outs[t][1]), t
        assert torch.equal(term[t], outs[t][2]) and torch.equal(trunc[t], outs[t][3]), t
    for e in (whole, lo, hi, many):
        e.close()
