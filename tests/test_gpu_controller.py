"""GPU parity of ActionType.PID / VEL / ONE_D_PID: the DSL PID controller runs inside the step
kernel (BaseRLAviary.py:193-235, control/DSLPIDControl.py:82-246).

The reference's attitude loop saturates its torque clip and chatters at 30-48 Hz control rates,
so closed-loop trajectories amplify rounding differences by ~3x per control step (see
tests/test_oracle_golden.py): free-running comparisons are made over the horizon where the
reference itself is still reproducible, and the WHOLE golden trajectory is checked step by step
from the reference's own previous state (kinematics + controller memory injected).

Tolerances (|a-b| / max(|b|, floor)):
  fp64, float64 actions : 1e-9 per step from an identical state (rpm: floor 1e4), 1e-9 free-running
                          over the first 16 steps (whole trajectory for the two non-chattering cases)
  fp64, float32 actions : same; numpy's float32 VEL arithmetic is reproduced operation by operation
  fp32                  : 2e-4 on the commanded rpm (the attitude loop multiplies fp32 rounding of
                          rpy by D/dt = 6e5..9.6e5 PWM per rad), 1e-4 on the state after one step
"""
import numpy as np
import pytest
import torch

from _util import batch_from_cfg, golden_names, load_golden, oracle_from_cfg, oracle_inject, rel_err

pytestmark = pytest.mark.gpu

STABLE = ("hover_one_d_pid", "multihover2_vel_cf2p")


def _adt(A, precision):
    return torch.float32 if (precision == "fp32" or A.dtype == np.float32) else torch.float64


def _dev(a, dtype, n=1):
    return torch.as_tensor(a).to("cuda", dtype)[None].expand(n, -1, -1).contiguous()


def _kin13(states, rates):
    """golden `states` (M,20) + body rates (M,3) -> [pos3 quat4 vel3 rates3]."""
    return np.concatenate([states[:, 0:7], states[:, 10:13], rates], axis=1)


def _inject(env, g, t, N):
    """Put every env of the batch into the reference's state after golden step t."""
    k = torch.as_tensor(_kin13(g["states"][t], g["rpy_rates"][t]))[None].expand(N, -1, -1)
    env.set_state(kin13=k, step_counter=torch.full((N,), 8, dtype=torch.int32))
    env.set_controller_state(torch.as_tensor(g["ctrl_state"][t])[None].expand(N, -1, -1))


@pytest.mark.parametrize("name", golden_names(controller=True))
def test_fp64_free_run_matches_reference(name):
    cfg, g = load_golden(name)
    A = g["actions"]
    adt = _adt(A, "fp64")
    N = 3
    env = batch_from_cfg(cfg, g["init_xyzs"], g["init_rpys"], num_envs=N, precision="fp64", action_dtype=adt)
    assert env.ACTION_DIM == A.shape[2] and env.OBS_DIM == g["obs"].shape[2]
    obs0 = env.reset_device().cpu().numpy()
    assert rel_err(obs0[N - 1], g["obs0"]) <= 2.5e-7
    resets = [int(t) for t in g["reset_at"]] if "reset_at" in g.files else []
    horizon = A.shape[0] if name in STABLE else 16
    for t in range(horizon):
        if t in resets:      # env.reset() on the same object: controller memory is kept (BaseRLAviary.py:73-78)
            env.set_initial_poses(g["reset_init_xyzs"][resets.index(t)], g["init_rpys"])
            o = env.reset_device().cpu().numpy()
            assert rel_err(o[0], g["reset_obs"][resets.index(t)]) <= 2.5e-7, (name, t)
        r = env.step_device(_dev(A[t], adt, N))
        st, rates = env.get_state(with_rates=True)
        st, rates = st.cpu().numpy(), rates.cpu().numpy()
        cs = env.get_controller_state().cpu().numpy()
        for e in (0, N - 1):
            assert rel_err(st[e][:, :13], g["states"][t][:, :13]) <= 1e-9, (name, t)
            assert rel_err(rates[e], g["rpy_rates"][t]) <= 1e-9, (name, t)
            assert rel_err(st[e][:, 16:20], g["states"][t][:, 16:20], floor=1e4) <= 1e-9, (name, t)
            assert rel_err(cs[e], g["ctrl_state"][t]) <= 1e-9, (name, t)
        assert rel_err(r.obs.cpu().numpy()[N - 1], g["obs"][t]) <= 2.5e-7, (name, t)
        assert rel_err(r.reward.cpu().numpy(), np.full(N, g["reward"][t])) <= 1e-9, (name, t)
        assert r.terminated.cpu().numpy().tolist() == [bool(g["terminated"][t])] * N, (name, t)
        assert r.truncated.cpu().numpy().tolist() == [bool(g["truncated"][t])] * N, (name, t)
    env.close()


@pytest.mark.parametrize("name", golden_names(controller=True))
@pytest.mark.parametrize("precision,tol_state,tol_rpm", [("fp64", 1e-9, 1e-9), ("fp32", 1e-4, 2e-4)])
def test_every_step_from_the_reference_state(name, precision, tol_state, tol_rpm):
    cfg, g = load_golden(name)
    A = g["actions"]
    adt = _adt(A, precision)
    N = 2
    env = batch_from_cfg(cfg, g["init_xyzs"], g["init_rpys"], num_envs=N, precision=precision, action_dtype=adt)
    env.reset_device()
    resets = [int(t) for t in g["reset_at"]] if "reset_at" in g.files else []
    worst = 0.0
    for t in range(A.shape[0]):
        if t in resets:
            continue
        if t > 0:
            _inject(env, g, t - 1, N)
        env.step_device(_dev(A[t], adt, N))
        st, rates = env.get_state(with_rates=True)
        st, rates = st.cpu().numpy()[N - 1], rates.cpu().numpy()[N - 1]
        cs = env.get_controller_state().cpu().numpy()[N - 1]
        e_rpm = rel_err(st[:, 16:20], g["states"][t][:, 16:20], floor=1e4)
        e_state = max(rel_err(st[:, :13], g["states"][t][:, :13]), rel_err(rates, g["rpy_rates"][t]),
                      rel_err(cs, g["ctrl_state"][t]))
        assert e_rpm <= tol_rpm and e_state <= tol_state, (name, t, e_rpm, e_state)
        worst = max(worst, e_rpm, e_state)
    env.close()


def test_pid_actions_batch_matches_oracle_many_envs():
    """64 envs x 3 drones with different VEL actions and per-env spawn poses against 64 oracles
    (ragged tile: 192 drones = 1.5 CTAs)."""
    from oracle.aviary_oracle import OracleAviary
    rng = np.random.default_rng(5)
    N, M, T = 64, 3, 6
    xyz = rng.uniform(-0.5, 0.5, (N, M, 3)) + np.array([[0, 0, 1.0], [1.5, 0, 1.0], [0, 1.5, 1.0]])
    rpy = rng.uniform(-0.2, 0.2, (N, M, 3))
    from marl_gym_pybullet_drones_b200 import BatchAviary
    env = BatchAviary(task="spiral", num_envs=N, num_drones=M, pyb_freq=240, ctrl_freq=48, act="vel",
                      precision="fp64", auto_reset=False, action_dtype=torch.float64)
    env.set_initial_poses(xyz, rpy)
    env.reset_device()
    oracles = []
    for e in range(N):
        o = OracleAviary(task="spiral", num_drones=M, pyb_freq=240, ctrl_freq=48, act="vel",
                         initial_xyzs=xyz[e], initial_rpys=rpy[e])
        o.reset()
        oracles.append(o)
    for t in range(T):
        a = rng.uniform(-1, 1, (N, M, 4))
        r = env.step_device(torch.as_tensor(a).cuda())
        obs = r.obs.cpu().numpy()
        rew = r.reward.cpu().numpy()
        for e in range(N):
            oo, rr, te, tr, _ = oracles[e].step(a[e])
            assert rel_err(obs[e], oo) <= 1e-6, (t, e)
            assert abs(rew[e] - rr) <= 1e-9 * max(1.0, abs(rr)), (t, e)
    st = env.get_state().cpu().numpy()
    for e in range(N):
        ref = np.array([oracles[e].state_vector(i) for i in range(M)])
        assert rel_err(st[e][:, :13], ref[:, :13]) <= 1e-9
        assert rel_err(st[e][:, 16:20], ref[:, 16:20], floor=1e4) <= 1e-9
    env.close()


def test_controller_memory_survives_autoreset_unless_asked():
    """SubprocVecEnv keeps the env objects, and env.reset() never touches `self.ctrl`
    (BaseRLAviary.py:73-78): integrators carry over episode boundaries.  `reset_controllers=True`
    is this repo's opt-in alternative."""
    from marl_gym_pybullet_drones_b200 import BatchAviary
    xyz = np.array([[0.0, 0.0, 0.12], [1.0, 0.0, 0.12]])
    for flag in (False, True):
        env = BatchAviary(task="multihover", num_envs=4, num_drones=2, initial_xyzs=xyz, act="pid",
                          precision="fp64", reset_mode="fixed", auto_reset=True, reset_controllers=flag)
        env.reset_device()
        a = torch.zeros(4, 2, 3, dtype=torch.float64, device="cuda")
        a[..., 2] = -5.0                     # waypoint below the floor: the drones descend and terminate
        done_at = None
        for t in range(120):
            r = env.step_device(a)
            if bool(r.terminated.any() | r.truncated.any()):
                done_at = t
                break
        assert done_at is not None
        cs = env.get_controller_state().cpu().numpy()
        sc = env.get_state(with_step_counter=True)[1].cpu().numpy()
        assert (sc == 0).all()               # every env was re-spawned in this step
        if flag:
            assert np.all(cs == 0.0)
        else:
            assert np.abs(cs[:, :, 2]).max() > 1e-3      # integral z error kept
        env.close()


def test_vel_with_drag_uses_commanded_rpm_history():
    """DYN + drag: the drag model reads last_clipped_action = the controller's previous rpm
    (BaseAviary.py:372,773), zero right after a reset."""
    from oracle.aviary_oracle import AERO_DRAG, OracleAviary
    from marl_gym_pybullet_drones_b200 import BatchAviary
    xyz = np.array([[0.0, 0.0, 1.0], [0.8, 0.3, 1.2]])
    env = BatchAviary(task="multihover", num_envs=2, num_drones=2, initial_xyzs=xyz, act="vel", physics="dyn_drag",
                      precision="fp64", reset_mode="fixed", auto_reset=False, action_dtype=torch.float64)
    env.reset_device()
    o = OracleAviary(task="multihover", num_drones=2, initial_xyzs=xyz, act="vel", aero=AERO_DRAG)
    o.reset(fixed=True)
    rng = np.random.default_rng(9)
    for t in range(10):
        a = rng.uniform(-1, 1, (2, 4))
        env.step_device(_dev(a, torch.float64, 2))
        o.step(a)
        st = env.get_state().cpu().numpy()[1]
        ref = np.array([o.state_vector(i) for i in range(2)])
        assert rel_err(st[:, :13], ref[:, :13]) <= 1e-9, t
        assert rel_err(st[:, 16:20], ref[:, 16:20], floor=1e4) <= 1e-9, t
    env.close()


def test_controller_api_errors():
    from marl_gym_pybullet_drones_b200 import BatchAviary
    from marl_gym_pybullet_drones_b200._native import NativeError
    with pytest.raises(ValueError, match="no controller is available"):
        BatchAviary(task="multihover", num_drones=2, drone_model="racer", act="vel")
    env = BatchAviary(task="multihover", num_envs=2, num_drones=2, act="rpm")
    with pytest.raises(NativeError):
        env.get_controller_state()
    env.close()
    env = BatchAviary(task="hover", num_envs=2, act="one_d_pid")
    assert env.ACTION_DIM == 1 and env.OBS_DIM == 12 + 15
    env.set_controller_state(None)
    assert float(env.get_controller_state().abs().max()) == 0.0
    env.close()


def test_spiral_view_defaults_to_vel_like_the_reference():
    """SpiralAviary.py:32: act defaults to ActionType.VEL; the Gymnasium view reproduces the reference."""
    from marl_gym_pybullet_drones_b200 import ActionType, SpiralFormationAviary
    cfg, g = load_golden("spiral3_vel")
    env = SpiralFormationAviary()
    assert env.ACT_TYPE == ActionType.VEL and env.NUM_DRONES == 3 and env.observation_space.shape == (3, 119)
    obs, _ = env.reset()
    assert rel_err(obs, g["obs0"]) <= 2.5e-7
    for t in range(12):
        o, r, te, tr, _ = env.step(g["actions"][t])
        assert rel_err(o, g["obs"][t]) <= 1e-6 and abs(r - g["reward"][t]) <= 1e-8
    env.close()
