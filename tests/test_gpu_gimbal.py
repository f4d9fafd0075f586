"""Gimbal and 90-degree branches of the Euler extraction on the GPU.

`pybullet.getEulerFromQuaternion` (call site `BaseAviary.py:518`) has hard branches at |sarg| >= 0.99999
(|pitch| within 4.47e-3 rad of pi/2): roll = 0, pitch = +-pi/2, yaw = 2 atan2(+-x, -+y).  The kernels carry them in
`quat_to_euler` (exact flavour), `quat_to_euler_fast` (float throughput kernel) and in the first line of
`tilt_below_half_pi` (the ground-effect gate, `BaseAviary.py:735`, which also switches at |roll| = pi/2).
These tests START inside / at the edges of those branches and CROSS them, per step against the fp64 oracle.

Stated tolerances: fp64 <= 1e-9 relative on the whole state incl. the Euler angles at every step.
fp32: position / quaternion / velocity <= 2e-4 free-running over the 24-step horizon; Euler angles <= 2e-4 where the
extraction is well conditioned and <= 2e-3 within 0.02 rad of the singularity (roll and yaw are individually
ill-conditioned there: their error is the float32 rounding of the quaternion, 6e-8, divided by cos(pitch)); the branch
taken must be the oracle's whenever the oracle's |sarg| is further than 3e-6 from 0.99999.
"""
import numpy as np
import pytest
import torch

from _util import oracle_inject, rel_err
from oracle import bullet_math as bm
from oracle.aviary_oracle import OracleAviary

pytestmark = pytest.mark.gpu

HALF_PI = 0.5 * np.pi
BAND = 0.99999


def _quat(rpy):
    return np.array(bm.pose_roundtrip(bm.quaternion_from_euler(rpy)))


def _pair(task, M, N, precision, physics="dyn", aero=0, ctrl_freq=30, xyz=None, rpy=None):
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    if xyz is None:
        xyz = np.array([[float(i), 0.0, 1.0] for i in range(M)])
    env = BatchAviary(task=task, num_envs=N, num_drones=M, initial_xyzs=xyz, initial_rpys=rpy, pyb_freq=240,
                      ctrl_freq=ctrl_freq, act="rpm", physics=physics, precision=precision, auto_reset=False,
                      reset_mode="fixed", action_dtype=torch.float32, keep_ang_vel=(precision == "fp64"))
    oracles = [OracleAviary(task=task, num_drones=M, initial_xyzs=xyz, initial_rpys=rpy, pyb_freq=240, ctrl_freq=ctrl_freq,
                            act="rpm", aero=aero) for _ in range(N)]
    for o in oracles:
        o.reset(fixed=True)
    return env, oracles


def _inject(env, oracles, pos, rpy, vel, rates):
    """pos/rpy/vel/rates: (N,M,3).  Same canonical unit quaternion on both sides."""
    N, M = pos.shape[:2]
    quat = np.array([[_quat(rpy[e, i]) for i in range(M)] for e in range(N)])
    kin = np.concatenate([pos, quat, vel, rates], axis=-1)
    env.set_state(torch.as_tensor(kin))
    for e, o in enumerate(oracles):
        st = np.zeros((M, 20))
        st[:, 0:3], st[:, 3:7], st[:, 10:13] = pos[e], quat[e], vel[e]
        oracle_inject(o, st, rates[e])
    return quat


def _check(env, oracles, r, precision, t, flags=True):
    st = env.get_state().cpu().numpy()
    obs, rew = r.obs.cpu().numpy(), r.reward.cpu().numpy()
    for e, o in enumerate(oracles):
        ost = np.array([o.state_vector(i) for i in range(o.NUM_DRONES)])
        if precision == "fp64":
            assert rel_err(st[e][:, :16], ost[:, :16]) <= 1e-9, (t, e, st[e][:, 7:10], ost[:, 7:10])
            assert rel_err(obs[e][:, :12], o._last_obs[:, :12]) <= 2.5e-7, (t, e)
            assert rel_err(rew[e], o._last_rew) <= 1e-9, (t, e)
        else:
            assert rel_err(st[e][:, 0:7], ost[:, 0:7]) <= 2e-4, (t, e)
            assert rel_err(st[e][:, 10:13], ost[:, 10:13]) <= 2e-4, (t, e)
            for i in range(o.NUM_DRONES):
                x, y, z, w = ost[i, 3:7]
                sarg = -2 * (x * z - w * y)
                if abs(abs(sarg) - BAND) < 3e-6:
                    continue          # the branch decision itself is within float32 rounding
                near = abs(sarg) > np.cos(0.02)
                # angles modulo 2 pi (a roll of +pi and -pi are the same attitude)
                ang = obs[e][i, 3:6].astype(np.float64)      # the step kernel's own extraction (quat_to_euler_fast)
                d = np.abs((ang - ost[i, 7:10] + np.pi) % (2 * np.pi) - np.pi)
                assert d.max() <= (2e-3 if near else 2e-4), (t, e, i, ang, ost[i, 7:10], sarg)
                in_band = abs(sarg) >= BAND
                assert (ang[0] == 0.0 and abs(abs(ang[1]) - HALF_PI) < 1e-6) == in_band, (t, e, i, sarg)
                d2 = np.abs((st[e][i, 7:10] - ost[i, 7:10] + np.pi) % (2 * np.pi) - np.pi)      # bd_get_state's (exact flavour)
                assert d2.max() <= (2e-3 if near else 2e-4), (t, e, i)
        if flags:
            assert bool(r.terminated[e]) == bool(o._last_te) and bool(r.truncated[e]) == bool(o._last_tr), (t, e)


def _step_all(env, oracles, a):
    r = env.step_device(torch.as_tensor(a, device="cuda"))
    for e, o in enumerate(oracles):
        o._last_obs, o._last_rew, o._last_te, o._last_tr, _ = o.step(a[e])
        o._last_obs = np.asarray(o._last_obs, dtype=np.float64)
    return r


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_euler_branches_from_static_attitudes(precision):
    """Attitudes at, inside, at the edge of and just outside both gimbal bands, beyond 90 degrees of pitch, and at /
    around roll = +-pi/2; zero body rates and the hover action, so the attitude is exactly what was injected."""
    deltas = [0.0, 1e-4, 1e-3, 4e-3, 4.4e-3, 4.6e-3, 5e-3, 1e-2, 3e-2]
    rpys = []
    for sgn in (+1, -1):
        for k, dlt in enumerate(deltas):
            rpys.append([0.3 - 0.1 * k, sgn * (HALF_PI - dlt), -0.4 + 0.15 * k])
        rpys.append([0.25, sgn * (HALF_PI + 0.2), 0.1])            # beyond 90 degrees: same attitude, other Euler triple
        for dlt in (0.0, 1e-6, -1e-6, 1e-3, -1e-3):
            rpys.append([sgn * (HALF_PI + dlt), 0.2, -0.3])            # roll at / around +-pi/2
    rpys.append([0.0, HALF_PI, 0.0])
    rpys.append([0.0, -HALF_PI, 0.0])
    M = 2
    if len(rpys) % M:
        rpys.append([0.1, 0.2, 0.3])
    N = len(rpys) // M
    rpy = np.array(rpys).reshape(N, M, 3)
    env, oracles = _pair("multihover", M, N, precision)
    pos = np.tile(np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 1.2]]), (N, 1, 1))
    zeros = np.zeros((N, M, 3))
    quat = _inject(env, oracles, pos, rpy, zeros, zeros)
    n_band = int(np.sum(np.abs(-2 * (quat[..., 0] * quat[..., 2] - quat[..., 3] * quat[..., 1])) >= BAND))
    assert n_band >= 10                      # both gimbal branches are really entered
    a = np.zeros((N, M, 4), dtype=np.float32)
    for t in range(2):
        r = _step_all(env, oracles, a)
        _check(env, oracles, r, precision, t)
    env.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("sgn", [+1, -1])
def test_pitch_crosses_half_pi_through_the_gimbal_band(precision, sgn):
    """Pitch drifts at ~1e-3 rad per control step from 8e-3 below +-pi/2 to 16e-3 beyond it: enters the band, spends
    ~9 control steps inside it (roll = 0, pitch = +-pi/2 exactly), leaves on the far side (roll and yaw jump by pi)."""
    M, N = 2, 3
    env, oracles = _pair("multihover", M, N, precision)
    rpy = np.zeros((N, M, 3))
    rates = np.zeros((N, M, 3))
    for e in range(N):
        for i in range(M):
            roll = 0.3 - 0.25 * e + 0.1 * i
            rpy[e, i] = [roll, sgn * (HALF_PI - 8e-3 - 2e-4 * i), -0.4 + 0.3 * e]
            rates[e, i] = [0.0, sgn * 0.03 / np.cos(roll), 0.0]       # pitch rate = wy cos(roll)
    pos = np.tile(np.array([[0.0, 0.0, 2.0], [1.0, 0.0, 2.2]]), (N, 1, 1))
    _inject(env, oracles, pos, rpy, np.zeros((N, M, 3)), rates)
    a = np.zeros((N, M, 4), dtype=np.float32)
    seen_in, seen_far = 0, 0
    for t in range(24):
        r = _step_all(env, oracles, a)
        _check(env, oracles, r, precision, t, flags=True)
        for o in oracles:
            seen_in += int(np.sum((o.rpy[:, 0] == 0.0) & (np.abs(np.abs(o.rpy[:, 1]) - HALF_PI) < 1e-12)))
            seen_far += int(np.sum(np.abs(o.rpy[:, 0]) > 2.0))
    assert seen_in >= 20 and seen_far >= 10
    env.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_ground_effect_gate_crossing_roll_and_pitch_half_pi(precision, off=6.1e-3):
    """Ground effect (`BaseAviary.py:735-742`) close to the ground while roll crosses +-pi/2 and pitch crosses the
    gimbal band: the gate (evaluated every substep) must switch exactly where the oracle's does.

    Not asserted, on purpose: off = 6e-3 makes the roll of drone 1 reach pi/2 to within 5e-16 after exactly 30 substeps
    (2e-4 rad each).  There the reference's own decision `abs(atan2(ys, xs)) < pi/2` hinges on the last two bits of
    xs = w^2 - x^2 - y^2 + z^2 ~ 5.6e-16; the kernel's xs differs at that level (fused multiply-adds, 30 substeps of
    rounding history) and the gate may switch one substep apart, a 7e-3 relative velocity difference.  Any re-ordering
    of the reference's arithmetic has the same property; measured in round 2 (`profiles/README.md`)."""
    M, N = 4, 2
    xyz = np.array([[0.0, 0.0, 0.06], [0.6, 0.0, 0.05], [0.0, 0.6, 0.07], [0.6, 0.6, 0.05]])
    env, oracles = _pair("multihover", M, N, precision, physics="dyn_gnd", aero=1, ctrl_freq=48, xyz=xyz)
    rpy = np.zeros((N, M, 3))
    rates = np.zeros((N, M, 3))
    for e in range(N):
        s = 1.0 if e == 0 else -1.0
        rpy[e, 0] = [s * (HALF_PI - off), 0.2, 0.1];            rates[e, 0] = [s * 0.048, 0.0, 0.0]      # roll up through pi/2
        rpy[e, 1] = [s * (HALF_PI + off), -0.1, 0.4];           rates[e, 1] = [-s * 0.048, 0.0, 0.0]     # roll down through pi/2
        rpy[e, 2] = [0.0, s * (HALF_PI - 8e-3), 0.3];            rates[e, 2] = [0.0, s * 0.048, 0.0]      # pitch through the band
        rpy[e, 3] = [0.0, s * HALF_PI, -0.2];                    rates[e, 3] = [0.0, 0.0, 0.0]            # exactly at the pole
    pos = np.tile(xyz, (N, 1, 1))
    _inject(env, oracles, pos, rpy, np.zeros((N, M, 3)), rates)
    rng = np.random.default_rng(5)
    for t in range(24):
        a = np.zeros((N, M, 4), dtype=np.float32)
        a[:] = (0.05 * rng.standard_normal((N, M, 1))).astype(np.float32)     # collective thrust only: no torque
        r = _step_all(env, oracles, a)
        _check(env, oracles, r, precision, t, flags=True)
    env.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_reset_at_the_pole(precision):
    """INIT_RPYS with pitch = +-pi/2 exactly: the reset observation goes through the gimbal branch (reset kernel)."""
    M, N = 2, 2
    rpy = np.array([[0.3, HALF_PI, -0.2], [0.1, -HALF_PI, 0.5]])
    env, oracles = _pair("multihover", M, N, precision, rpy=rpy)
    obs0 = env.reset_device().cpu().numpy()
    oo, _ = oracles[0].reset(fixed=True)
    assert rel_err(obs0[1], np.asarray(oo, dtype=np.float64)) <= 2.5e-7
    assert obs0[0, 0, 3] == 0.0 and abs(obs0[0, 0, 4] - HALF_PI) < 1e-6 and abs(obs0[0, 1, 4] + HALF_PI) < 1e-6
    env.close()
