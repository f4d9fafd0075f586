"""Generate the golden trajectories under tests/golden/ by RUNNING THE REFERENCE.

Run once in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It imports the unmodified reference env classes through `oracle/ref_harness.py`
(PyBullet / gymnasium replaced by shims; see that file for what is and is not
pinned), steps them with `Physics.DYN` on seeded inputs and stores inputs and
outputs as `.npz`.  The fixtures travel to the GPU box; this script and
/root/reference do not need to.

Every case stores: `cfg` (json), `init_xyzs` (M,3) actually used after reset,
`actions` (T,M,A), `obs` (T,M,D) as returned (cast to fp64 losslessly),
`obs0` reset observation, `reward` (T,), `terminated`, `truncated` (T,) and
`states` (T,M,20) = `_getDroneStateVector` after each step, `rpy_rates` (T,M,3).
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as rh  # noqa: E402


def _make_env(c, task, **kw):
    E = c["enums"]
    cls = {"hover": c["HoverAviary"], "multihover": c["MultiHoverAviary"],
           "spiral": c["SpiralFormationAviary"], "meetup": c["MeetupAviary"], "flock": c["FlockAviary"],
           "leaderfollower": c["LeaderFollowerAviary"]}[task]
    kw = dict(kw)
    kw["drone_model"] = E.DroneModel(kw.pop("drone_model", "cf2x"))
    kw["act"] = E.ActionType(kw.pop("act", "rpm"))
    kw["physics"] = E.Physics.DYN
    kw["obs"] = E.ObservationType.KIN
    with redirect_stdout(io.StringIO()):
        env = cls(**kw)
    return env


def rollout(c, task, actions, np_seed=None, reset_at=(), **kw):
    """`reset_at`: step indices before which `env.reset()` is called again on the SAME env object
    (the controllers of the PID action types are not reset by it, BaseRLAviary.py:73-78)."""
    env = _make_env(c, task, **kw)
    if np_seed is not None:
        np.random.seed(np_seed)
    obs0, _ = env.reset()
    T, M = actions.shape[0], env.NUM_DRONES
    out = dict(init_xyzs=np.array(env.INIT_XYZS, dtype=np.float64),
               init_rpys=np.array(env.INIT_RPYS, dtype=np.float64),
               obs0=np.asarray(obs0, dtype=np.float64), actions=actions,
               obs=[], reward=[], terminated=[], truncated=[], states=[], rpy_rates=[])
    if hasattr(env, "TARGET_POS"):
        out["target_pos"] = np.array(env.TARGET_POS, dtype=np.float64).reshape(-1, 3)
    has_ctrl = hasattr(env, "ctrl")
    if has_ctrl:
        out["ctrl_state"] = []
    if len(reset_at):
        out["reset_at"] = np.array(reset_at, dtype=np.int64)
        out["reset_obs"], out["reset_init_xyzs"] = [], []
    for t in range(T):
        if t in reset_at:
            ro, _ = env.reset()
            out["reset_obs"].append(np.asarray(ro, dtype=np.float64))
            out["reset_init_xyzs"].append(np.array(env.INIT_XYZS, dtype=np.float64))
        o, r, te, tr, _ = env.step(actions[t])
        if has_ctrl:
            out["ctrl_state"].append(np.array([np.concatenate([k.integral_pos_e, k.integral_rpy_e, k.last_rpy])
                                               for k in env.ctrl]))
        out["obs"].append(np.asarray(o, dtype=np.float64))
        out["reward"].append(float(r))
        out["terminated"].append(bool(te))
        out["truncated"].append(bool(tr))
        out["states"].append(np.array([env._getDroneStateVector(i) for i in range(M)]))
        out["rpy_rates"].append(env.rpy_rates.copy())
    for k in ("obs", "reward", "terminated", "truncated", "states", "rpy_rates", "ctrl_state", "reset_obs",
              "reset_init_xyzs"):
        if k in out:
            out[k] = np.array(out[k])
    env.close()
    return out


def save(name, cfg, data):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, cfg=json.dumps(cfg), **data)
    print(f"{name}: T={data['actions'].shape[0]} obs{data['obs'].shape} "
          f"term@{int(np.argmax(data['terminated'])) if data['terminated'].any() else None} "
          f"trunc@{int(np.argmax(data['truncated'])) if data['truncated'].any() else None} "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


def controller_cases(c):
    """ActionType.PID / VEL / ONE_D_PID: DSLPIDControl inside _preprocessAction (BaseRLAviary.py:193-235)."""
    rng = np.random.default_rng
    # Spiral's default action type is VEL (SpiralAviary.py:32), M=3 default, 240/48; fp64 actions
    cfg = dict(task="spiral", drone_model="cf2x", num_drones=3, pyb_freq=240, ctrl_freq=48, act="vel")
    kw = {k: v for k, v in cfg.items() if k != "task"}
    a = rng(21).uniform(-1, 1, (160, 3, 4))
    a[40:44] = 0.0                                   # zero direction: v_unit_vector = 0 branch (:210-213)
    save("spiral3_vel", cfg, rollout(c, "spiral", a, reset_at=(96,), **kw))
    # same with float32 actions (numpy keeps the unit vector and the speed factor in float32)
    a = rng(22).uniform(-1, 1, (64, 3, 4)).astype(np.float32)
    save("spiral3_vel_f32", cfg, rollout(c, "spiral", a, **kw))
    # MultiHover M=2 on CF2P with VEL, 240/30, jittered spawn, two resets on the same env object
    cfg = dict(task="multihover", drone_model="cf2p", num_drones=2, pyb_freq=240, ctrl_freq=30, act="vel",
               initial_xyzs=[[0, 0, .5], [1.2, 0, .6]])
    a = rng(23).uniform(-1, 1, (120, 2, 4))
    save("multihover2_vel_cf2p", cfg, rollout(c, "multihover", a, np_seed=4, reset_at=(50, 90), drone_model="cf2p",
                                              num_drones=2, pyb_freq=240, ctrl_freq=30, act="vel",
                                              initial_xyzs=np.array(cfg["initial_xyzs"])))
    # PID waypoint targets (A=3): near (< 1 m: target itself) and far (> 1 m: unit step) destinations
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=2, pyb_freq=240, ctrl_freq=30, act="pid",
               initial_xyzs=[[0, 0, .5], [1.0, 0, .5]])
    a = np.zeros((150, 2, 3))
    a[:, 0] = [0.3, 0.2, 1.0]
    a[:, 1] = [1.0, -0.4, 0.8]
    a[60:] = rng(24).uniform(-1, 1, (90, 2, 3)) * [2.5, 2.5, 1.0] + [0, 0, 1.2]
    save("multihover2_pid", cfg, rollout(c, "multihover", a, np_seed=6, reset_at=(100,), drone_model="cf2x",
                                         num_drones=2, pyb_freq=240, ctrl_freq=30, act="pid",
                                         initial_xyzs=np.array(cfg["initial_xyzs"])))
    # HoverAviary with ONE_D_PID (A=1), float32 actions
    cfg = dict(task="hover", drone_model="cf2x", pyb_freq=240, ctrl_freq=30, act="one_d_pid")
    a = rng(25).uniform(-1, 1, (120, 1, 1)).astype(np.float32)
    a[:30] = 1.0
    save("hover_one_d_pid", cfg, rollout(c, "hover", a, **{k: v for k, v in cfg.items() if k != "task"}))


def aero_samples(c):
    """Sample the reference's own _groundEffect/_drag/_downwash force formulas."""
    E = c["enums"]
    fake = rh.install()
    with redirect_stdout(io.StringIO()):
        env = c["MultiHoverAviary"](num_drones=3, physics=E.Physics.DYN, pyb_freq=240, ctrl_freq=48,
                                    initial_xyzs=np.array([[0, 0, .05], [.05, .02, .45], [-.3, .1, 1.2]]),
                                    initial_rpys=np.array([[.1, -.2, .3], [0, 0, 0], [.4, .1, -1.]]),
                                    act=E.ActionType.RPM)
    rng = np.random.default_rng(7)
    rows = []
    for trial in range(24):
        M = env.NUM_DRONES
        pos = rng.uniform(-.4, .4, (M, 3))
        pos[:, 2] = rng.uniform(0.005, 1.5, M)
        if trial % 3 == 0:       # stacked pair, small lateral offset: strong downwash
            pos[1, :2] = pos[0, :2] + rng.uniform(-.05, .05, 2)
        rpy = rng.uniform(-.6, .6, (M, 3))
        if trial % 5 == 4:
            rpy[0, 0] = 2.0      # |roll| > pi/2: ground effect switched off
        vel = rng.uniform(-2, 2, (M, 3))
        rpm = env.HOVER_RPM * (1 + 0.05 * rng.uniform(-1, 1, (M, 4)))
        for i in range(M):
            q = fake.getQuaternionFromEuler(rpy[i])
            fake.resetBasePositionAndOrientation(env.DRONE_IDS[i], pos[i], q, physicsClientId=env.CLIENT)
            fake.resetBaseVelocity(env.DRONE_IDS[i], vel[i], [0, 0, 0], physicsClientId=env.CLIENT)
        env._updateAndStoreKinematicInformation()
        for i in range(M):
            fake.force_log.clear()
            env._groundEffect(rpm[i], i)
            gnd = np.zeros(4)
            for (_, _, link, f, _) in fake.force_log:
                gnd[link] = f[2]
            fake.force_log.clear()
            env._drag(rpm[i], i)
            drag_link = np.array(fake.force_log[0][3])
            fake.force_log.clear()
            env._downwash(i)
            dw = sum(f[3][2] for f in fake.force_log)
            rows.append(dict(trial=trial, drone=i, pos=env.pos.copy(), quat=env.quat.copy(),
                             rpy=env.rpy.copy(), vel=env.vel.copy(), rpm=rpm[i].copy(),
                             gnd=gnd, drag_link=drag_link, dw=dw))
    data = {k: np.array([r[k] for r in rows]) for k in rows[0]}
    path = os.path.join(HERE, "aero_formulas.npz")
    np.savez_compressed(path, **data)
    print(f"aero_formulas: {len(rows)} samples, dw nonzero in {(data['dw'] != 0).sum()}")


def main():
    c = rh.reference_classes()
    rng = np.random.default_rng

    # cfg 1: HoverAviary, 1 drone CF2X, 240/30, KIN, RPM; U(-1,1) fp64 actions
    cfg = dict(task="hover", drone_model="cf2x", pyb_freq=240, ctrl_freq=30, act="rpm")
    a = rng(0).uniform(-1, 1, (242, 1, 4))
    save("hover_rpm_uniform", cfg, rollout(c, "hover", a, **{k: v for k, v in cfg.items() if k != "task"}))
    # gentle fp32 actions (policy-like): survives to the 8 s truncation at step 242
    a = (0.02 * rng(10).standard_normal((244, 1, 4))).astype(np.float32)
    save("hover_rpm_gentle_f32", cfg, rollout(c, "hover", a, **{k: v for k, v in cfg.items() if k != "task"}))
    # ONE_D_RPM, a = 0 holds altitude indefinitely; truncation fires on the 242nd step
    cfg1 = dict(cfg, act="one_d_rpm")
    a = np.zeros((244, 1, 1))
    a[5:40] = 0.3
    a[40:80] = -0.25
    save("hover_one_d_rpm", cfg1, rollout(c, "hover", a, **{k: v for k, v in cfg1.items() if k != "task"}))
    # tilted start: exercises rpy extraction / gyroscopic term from step 0
    cfg1b = dict(cfg, initial_xyzs=[[0.2, -0.1, 0.8]], initial_rpys=[[0.3, -0.25, 1.1]])
    a = (0.2 * rng(11).standard_normal((60, 1, 4))).astype(np.float32)
    save("hover_rpm_tilted", cfg1b, rollout(c, "hover", a, pyb_freq=240, ctrl_freq=30, act="rpm",
                                            drone_model="cf2x",
                                            initial_xyzs=np.array(cfg1b["initial_xyzs"]),
                                            initial_rpys=np.array(cfg1b["initial_rpys"])))

    # cfg 2: MultiHover M=2, default spacing + reference jitter (np.random seeded), 240/30, RPM
    cfg2 = dict(task="multihover", drone_model="cf2x", num_drones=2, pyb_freq=240, ctrl_freq=30, act="rpm")
    kw2 = {k: v for k, v in cfg2.items() if k != "task"}
    a = (0.3 * rng(2).standard_normal((128, 2, 4))).astype(np.float32)   # unclipped Gaussian, fp32
    save("multihover2_gauss_f32", cfg2, rollout(c, "multihover", a, np_seed=1, **kw2))
    a = rng(3).uniform(-1, 1, (48, 2, 4))                                  # fp64 uniform
    save("multihover2_uniform", cfg2, rollout(c, "multihover", a, np_seed=5, **kw2))
    a = (0.03 * rng(4).standard_normal((250, 2, 4))).astype(np.float32)  # long: reaches truncation
    a[:, :, :] += 0.02
    save("multihover2_long", cfg2, rollout(c, "multihover", a, np_seed=9, **kw2))
    # M=4 on a 1 m grid (the cfg-4 layout), CF2P airframe
    grid = np.array([[0., 0, .5], [1, 0, .5], [0, 1, .5], [1, 1, .5]])
    cfg4 = dict(task="multihover", drone_model="cf2p", num_drones=4, pyb_freq=240, ctrl_freq=30,
                act="rpm", initial_xyzs=grid.tolist())
    a = (0.25 * rng(6).standard_normal((64, 4, 4))).astype(np.float32)
    save("multihover4_cf2p", cfg4, rollout(c, "multihover", a, np_seed=2, drone_model="cf2p", num_drones=4,
                                           pyb_freq=240, ctrl_freq=30, act="rpm", initial_xyzs=grid))
    # racer airframe, ONE_D_RPM, 240/48
    cfgr = dict(task="multihover", drone_model="racer", num_drones=2, pyb_freq=240, ctrl_freq=48,
                act="one_d_rpm", initial_xyzs=[[0, 0, .4], [1.5, 0, .6]])
    a = (0.5 * rng(8).standard_normal((40, 2, 1))).astype(np.float32)
    save("multihover2_racer_1d", cfgr, rollout(c, "multihover", a, np_seed=3, drone_model="racer",
                                               num_drones=2, pyb_freq=240, ctrl_freq=48, act="one_d_rpm",
                                               initial_xyzs=np.array(cfgr["initial_xyzs"])))

    # cfg 3 (DYN part): Spiral M=5, 240/48, RPM, ring init; covers the 12 s truncation (step 578)
    cfg3 = dict(task="spiral", drone_model="cf2x", num_drones=5, pyb_freq=240, ctrl_freq=48, act="rpm")
    kw3 = {k: v for k, v in cfg3.items() if k != "task"}
    a = (0.2 * rng(3).standard_normal((96, 5, 4))).astype(np.float32)
    save("spiral5_gauss_f32", cfg3, rollout(c, "spiral", a, **kw3))
    a = np.zeros((580, 5, 4), dtype=np.float32)
    a += (0.004 * rng(12).standard_normal((580, 5, 4))).astype(np.float32)
    save("spiral5_long", cfg3, rollout(c, "spiral", a, **kw3))

    controller_cases(c)
    aero_samples(c)


if __name__ == "__main__":
    main()
