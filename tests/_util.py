"""Shared helpers of the test-suite (golden loading, oracle/kernel drivers)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AERO = {"gnd": 1, "drag": 2, "dw": 4}


TRAINER_FIXTURES = ("normalizers", "checkpoint_layout")   # not trajectories (test_trainer_golden.py)
CONTROLLER_CASES = ("spiral3_vel", "spiral3_vel_f32", "multihover2_vel_cf2p", "multihover2_pid", "hover_one_d_pid")


def golden_names(controller=False):
    """Golden trajectories; `controller=True` selects the PID / VEL / ONE_D_PID cases instead."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                   if not os.path.basename(p).startswith("aero") and os.path.basename(p)[:-4] not in TRAINER_FIXTURES)
    return [n for n in names if (n in CONTROLLER_CASES) == controller]


def oracle_inject(env, states, rpy_rates, ctrl_state=None):
    """Put an OracleAviary into the state of a golden row: states (M,20), body rates (M,3) — the
    dynamics integrate `rpy_rates`, the state vector only carries R*rates (BaseAviary.py:873) —
    and ctrl_state (M,9)."""
    for i in range(env.NUM_DRONES):
        env._store_pos[i] = tuple(states[i, 0:3])
        env._store_quat[i] = tuple(states[i, 3:7])
        env._store_vel[i] = tuple(states[i, 10:13])
        env._store_angv[i] = tuple(states[i, 13:16])
        if ctrl_state is not None:
            env.ctrl[i].integral_pos_e = ctrl_state[i, 0:3].copy()
            env.ctrl[i].integral_rpy_e = ctrl_state[i, 3:6].copy()
            env.ctrl[i].last_rpy = ctrl_state[i, 6:9].copy()
    env.last_clipped_action = states[:, 16:20].copy()
    env.rpy_rates = np.array(rpy_rates, dtype=np.float64).copy()
    env._refresh_kinematics()


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = json.loads(str(g["cfg"]))
    return cfg, g


def oracle_from_cfg(cfg, init_xyzs, init_rpys, aero=0, integrator="quat"):
    """OracleAviary configured like a golden case, positions injected (no jitter)."""
    from oracle.aviary_oracle import OracleAviary
    c = {k: v for k, v in cfg.items() if k not in ("initial_xyzs", "initial_rpys")}
    task = c.pop("task")
    env = OracleAviary(task=task, initial_xyzs=init_xyzs, initial_rpys=init_rpys, aero=aero,
                       integrator=integrator, **c)
    env.reset(fixed=True)
    return env


def batch_from_cfg(cfg, init_xyzs, init_rpys, num_envs=1, precision="fp64", physics="dyn",
                   action_dtype=None, auto_reset=False, reset_mode="fixed", **kw):
    """BatchAviary configured like a golden case (GPU)."""
    import torch
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    c = {k: v for k, v in cfg.items() if k not in ("initial_xyzs", "initial_rpys")}
    if action_dtype is None:
        action_dtype = torch.float64 if precision == "fp64" else torch.float32
    return BatchAviary(task=c["task"], num_envs=num_envs, drone_model=c.get("drone_model", "cf2x"),
                       num_drones=c.get("num_drones", 1), initial_xyzs=init_xyzs, initial_rpys=init_rpys,
                       physics=physics, pyb_freq=c["pyb_freq"], ctrl_freq=c["ctrl_freq"], act=c["act"],
                       precision=precision, auto_reset=auto_reset, reset_mode=reset_mode,
                       action_dtype=action_dtype, **kw)


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
