"""Shared helpers of the test-suite (golden loading, oracle/kernel drivers)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AERO = {"gnd": 1, "drag": 2, "dw": 4}


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith("aero"))


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
