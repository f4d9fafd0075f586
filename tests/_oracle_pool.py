"""Run many independent `OracleAviary` trajectories on all host cores (test infrastructure).

BASELINE.json configs[1] (176 envs x 256 steps) and configs[2] (64 envs x 578 steps, all aero terms) take
1-3 minutes of single-core Python; spread over a spawn pool they finish in seconds on the GPU box.
One job = one environment driven like a `SubprocVecEnv` worker drives it
(`subproc_vec_env.py:186-207`): optional reset-on-done with injected jitter draws.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def accepted_jitter(rng, orig_xyz):
    """One `np.random.uniform(-0.25, 0.25, (M,3))` draw that passes MultiHoverAviary.reset's
    rejection rule (`MultiHoverAviary.py:83-102`): z clipped to [0.1, 1], every pair >= 0.5 m apart."""
    M = orig_xyz.shape[0]
    while True:
        j = rng.uniform(-0.25, 0.25, (M, 3))
        cand = orig_xyz + j
        cand[:, 2] = np.clip(cand[:, 2], 0.1, 1.0)
        d = np.linalg.norm(cand[:, None, :] - cand[None, :, :], axis=2)
        np.fill_diagonal(d, np.inf)
        if not np.any(d < 0.5):
            return j


def _run_one(job):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle.aviary_oracle import OracleAviary, step_env_autoreset
    kw = dict(job["kw"])
    env = OracleAviary(**kw)
    actions = job["actions"]                     # (T, M, A)
    jit = job.get("jitter")                      # (T+1, M, 3) accepted draws: [0] for the first reset, [t+1] for step t
    auto = bool(job.get("auto_reset", False))
    if jit is not None:
        obs0, _ = env.reset(jitter=[jit[0]])
    else:
        obs0, _ = env.reset(fixed=True)
    T, M = actions.shape[0], env.NUM_DRONES
    out = dict(obs0=np.asarray(obs0, dtype=np.float32), obs=[], reward=[], terminated=[], truncated=[], states=[],
               rates=[], stepc=[], targets=[], term_obs=[])
    for t in range(T):
        if auto:
            # the flags of the step itself (before the reset) are what the kernel reports
            o, r, te, tr, info = env.step(actions[t])
            tob = None
            if te or tr:
                tob = np.array(o, copy=True)
                o, _ = env.reset(jitter=[jit[t + 1]]) if jit is not None else env.reset(fixed=(env.task == "multihover"))
        else:
            o, r, te, tr, info = env.step(actions[t])
            tob = None
        out["obs"].append(np.asarray(o, dtype=np.float32))
        out["term_obs"].append(np.asarray(tob if tob is not None else np.zeros_like(o), dtype=np.float32))
        out["reward"].append(float(r))
        out["terminated"].append(bool(te))
        out["truncated"].append(bool(tr))
        out["states"].append(np.array([env.state_vector(i) for i in range(M)]))
        out["rates"].append(env.rpy_rates.copy())
        out["stepc"].append(int(env.step_counter))
        out["targets"].append(np.array(env.TARGET_POS, dtype=np.float64).reshape(-1, 3).copy()
                              if hasattr(env, "TARGET_POS") else np.zeros((M, 3)))
    return {k: np.array(v) for k, v in out.items()}


def run_oracles(jobs, processes=None):
    """-> dict of arrays stacked over envs on axis 1: obs (T,N,M,D), states (T,N,M,20), ..."""
    if processes is None:
        try:
            processes = len(os.sched_getaffinity(0))
        except Exception:
            processes = os.cpu_count() or 1
    processes = max(1, min(processes, len(jobs)))
    if processes == 1:
        res = [_run_one(j) for j in jobs]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Pool(processes) as pool:
            res = pool.map(_run_one, jobs, chunksize=max(1, len(jobs) // (4 * processes)))
    out = {}
    for k in res[0]:
        out[k] = np.stack([r[k] for r in res], axis=0 if k == "obs0" else 1)
    return out
