"""Print per-case CUDA-vs-golden errors (exploration aid; the gates live in tests/)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import batch_from_cfg, golden_names, load_golden, rel_err  # noqa: E402


def run(name, precision):
    cfg, g = load_golden(name)
    A = g["actions"]
    f32_actions = A.dtype == np.float32
    adt = torch.float32 if (precision == "fp32" or f32_actions) else torch.float64
    env = batch_from_cfg(cfg, g["init_xyzs"], g["init_rpys"], num_envs=3, precision=precision,
                         action_dtype=adt, keep_ang_vel=True)
    obs0 = env.reset_device().cpu().numpy()
    e0 = np.abs(obs0[1] - g["obs0"]).max()
    worst_state = worst_obs = worst_rew = 0.0
    first_bad = None
    flags_bad = 0
    T = A.shape[0]
    for t in range(T):
        a = torch.as_tensor(A[t]).to("cuda", adt)[None].expand(3, -1, -1).contiguous()
        r = env.step_device(a)
        st = env.get_state().cpu().numpy()[1]
        es = rel_err(st[:, :16], g["states"][t][:, :16])
        eo = rel_err(r.obs.cpu().numpy()[1], g["obs"][t])
        er = rel_err(r.reward.cpu().numpy()[1], g["reward"][t])
        tb = bool(r.terminated[1].item()) != bool(g["terminated"][t]) or bool(r.truncated[1].item()) != bool(g["truncated"][t])
        flags_bad += int(tb)
        worst_state, worst_obs, worst_rew = max(worst_state, es), max(worst_obs, eo), max(worst_rew, er)
        thr = 1e-9 if precision == "fp64" else 1e-3
        if first_bad is None and es > thr:
            first_bad = t
    print(f"{precision} {name:28s} T={T:4d} obs0 {e0:.1e} state {worst_state:.2e} obs {worst_obs:.2e} "
          f"rew {worst_rew:.2e} flag_mismatch {flags_bad} first>thr {first_bad}", flush=True)
    env.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    t0 = time.time()
    for prec in ("fp64", "fp32"):
        for n in golden_names():
            run(n, prec)
    print("elapsed", time.time() - t0)
