"""Per-launch CUDA-event timing of the step kernel (steady state, back-to-back launches)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary, StepResult  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--drones", type=int, default=4)
ap.add_argument("--slots", type=int, default=16)
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--task", default="multihover")
ap.add_argument("--physics", default="dyn")
ap.add_argument("--ctrl-freq", type=int, default=30)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--act", default="rpm")
args = ap.parse_args()
N, M = args.envs, args.drones
side = int(np.ceil(np.sqrt(M)))
xyz = np.array([[float(i % side), float(i // side), 0.5] for i in range(M)])
env = BatchAviary(task=args.task, num_envs=N, num_drones=M, initial_xyzs=None if args.task == "spiral" else xyz,
                  physics=args.physics, ctrl_freq=args.ctrl_freq, precision=args.precision, auto_reset=True, seed=1,
                  act=args.act)
S = args.slots
acts = (torch.rand((S, N, M, env.ACTION_DIM), device="cuda") * 2 - 1).to(env.action_dtype)
obs = torch.empty((S, N, M, env.OBS_DIM), device="cuda")
rew = torch.empty((S, N), device="cuda", dtype=env.real_dtype)
te = torch.empty((S, N), dtype=torch.uint8, device="cuda")
tr = torch.empty((S, N), dtype=torch.uint8, device="cuda")
outs = [StepResult(obs[i], rew[i], te[i].view(torch.bool), tr[i].view(torch.bool), None) for i in range(S)]
env.reset_device(out=obs[0])
for k in range(50):
    env.step_device(acts[k % S], out=outs[k % S])
torch.cuda.synchronize()
# (a) events around every launch
evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.n + 1)]
evs[0].record()
for k in range(args.n):
    env.step_device(acts[k % S], out=outs[k % S])
    evs[k + 1].record()
torch.cuda.synchronize()
d = np.array([evs[k].elapsed_time(evs[k + 1]) * 1e3 for k in range(args.n)])
print(f"per-launch (event to event) us: median {np.median(d):.1f} min {d.min():.1f} p90 {np.percentile(d, 90):.1f}")
# (b) whole loop, no events inside
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(args.n):
    env.step_device(acts[k % S], out=outs[k % S])
e1.record()
torch.cuda.synchronize()
print(f"loop: {e0.elapsed_time(e1) * 1e3 / args.n:.1f} us/step")
# (c) CUDA graph of 16 steps
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for k in range(3):
        env.step_device(acts[k % S], out=outs[k % S])
    with torch.cuda.graph(g, stream=s):
        for k in range(S):
            env.step_device(acts[k], out=outs[k])
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
e0.record()
reps = max(1, args.n // S)
for _ in range(reps):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph: {e0.elapsed_time(e1) * 1e3 / (reps * S):.1f} us/step")
# (d) host-side cost of one step_device call
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(args.n):
    env.step_device(acts[k % S], out=outs[k % S])
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host call cost (async enqueue): {(t1 - t0) * 1e6 / args.n:.1f} us/step")
