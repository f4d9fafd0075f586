#!/usr/bin/env python
"""Throughput bench of the batched drone-step hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one control step of every environment on this GPU = ONE launch of
`step_kernel` (S = pyb_freq/ctrl_freq physics substeps per drone).  The metric is
drone-substeps/s = N_envs * M * S * K / time, aggregated over all ranks.

* default workload: BASELINE.json configs[3] — MultiHoverAviary, M = 4 drones, RPM
  action, 240 Hz physics / 30 Hz control, fp32, auto-reset on, random actions;
  65,536 envs PER GPU (weak scaling: every rank steps its own env shard, no
  data-path collective; NCCL is used only for the barrier / max-time reduction).
* `value`  : inputs resident in HBM (pre-generated action pool, observations
  written into a rotating rollout buffer so the working set exceeds L2).
* `e2e`    : the same step through the host-buffer C-ABI call `bd_step_host`
  (pinned numpy in, pinned numpy out: H2D + kernel + D2H + sync each step).
* `roofline`: algorithmic bytes / measured kernel time vs the
  measured HBM copy peak in MEASURED_PEAKS.json.
* `cpu_baseline` / `--impl reference`: the fp64 numpy oracle of the reference's
  Physics.DYN path (the reference itself needs PyBullet, which is not
  installable), run the way the reference runs it: one Python env object per env,
  SubprocVecEnv-style spawn workers + Pipes on all host cores, auto-reset.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "drone_substeps_per_sec"
UNIT = "drone-substeps/s"
GRID_1M = [[0.0, 0.0, 0.5], [1.0, 0.0, 0.5], [0.0, 1.0, 0.5], [1.0, 1.0, 0.5]]


def grid_xyzs(m: int):
    side = int(np.ceil(np.sqrt(m)))
    return np.array([[float(i % side), float(i // side), 0.5] for i in range(m)])


def algorithmic_bytes_per_drone_step(A: int, B: int, M: int, e: int = 4, extra_obs: int = 0) -> float:
    """read state 13 + action A + history (B-1)A + target 3;
    write state 13 + obs 12 + B*A (+extras); per env 14 B (reward, 2 flags, counter r/w)."""
    return e * (41 + 2 * B * A + extra_obs) + 14.0 / M


# --------------------------------------------------------------------------- CPU arm
def _cpu_worker(remote, n_envs, m, seed):
    """SubprocVecEnv worker stand-in (subproc_vec_env.py:186-261) around the oracle."""
    sys.path.insert(0, ROOT)
    from oracle.aviary_oracle import OracleAviary, step_env_autoreset
    np.random.seed(seed)
    xyz = grid_xyzs(m)
    envs = [OracleAviary(task="multihover", num_drones=m, initial_xyzs=xyz, pyb_freq=240, ctrl_freq=30,
                         act="rpm") for _ in range(n_envs)]
    for e in envs:
        e.reset()
    remote.send("ready")
    while True:
        cmd, data = remote.recv()
        if cmd == "step":
            remote.send([step_env_autoreset(env, a) for env, a in zip(envs, data)])
        elif cmd == "close":
            remote.close()
            break


class CpuVecBaseline:
    """The oracle run like the reference's CPU path: spawn workers, Pipes, auto-reset."""

    def __init__(self, n_workers, envs_per_worker, m):
        ctx = mp.get_context("spawn")
        self.n_workers, self.envs_per_worker, self.m = n_workers, envs_per_worker, m
        self.remotes, self.procs = [], []
        for w in range(n_workers):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(child, envs_per_worker, m, 1000 + w), daemon=True)
            p.start()
            child.close()
            self.remotes.append(parent)
            self.procs.append(p)
        for r in self.remotes:
            assert r.recv() == "ready"
        self.rng = np.random.default_rng(2)

    @property
    def num_envs(self):
        return self.n_workers * self.envs_per_worker

    def step(self):
        acts = self.rng.uniform(-1, 1, (self.num_envs, self.m, 4)).astype(np.float32)
        for r, a in zip(self.remotes, np.array_split(acts, self.n_workers)):
            r.send(("step", a))
        res = [x for r in self.remotes for x in r.recv()]
        obs, rews, dones, infos = zip(*res)
        return np.stack(obs), np.stack(rews), np.stack(dones)

    def close(self):
        for r in self.remotes:
            try:
                r.send(("close", None))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_baseline(steps: int, warmup: int, m: int, envs_per_worker: int = 4):
    cores = host_cores()
    vec = CpuVecBaseline(cores, envs_per_worker, m)
    try:
        for _ in range(warmup):
            vec.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            vec.step()
        dt = time.perf_counter() - t0
    finally:
        vec.close()
    S = 8
    value = vec.num_envs * m * S * steps / dt
    sample = (f"{vec.num_envs} MultiHover envs x {m} drones ({cores} spawn workers x {envs_per_worker} envs, "
              f"Pipe IPC, auto-reset), {steps} control steps after {warmup} warm-up, fp64 numpy oracle")
    return dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample), dt / steps * 1e3


# ----------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary, StepResult

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, M, S, A, Bf = args.envs_per_gpu, args.drones, 8, 4, 15
    env = BatchAviary(task="multihover", num_envs=N, num_drones=M, initial_xyzs=grid_xyzs(M), pyb_freq=240,
                      ctrl_freq=30, act="rpm", precision="fp32", device=dev, auto_reset=True,
                      reset_mode="jitter_philox", seed=1234 + rank)
    D = env.OBS_DIM
    slots = args.rollout_slots
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    act_pool = (torch.rand((slots, N, M, A), generator=gen, device=dev) * 2 - 1).contiguous()
    obs_buf = torch.empty((slots, N, M, D), dtype=torch.float32, device=dev)
    rew_buf = torch.empty((slots, N), dtype=torch.float32, device=dev)
    term_buf = torch.empty((slots, N), dtype=torch.uint8, device=dev)
    trunc_buf = torch.empty((slots, N), dtype=torch.uint8, device=dev)
    outs = [StepResult(obs_buf[i], rew_buf[i], term_buf[i].view(torch.bool), trunc_buf[i].view(torch.bool), None)
            for i in range(slots)]
    env.reset_device(out=obs_buf[0])

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def gpu_steps(n, start):
        for k in range(n):
            i = (start + k) % slots
            env.step_device(act_pool[i], out=outs[i])

    gpu_steps(args.warmup, 0)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    gpu_steps(args.steps, args.warmup)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_count - l0
    if ms < 300:   # keep the GPU busy a little longer so nvidia-smi sees clocks under load
        t_end = time.time() + 0.4
        while time.time() < t_end:
            gpu_steps(50, 0)
            torch.cuda.synchronize(dev)
    clocks = sampler.stop()

    # ---- e2e: host buffers through bd_step_host (H2D + kernel + D2H + sync per step)
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    host_actions = env.pinned_array((min(slots, 4), N, M, A), np.float32)   # this step's inputs wait in pinned host memory
    host_actions[...] = act_pool[:min(slots, 4)].cpu().numpy()
    for k in range(3):
        env.step_host(host_actions[k % host_actions.shape[0]], actions_pinned=True)
    sync_all()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        res = env.step_host(host_actions[k % host_actions.shape[0]], actions_pinned=True)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    assert np.isfinite(res["reward"]).all()
    h2d = N * M * A * 4
    d2h = N * M * D * 4 + N * 4 + 2 * N

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    units = world * N * M * S
    value = units * args.steps / (ms_max * 1e-3)
    e2e_value = units * e2e_steps / (e2e_ms_max * 1e-3)

    bytes_per_drone_step = algorithmic_bytes_per_drone_step(A, Bf, M)
    bytes_per_launch = bytes_per_drone_step * N * M
    kernel_ms = ms / max(launches, 1)
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("envs_per_gpu") == N and tj.get("drones") == M:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass

    line = None
    if rank == 0:
        cpu = None
        if world == 1:
            cpu, _ = run_cpu_baseline(args.cpu_steps, 2, M)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": (f"BASELINE configs[3]: MultiHoverAviary M={M} CF2X Physics.DYN 240Hz/30Hz KIN obs RPM "
                             f"action, {N} envs per GPU x {world} GPU(s), auto-reset, U(-1,1) actions"),
                "envs_per_gpu": N, "drones_per_env": M, "substeps_per_step": S, "obs_dim": D,
                "l2_policy": (f"inputs larger than L2: obs written to a {slots}-slot rotating rollout buffer "
                              f"({slots * N * M * D * 4 / 1e6:.0f} MB) + {slots}-slot action pool; state+history "
                              f"{(N * M * (64 + Bf * A * 4)) / 1e6:.0f} MB"),
                "parallelism": f"env-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "bd::step_kernel_tile<MULTIHOVER,4,4,false>",
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "algorithmic_bytes_per_drone_substep": bytes_per_drone_step / S,
                         "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "BatchAviary.step_host -> bd_step_host (pinned numpy in/out)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    M = args.drones
    cpu, ms_step = run_cpu_baseline(args.steps, args.warmup, M)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", str(args.gpus))), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"BASELINE configs[3]: MultiHoverAviary M={M} CF2X Physics.DYN 240Hz/30Hz KIN obs RPM "
                                f"action; bounded sample on host cores: {cpu['sample']}")},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--drones", type=int, default=4)
    ap.add_argument("--rollout-slots", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-steps", type=int, default=200)
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 24 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        run_reference(args)
    else:
        args.steps = 1000 if args.steps is None else args.steps
        args.warmup = 100 if args.warmup is None else max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    # stdout carries exactly ONE line, the JSON result: anything libraries print on file descriptor 1 meanwhile
    # (e.g. NCCL's version banner) goes to stderr instead
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real_stdout, "w", buffering=1)
    main()
