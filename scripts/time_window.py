"""Where do short timed windows lose time?  bench.py's device-timed region is `--steps K` launches between two
synchronisations (the driver runs K = 20 after 5 warm-up steps); in steady state the launches overlap tile by tile and
cost 32.6 us each at the headline size, the 20-step window costs 36.8.  This script repeats the window and varies one
thing at a time:

    python scripts/time_window.py [--envs 65536] [--drones 4] [--steps 20] [--reps 30]

  as_bench      warm-up 5, sync, record, K steps in runs of consecutive slots (16-slot buffers), record, sync
  one_call      the same K steps issued by ONE host call (K-slot buffers)
  gated         the K launches are enqueued behind a short spin kernel so that the host's enqueue latency is outside
                the GPU's timeline (what the window would cost if the host were infinitely fast)
  K = 40/80/200 the ramp amortised
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--drones", type=int, default=4)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--reps", type=int, default=30)
args = ap.parse_args()

N, M, A = args.envs, args.drones, 4
dev = torch.device("cuda", 0)
side = int(np.ceil(np.sqrt(M)))
xyz = np.array([[float(i % side), float(i // side), 0.5] for i in range(M)])      # bench.py's layout
env = BatchAviary(task="multihover", num_envs=N, num_drones=M, initial_xyzs=xyz, pyb_freq=240, ctrl_freq=30, act="rpm",
                  precision="fp32", device=dev, auto_reset=True, reset_mode="jitter_philox", seed=1234)
D = env.OBS_DIM
slots = 32
act = (torch.rand((slots, N, M, A), device=dev) * 2 - 1).contiguous()
obs = torch.empty((slots, N, M, D), device=dev)
rew = torch.empty((slots, N), device=dev)
term = torch.empty((slots, N), dtype=torch.uint8, device=dev)
trunc = torch.empty((slots, N), dtype=torch.uint8, device=dev)
env.reset_device(out=obs[0])


def steps(n, start, nslots):
    k = 0
    while k < n:
        i = (start + k) % nslots
        run = min(n - k, nslots - i)
        env.step_many(act[i:i + run], obs[i:i + run], rew[i:i + run], term[i:i + run], trunc[i:i + run])
        k += run


def window(K, nslots, warm=5, idle_ms=0.0):
    steps(warm, 0, nslots)
    torch.cuda.synchronize()
    if idle_ms:
        time.sleep(idle_ms * 1e-3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    steps(K, warm, nslots)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K


def gated(K, nslots, warm=5):
    steps(warm, 0, nslots)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)      # ~1 ms spin on the stream: everything below is enqueued before it ends
    e0.record()
    steps(K, warm, nslots)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K


def report(name, f):
    v = np.array([f() for _ in range(args.reps)])
    print(f"{name:28s} us/step: median {np.median(v):6.2f}  min {v.min():6.2f}  p90 {np.percentile(v, 90):6.2f}", flush=True)


K = args.steps
report(f"as_bench K={K} (16 slots)", lambda: window(K, 16))
report(f"one_call K={K} (32 slots)", lambda: window(K, 32))
report(f"gated K={K}", lambda: gated(K, 32))
report(f"as_bench K={K}, 50 ms idle", lambda: window(K, 16, idle_ms=50.0))
report(f"as_bench K={K}, warm 50", lambda: window(K, 16, warm=50))
for k in (40, 80, 200):
    report(f"as_bench K={k}", lambda: window(k, 16))

# bench.py starts `nvidia-smi -lms 100` right before its timed window: does the tool's start-up disturb the launches?
import subprocess  # noqa: E402


def with_smi(K, settle_s):
    pr = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if settle_s:
        t_end = time.time() + settle_s
        while time.time() < t_end:
            steps(50, 0, 16)
            torch.cuda.synchronize()
    v = window(K, 16)
    pr.terminate()
    pr.wait()
    return v


args.reps = 8
report(f"K={K}, nvidia-smi just started", lambda: with_smi(K, 0.0))
report(f"K={K}, nvidia-smi settled 0.5 s", lambda: with_smi(K, 0.5))


# the same samples taken in-process through NVML (no child process): does the polling itself disturb the window?
import threading  # noqa: E402

import pynvml  # noqa: E402

pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)


def with_nvml(K, period_s):
    stop = threading.Event()
    rows = []

    def poll():
        while not stop.is_set():
            rows.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                         pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)))
            stop.wait(period_s)
    th = threading.Thread(target=poll, daemon=True)
    th.start()
    v = window(K, 16)
    stop.set()
    th.join()
    return v


report(f"K={K}, NVML thread @100 ms", lambda: with_nvml(K, 0.1))
report(f"K={K}, NVML thread @10 ms", lambda: with_nvml(K, 0.01))
report(f"K={K}, no sampler again", lambda: window(K, 16))
