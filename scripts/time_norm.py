"""Timing of the running-normaliser kernels (bd_rms_update / bd_rms_normalize) against the HBM roofline."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.normalization import MeanStdNormalizer  # noqa: E402

peak = 6553.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for N, M, D in ((65536, 4, 72), (262144, 4, 72), (131072, 16, 72), (16384, 5, 119)):
    xs = [torch.randn(N, M, D, device="cuda") * 3 + 1 for _ in range(8)]      # 8 x 75 MB rotating > L2
    n = MeanStdNormalizer(shape=(M, D), device="cuda")
    y = torch.empty_like(xs[0])

    def timeit(fn, reps=40):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3

    t_up = timeit(lambda i: n.update(xs[i % 8]))
    t_no = timeit(lambda i: n.rms.normalize(xs[i % 8], 10.0, out=y))
    by = xs[0].numel() * 4
    print(f"N={N} M={M} D={D} ({by / 1e6:.1f} MB): update {t_up:.1f} us = {by / t_up / 1e3:.0f} GB/s ({by / t_up / 1e3 / peak:.2f} of peak), "
          f"normalize {t_no:.1f} us = {2 * by / t_no / 1e3:.0f} GB/s ({2 * by / t_no / 1e3 / peak:.2f})")
