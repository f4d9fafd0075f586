"""Observation / reward normalisers of the MAPPO trainer, on the device.

Same classes, constructor keywords, `__call__`, `state_dict` / `load_state_dict`, `set_read_only` /
`unset_read_only` as `safe_control_gym/math_and_models/normalization.py:13-160`, operating on CUDA
tensors.  `MeanStdNormalizer` runs on the `bd_rms_*` kernels (csrc/bd_norm.cu): one HBM pass for the
batch moments, an fp64 parallel-variance merge, and either a standalone normalise pass
(`__call__`) or, in the rollout, float (mean, 1/std) vectors that the fused actor kernel applies to
its input tile (`update` + `stats`), so normalised observations are never materialised.
`RewardStdNormalizer` works on (N,) vectors with a handful of torch ops.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _native


class BaseNormalizer:
    """`normalization.py:61-88`: identity, with the read-only switch."""

    def __init__(self, read_only=False):
        self.read_only = read_only

    def set_read_only(self):
        self.read_only = True

    def unset_read_only(self):
        self.read_only = False

    def __call__(self, x, *args, **kwargs):
        return x

    def state_dict(self):
        return {}

    def load_state_dict(self, _):
        pass


class RunningMeanStd:
    """`normalization.py:13-58` with the statistics resident on the GPU (fp64)."""

    def __init__(self, epsilon=1e-4, shape=(), device=None, eps_div=1e-8):
        if not torch.cuda.is_available():
            raise RuntimeError("RunningMeanStd needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.shape = tuple(shape)
        self.cols = int(np.prod(self.shape)) if self.shape else 1
        self._lib = _native.load()
        self._h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = self._lib.bd_rms_create(self.cols, idx, float(epsilon), float(eps_div), C.byref(self._h))
        if rc != 0:
            raise _native.NativeError(f"bd_rms_create failed ({rc}): {self._lib.bd_rms_last_error().decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        if rc != 0:
            raise _native.NativeError(f"{what} failed ({rc}): {self._lib.bd_rms_last_error().decode()}")

    def update(self, arr: torch.Tensor):
        """arr: (batch, *shape) float32 CUDA tensor."""
        if arr.dtype != torch.float32 or not arr.is_contiguous() or arr.device != self.device:
            raise ValueError("arr must be a contiguous float32 CUDA tensor")
        rows = arr.numel() // self.cols
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # envs sharded over ranks: the batch is the union of every rank's rows (one small all-gather)
            mine = self.batch_moments(arr)
            parts = torch.empty((dist.get_world_size(), mine.numel()), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(parts, mine)
            self.merge_moments(parts)
            return
        self._check(self._lib.bd_rms_update(self._h, C.c_void_p(arr.data_ptr()), rows, self._stream()), "bd_rms_update")

    def batch_moments(self, arr: torch.Tensor) -> torch.Tensor:
        """[mean(cols) | var(cols) | count] of the local batch (float64, on the device)."""
        out = torch.empty(2 * self.cols + 1, dtype=torch.float64, device=self.device)
        self._check(self._lib.bd_rms_batch_moments(self._h, C.c_void_p(arr.data_ptr()), arr.numel() // self.cols,
                                                   C.c_void_p(out.data_ptr()), self._stream()), "bd_rms_batch_moments")
        return out

    def merge_moments(self, parts: torch.Tensor):
        """Fold (parts, 2*cols+1) batch moments — combined as one batch — into the running statistics."""
        if parts.dtype != torch.float64 or not parts.is_contiguous() or parts.shape[-1] != 2 * self.cols + 1:
            raise ValueError("parts must be contiguous float64 (parts, 2*cols+1)")
        self._check(self._lib.bd_rms_merge_moments(self._h, C.c_void_p(parts.data_ptr()), parts.numel() // parts.shape[-1],
                                                   self._stream()), "bd_rms_merge_moments")

    def normalize(self, x: torch.Tensor, clip: float, out: torch.Tensor = None) -> torch.Tensor:
        if x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device:
            raise ValueError("x must be a contiguous float32 CUDA tensor")
        y = out if out is not None else torch.empty_like(x)
        self._check(self._lib.bd_rms_normalize(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()),
                                               x.numel() // self.cols, float(clip), self._stream()), "bd_rms_normalize")
        return y

    def _get(self, want):
        t = {k: torch.empty(n, dtype=dt, device=self.device) for k, (n, dt) in want.items()}
        p = lambda k: C.c_void_p(t[k].data_ptr()) if k in t else None  # noqa: E731
        self._check(self._lib.bd_rms_get(self._h, p("mean"), p("var"), p("count"), p("mean_f"), p("rstd_f"), self._stream()),
                    "bd_rms_get")
        return t

    @property
    def mean(self):
        return self._get({"mean": (self.cols, torch.float64)})["mean"].view(self.shape)

    @property
    def var(self):
        return self._get({"var": (self.cols, torch.float64)})["var"].view(self.shape)

    @property
    def count(self):
        return float(self._get({"count": (1, torch.float64)})["count"].item())

    def stats(self, mean_out: torch.Tensor = None, rstd_out: torch.Tensor = None):
        """float (mean, 1/sqrt(var+eps)) vectors, optionally into caller-owned (cols,) tensors."""
        m = mean_out if mean_out is not None else torch.empty(self.cols, device=self.device)
        r = rstd_out if rstd_out is not None else torch.empty(self.cols, device=self.device)
        self._check(self._lib.bd_rms_get(self._h, None, None, None, C.c_void_p(m.data_ptr()), C.c_void_p(r.data_ptr()),
                                         self._stream()), "bd_rms_get")
        return m, r

    def set(self, mean=None, var=None, count=None):
        ts = []

        def prep(v, n):
            if v is None:
                return None
            t = torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(n)).to(self.device)
            ts.append(t)
            return C.c_void_p(t.data_ptr())
        self._check(self._lib.bd_rms_set(self._h, prep(mean, self.cols), prep(var, self.cols), prep(count, 1), self._stream()),
                    "bd_rms_set")
        torch.cuda.current_stream(self.device).synchronize()

    def close(self):
        if self._h:
            self._lib.bd_rms_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MeanStdNormalizer(BaseNormalizer):
    """`normalization.py:64-96`: normalise by the running average."""

    def __init__(self, shape=(), read_only=False, clip=10.0, epsilon=1e-8, device=None):
        super().__init__(read_only)
        self.rms = RunningMeanStd(shape=shape, device=device, eps_div=epsilon)
        self.clip = clip
        self.epsilon = epsilon

    def update(self, x: torch.Tensor):
        """Statistics update only (the rollout's fused path normalises inside the consumer)."""
        if not self.read_only:
            self.rms.update(x)

    def __call__(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        self.update(x)
        return self.rms.normalize(x, self.clip, out=out)

    def state_dict(self):
        return {"mean": self.rms.mean.cpu().numpy(), "var": self.rms.var.cpu().numpy()}

    def load_state_dict(self, saved):
        self.rms.set(mean=saved["mean"], var=saved["var"])


class RewardStdNormalizer(BaseNormalizer):
    """`normalization.py:99-141`: scale rewards by the running std of the discounted return."""

    def __init__(self, gamma=0.99, read_only=False, clip=10.0, epsilon=1e-8, device=None):
        super().__init__(read_only)
        self.device = torch.device(device if device is not None else "cuda")
        self.gamma, self.clip, self.epsilon = gamma, clip, epsilon
        self.mean = torch.zeros((), dtype=torch.float64, device=self.device)
        self.var = torch.ones((), dtype=torch.float64, device=self.device)
        self.count = torch.full((), 1e-4, dtype=torch.float64, device=self.device)
        self.ret = None

    def __call__(self, x: torch.Tensor, dones: torch.Tensor) -> torch.Tensor:
        if not self.read_only:
            if self.ret is None:
                self.ret = torch.zeros_like(x)
            self.ret = self.ret * self.gamma + x
            r64 = self.ret.double()
            bm, bv, bc = r64.mean(), r64.var(unbiased=False), float(r64.numel())
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                # moments of the union of every rank's returns
                p = torch.stack([r64.sum(), (r64 * r64).sum(), torch.tensor(bc, dtype=torch.float64, device=r64.device)])
                dist.all_reduce(p)
                bc = p[2]                       # stays on the device: no host sync in the rollout
                bm = p[0] / p[2]
                bv = (p[1] / p[2] - bm * bm).clamp_min(0.0)
            delta = bm - self.mean
            tot = self.count + bc
            m2 = self.var * self.count + bv * bc + delta * delta * self.count * bc / tot
            self.mean = self.mean + delta * bc / tot
            self.var = m2 / tot
            self.count = tot
            self.ret = torch.where(dones.bool(), torch.zeros_like(self.ret), self.ret)
        return (x / torch.sqrt(self.var + self.epsilon).to(x.dtype)).clamp(-self.clip, self.clip)

    def state_dict(self):
        return {"mean": self.mean.cpu().numpy(), "var": self.var.cpu().numpy()}

    def load_state_dict(self, saved):
        self.mean = torch.as_tensor(np.asarray(saved["mean"], dtype=np.float64)).to(self.device).reshape(())
        self.var = torch.as_tensor(np.asarray(saved["var"], dtype=np.float64)).to(self.device).reshape(())
