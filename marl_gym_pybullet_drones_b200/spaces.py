"""Minimal stand-ins for `gymnasium.spaces.Box` and `gymnasium.Env`.

Gymnasium is not a dependency of this package; callers of the reference only
read `shape`, `low`, `high`, `dtype` of the spaces (`BaseRLAviary.py:156,277`,
`mappo/mappo.py:60-75`) and call `sample()` / `contains()`.  When gymnasium is
importable the real classes are used instead, so `isinstance` checks in user
code keep working.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium import Env as _GymEnv
    from gymnasium.spaces import Box as _GymBox
except Exception:  # gymnasium absent (this image)
    _GymEnv = None
    _GymBox = None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class _Env:
    metadata = {"render_modes": []}

    def close(self):
        pass


Box = _GymBox if _GymBox is not None else _Box
Env = _GymEnv if _GymEnv is not None else _Env
