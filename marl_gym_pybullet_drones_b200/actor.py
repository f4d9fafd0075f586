"""`FusedActor`: the MAPPO actor's rollout-time forward pass as one tcgen05/TMEM kernel.

Drop-in for the batched branch of `MAPPOActorCritic.step` (reference `mappo/agent.py:389-415`):
`act, logp = FusedActor.forward(obs)` with obs (rows, obs_dim) on the GPU.  Weights are taken
from the fp32 `ActorCritic` (mappo.py here) and repacked to bf16 tensor-core tiles on the device
(`bd_actor_set_weights`), which is cheap enough to redo after every optimiser phase.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native


class FusedActor:
    def __init__(self, obs_dim: int, hidden: int, act_dim: int, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("FusedActor needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.obs_dim, self.hidden, self.act_dim = int(obs_dim), int(hidden), int(act_dim)
        self._lib = _native.load()
        self._h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = self._lib.bd_actor_create(self.obs_dim, self.hidden, self.act_dim, idx, C.byref(self._h))
        if rc != 0:
            raise _native.NativeError(f"bd_actor_create failed ({rc}): {self._lib.bd_actor_last_error().decode()}")
        self._calls = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        if rc != 0:
            raise _native.NativeError(f"{what} failed ({rc}): {self._lib.bd_actor_last_error().decode()}")

    @torch.no_grad()
    def set_weights(self, mlp, logstd: torch.Tensor):
        """`mlp.fcs` = three nn.Linear (obs_dim->hidden->hidden->act_dim), `logstd` (act_dim,)."""
        fcs = list(mlp.fcs)
        if len(fcs) != 3:
            raise ValueError("FusedActor fuses exactly two hidden layers")
        ts = []
        for fc in fcs:
            ts += [fc.weight.detach().float().contiguous(), fc.bias.detach().float().contiguous()]
        ts.append(logstd.detach().float().contiguous())
        assert tuple(ts[0].shape) == (self.hidden, self.obs_dim) and tuple(ts[2].shape) == (self.hidden, self.hidden)
        assert tuple(ts[4].shape) == (self.act_dim, self.hidden)
        self._check(self._lib.bd_actor_set_weights(self._h, *[C.c_void_p(t.data_ptr()) for t in ts], self._stream()),
                    "bd_actor_set_weights")
        self._keep = ts   # alive until the stream has consumed them

    @torch.no_grad()
    def forward(self, obs: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
                out_act: Optional[torch.Tensor] = None, out_logp: Optional[torch.Tensor] = None,
                want_mean: bool = False):
        rows = obs.numel() // self.obs_dim
        if obs.dtype != torch.float32 or not obs.is_contiguous() or obs.device != self.device:
            raise ValueError("obs must be a contiguous float32 CUDA tensor")
        act = out_act if out_act is not None else torch.empty((rows, self.act_dim), device=self.device)
        logp = out_logp if out_logp is not None else torch.empty((rows,), device=self.device)
        mean = torch.empty((rows, self.act_dim), device=self.device) if want_mean else None
        if noise is not None:
            noise = noise.to(self.device, torch.float32).contiguous()
        self._calls += 1
        self._check(self._lib.bd_actor_forward(
            self._h, C.c_void_p(obs.data_ptr()), rows, C.c_void_p(noise.data_ptr()) if noise is not None else None,
            int(seed) & 0xFFFFFFFFFFFFFFFF, self._calls, C.c_void_p(act.data_ptr()), C.c_void_p(logp.data_ptr()),
            C.c_void_p(mean.data_ptr()) if mean is not None else None, self._stream()), "bd_actor_forward")
        return (act, logp, mean) if want_mean else (act, logp)

    def set_input_norm(self, mean: Optional[torch.Tensor], rstd: Optional[torch.Tensor] = None, period: int = 1,
                       clip: float = 10.0):
        """Normalise input rows on load with float (mean, 1/std) vectors of `period * obs_dim` entries
        (`MeanStdNormalizer` of shape (M, D): period = M); `None` switches it off."""
        if mean is None:
            self._check(self._lib.bd_actor_set_input_norm(self._h, None, None, 1, 10.0), "bd_actor_set_input_norm")
            self._norm = None
            return
        if mean.dtype != torch.float32 or rstd.dtype != torch.float32 or mean.numel() != period * self.obs_dim:
            raise ValueError("mean / rstd must be float32 with period * obs_dim entries")
        self._norm = (mean, rstd)   # keep alive
        self._check(self._lib.bd_actor_set_input_norm(self._h, C.c_void_p(mean.data_ptr()), C.c_void_p(rstd.data_ptr()),
                                                      int(period), float(clip)), "bd_actor_set_input_norm")

    @property
    def launch_count(self):
        return int(self._lib.bd_actor_launch_count(self._h))

    def close(self):
        if self._h:
            self._lib.bd_actor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
