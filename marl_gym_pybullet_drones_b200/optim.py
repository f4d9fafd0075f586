"""Adam whose step can be switched off by a device-side gate, for a PPO update with no host sync.

The reference skips the actor's optimiser step — parameters, both Adam moments and the step count — when a
minibatch's approximate KL exceeds 1.5 x target_kl (`mappo/agent.py:731`).  Deciding that on the host costs one
device->host read per minibatch and makes the update uncapturable in a CUDA graph.  `GatedAdam` keeps all
parameters of an optimiser in ONE flat buffer (the module's parameters become views of it, their `.grad`
views of a flat gradient buffer), does `torch.optim.Adam`'s arithmetic (defaults: betas (0.9, 0.999), eps 1e-8,
no weight decay) as a handful of elementwise ops on the flat tensors, and applies the result through
`torch.where(gate, new, old)`.  The flat gradient buffer is also the unit of the multi-GPU all-reduce.
`state_dict()` / `load_state_dict()` speak `torch.optim.Adam`'s format (per-parameter `step`, `exp_avg`,
`exp_avg_sq`), so checkpoints stay interchangeable with the reference's `model_latest.pt`.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch


class GatedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = list(params)
        if not self.params:
            raise ValueError("GatedAdam got an empty parameter list")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        dev, dt = self.params[0].device, self.params[0].dtype
        self.flat = torch.cat([p.detach().reshape(-1) for p in self.params]).contiguous()
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_t = torch.zeros((), dtype=torch.float64, device=dev)   # fp64: bias corrections as torch's Python floats
        self._one = torch.ones((), dtype=torch.bool, device=dev)
        self._spans = []
        off = 0
        for p in self.params:
            n = p.numel()
            if p.device != dev or p.dtype != dt:
                raise ValueError("all parameters must share one device and dtype")
            p.data = self.flat[off:off + n].view(p.shape)     # the module now reads / writes the flat buffer
            p.grad = self.grad[off:off + n].view(p.shape)
            self._spans.append((off, n))
            off += n

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce_grad(self):
        """Average the flat gradient over ranks: one collective per optimiser step."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad)
            self.grad.div_(dist.get_world_size())

    @torch.no_grad()
    def step(self, gate: Optional[torch.Tensor] = None):
        """One Adam step; `gate` = 0-dim bool tensor on the device (False: nothing changes, not even the
        step count), None = unconditional."""
        b1, b2 = self.betas
        g = self.grad
        on = self._one if gate is None else gate
        step = self.step_t + 1.0
        m = torch.lerp(self.exp_avg, g, 1.0 - b1)
        v = self.exp_avg_sq * b2 + (g * g) * (1.0 - b2)
        step_size = (self.lr / (1.0 - torch.pow(b1, step))).to(g.dtype)
        bias2_sqrt = (1.0 - torch.pow(b2, step)).sqrt().to(g.dtype)
        denom = v.sqrt() / bias2_sqrt + self.eps
        p = self.flat - step_size * (m / denom)
        self.flat.copy_(torch.where(on, p, self.flat))
        self.exp_avg.copy_(torch.where(on, m, self.exp_avg))
        self.exp_avg_sq.copy_(torch.where(on, v, self.exp_avg_sq))
        self.step_t.copy_(torch.where(on, step, self.step_t))

    # ---- torch.optim.Adam's checkpoint format -------------------------------------------------------
    def state_dict(self):
        state = {}
        if float(self.step_t.item()) > 0:     # torch creates the state lazily, at the first step
            for i, (off, n) in enumerate(self._spans):
                shape = self.params[i].shape
                state[i] = {"step": self.step_t.detach().to(torch.float32).cpu(),
                            "exp_avg": self.exp_avg[off:off + n].view(shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd.get("param_groups", [])
        if groups:
            self.lr = float(groups[0].get("lr", self.lr))
            self.betas = tuple(float(b) for b in groups[0].get("betas", self.betas))
            self.eps = float(groups[0].get("eps", self.eps))
        state = sd.get("state", {})
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_t.zero_()
        for i, (off, n) in enumerate(self._spans):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            self.exp_avg[off:off + n].copy_(torch.as_tensor(st["exp_avg"]).reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(torch.as_tensor(st["exp_avg_sq"]).reshape(-1))
            self.step_t.fill_(float(torch.as_tensor(st["step"]).item()))
