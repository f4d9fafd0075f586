"""All-reduce of the trainer's flat gradient buffer over NVLink peer memory (`csrc/bd_peer.cu`).

Host side of `bd_peer_*`: every rank allocates the same block of device memory in the library, the 64-byte CUDA
IPC handles are all-gathered through `torch.distributed` (plumbing only: any transport would do), every rank maps
its peers' blocks, and from then on `all_reduce()` is ONE kernel launch per rank and call — no NCCL on the data
path.  The gradient tensors of the optimisers are views of the block's data region (`.data`), so the kernels that
produce the gradients write them where the collective reads them.

Replaces the `torch.distributed.all_reduce` pair of `DeviceMAPPO._native_minibatch` (the data-parallel form of
`MAPPOAgent.update`, reference gym_pybullet_drones/mappo/agent.py:702-772).  One node only (CUDA IPC); the caller
falls back to NCCL when `PeerAllReduce.create` returns None.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native


class PeerAllReduce:
    def __init__(self, handle, floats: int, device: torch.device, rank: int, world: int):
        self._lib = _native.load()
        self._h = handle
        self.floats, self.device, self.rank, self.world = int(floats), device, int(rank), int(world)
        self._group = None
        ptr = self._lib.bd_peer_data(self._h)

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        with torch.cuda.device(device):
            self.data = torch.as_tensor(a, device=device)      # view of library-owned memory

    @classmethod
    def create(cls, floats: int, device: torch.device, group=None) -> Optional["PeerAllReduce"]:
        """Collective over `group` (default: the world).  None when the ranks cannot map each other's memory
        (several nodes, no peer access): every rank takes the same decision."""
        import torch.distributed as dist
        lib = _native.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world < 2 or world > 16:
            return None
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        ok = lib.bd_peer_create(int(idx), int(rank), int(world), int(floats), C.byref(h)) == 0
        nbytes = lib.bd_peer_handle_size()
        mine = torch.zeros(nbytes + 1, dtype=torch.uint8)
        if ok:
            buf = (C.c_ubyte * nbytes)()
            ok = lib.bd_peer_get_handle(h, buf) == 0
            if ok:
                mine[:nbytes] = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
                mine[nbytes] = 1
        # the handles travel as a byte tensor through the process group (device tensors for NCCL, host for gloo)
        backend_dev = device if dist.get_backend(group) == "nccl" else torch.device("cpu")
        gathered = [torch.zeros_like(mine, device=backend_dev) for _ in range(world)]
        dist.all_gather(gathered, mine.to(backend_dev), group=group)
        allb = torch.stack([g.cpu() for g in gathered])
        good = bool(allb[:, nbytes].all())
        if good:
            raw = bytes(allb[:, :nbytes].contiguous().numpy().tobytes())
            good = lib.bd_peer_open(h, raw, int(world)) == 0
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=backend_dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if h:
                lib.bd_peer_destroy(h)
            return None
        obj = cls(h, floats, device, rank, world)
        obj._group = group
        return obj

    def all_reduce(self, floats: Optional[int] = None, extra: Optional[torch.Tensor] = None):
        """Sum `.data[:floats]` over the ranks in place, and `extra` (a contiguous float64 device tensor of <= 16
        entries, e.g. the KL pair) in place; one launch on the current stream."""
        n = self.floats if floats is None else int(floats)
        ne, ep = 0, None
        if extra is not None:
            if extra.dtype != torch.float64 or not extra.is_contiguous() or extra.device != self.data.device:
                raise ValueError("extra must be a contiguous float64 tensor on the collective's device")
            ne, ep = extra.numel(), C.c_void_p(extra.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self._lib.bd_peer_allreduce(self._h, n, ep, ne, st)
        if rc != 0:
            raise _native.NativeError(f"bd_peer_allreduce failed ({rc}): {self._lib.bd_peer_last_error().decode()}")

    @property
    def launch_count(self) -> int:
        return int(self._lib.bd_peer_launch_count(self._h))

    def close(self):
        """Collective: every rank unmaps its peers' blocks, the ranks meet, then every rank frees its own block (an
        exporter must not free memory a peer still maps)."""
        if self._h:
            self.data = None
            self._lib.bd_peer_unmap(self._h)
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.barrier(group=self._group)
            self._lib.bd_peer_destroy(self._h)
            self._h = None
