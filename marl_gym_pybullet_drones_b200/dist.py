"""Multi-GPU plumbing: one process per GPU, environments sharded, no data-path collective.

The reference parallelises by spreading env objects over worker processes with
`np.array_split` (`subproc_vec_env.py:28,53`).  Here every rank owns a contiguous
env slice resident in its own GPU's HBM for the whole run; stepping needs no
communication.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used
only for what genuinely crosses ranks: episode-statistics reduction here, and in the
trainer the flat-gradient all-reduce (`optim.GatedAdam.grad`, one collective per optimiser
step), the KL pair of the gate, the advantage moments and the normalisers' batch moments.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_envs(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(start, count) of this rank's contiguous env slice — np.array_split's partition."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def init_distributed(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise from the torchrun environment; returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def reduce_episode_stats(return_sum: torch.Tensor, length_sum: torch.Tensor, episodes: torch.Tensor):
    """Global mean episode return / length from per-rank sums (one small all-reduce).

    Replaces reading `VecRecordEpisodeStatistics.return_queue/length_queue`
    (`record_episode_statistics.py:162-164`, consumed at `mappo/mappo.py:1186-1226`) when
    the envs live on several GPUs."""
    packed = torch.stack([return_sum.double().sum(), length_sum.double().sum(), episodes.double().sum()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    n = packed[2].clamp_min(1.0)
    return (packed[0] / n).item(), (packed[1] / n).item(), int(packed[2].item())


def bind_host_thread_to_gpu(local_rank: int):
    """Pin the calling thread to the CPUs NVML reports as closest to GPU `local_rank` (its NUMA node), so that the
    page-locked buffers this rank allocates afterwards are first-touched on that node and its copies do not cross
    the socket interconnect.  Returns the resulting CPU set, or None when NVML is unavailable or the container's
    cpuset does not contain those CPUs (then nothing changes)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = os.sched_getaffinity(0)
        if not after:
            os.sched_setaffinity(0, before)
            return None
        return sorted(after)
    except Exception:
        return None

