"""torchrun check of the NVLink peer-memory all-reduce (csrc/bd_peer.cu) against NCCL on the same data.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        scripts/peer_allreduce_check.py

Checks: results identical on every rank; equal to NCCL's sum (bit-exact for 2 ranks, <= 1e-6 relative otherwise);
many back-to-back calls with changing data (flag protocol, buffer reuse); the call inside a CUDA graph; timing.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200.peer import PeerAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 445_191                       # actor + critic parameters of the bench's networks, not a multiple of 4
peer = PeerAllReduce.create(n, dev)
if peer is None:
    print(f"rank {rank}: peer memory not available")
    dist.destroy_process_group()
    sys.exit(1)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(200):
    x = torch.randn(n, device=dev, generator=gen)
    kl = torch.randn(2, device=dev, generator=gen, dtype=torch.float64)
    ref, ref_kl = x.clone(), kl.clone()
    dist.all_reduce(ref)
    dist.all_reduce(ref_kl)
    peer.data.copy_(x)
    peer.all_reduce(extra=kl)
    got = peer.data.clone()
    err = float((got - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    assert err <= (0.0 if world == 2 else 1e-6), (it, err)
    assert float((kl - ref_kl).abs().max()) <= 1e-12, (kl, ref_kl)
    # every rank holds the same bits
    chk = torch.stack([got.double().sum(), got.view(torch.int32).sum().double()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), (it, lo, hi)
torch.cuda.synchronize()
# inside a CUDA graph: 20 calls per replay, data changed between replays
kl = torch.zeros(2, device=dev, dtype=torch.float64)
g = torch.cuda.CUDAGraph()
peer.data.fill_(1.0)
kl.fill_(1.0)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(3):
            peer.all_reduce(extra=kl)
torch.cuda.current_stream().wait_stream(s)
for rep in range(3):
    peer.data.fill_(float(rank + 1 + rep))
    kl.fill_(1.0)
    g.replay()
    torch.cuda.synchronize()
    want = sum(r + 1 + rep for r in range(world)) * world ** 2
    assert float(peer.data[0]) == want and float(peer.data[n - 1]) == want, (float(peer.data[0]), want)
    assert float(kl[0]) == world ** 3


def timed(fn, k=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3


y = torch.randn(n, device=dev)
t_nccl = timed(lambda: (dist.all_reduce(y), dist.all_reduce(kl)))
t_nccl1 = timed(lambda: dist.all_reduce(y))
t_peer = timed(lambda: peer.all_reduce(extra=kl))
t = torch.tensor([t_nccl, t_nccl1, t_peer], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: {n} floats + 2 doubles; worst relative difference to NCCL {worst:.2e}; identical on all ranks; graph ok")
    print(f"NCCL gradients + KL pair: {float(t[0]):.1f} us   NCCL gradients only: {float(t[1]):.1f} us   peer kernel (both): {float(t[2]):.1f} us")
peer.close()
dist.destroy_process_group()
