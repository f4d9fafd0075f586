"""world_size=2 gloo test of the only cross-rank traffic: episode-stat reduction + gradient averaging."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from marl_gym_pybullet_drones_b200.dist import (allreduce_gradients, init_distributed, reduce_episode_stats,
                                                    shard_envs)
    r, lr, w = init_distributed("gloo")
    start, count = shard_envs(10, r, w)
    # each rank finished `count` episodes of return = env index
    rets = torch.arange(start, start + count, dtype=torch.float64)
    mean_r, mean_l, n = reduce_episode_stats(rets, torch.full((count,), 242.0), torch.ones(count))
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    x = torch.full((3, 4), float(r + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    calls = allreduce_gradients(net.parameters(), bucket_bytes=64)
    gathered = [None] * w
    dist.all_gather_object(gathered, [g.tolist() for g in local])
    avg_ok = all(torch.allclose(p.grad, sum(torch.tensor(gathered[k][i]) for k in range(w)) / w)
                 for i, p in enumerate(net.parameters()))
    q.put((r, mean_r, mean_l, n, calls, avg_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_reduction_and_gradient_average():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for r, mean_r, mean_l, n, calls, avg_ok in res:
        assert n == 10 and mean_r == pytest.approx(4.5) and mean_l == pytest.approx(242.0)
        assert calls >= 2 and avg_ok
