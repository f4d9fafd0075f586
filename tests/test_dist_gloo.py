"""world_size=2 gloo tests of the only cross-rank traffic: episode-stat reduction, the flat-gradient all-reduce of
`GatedAdam` (the unit the trainer reduces), minibatch-count agreement for unequal shards."""
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
    from marl_gym_pybullet_drones_b200.dist import init_distributed, reduce_episode_stats, shard_envs
    from marl_gym_pybullet_drones_b200.optim import GatedAdam
    r, lr, w = init_distributed("gloo")
    start, count = shard_envs(10, r, w)
    # each rank finished `count` episodes of return = env index
    rets = torch.arange(start, start + count, dtype=torch.float64)
    mean_r, mean_l, n = reduce_episode_stats(rets, torch.full((count,), 242.0), torch.ones(count))
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    x = torch.full((3, 4), float(r + 1))
    opt = GatedAdam(net.parameters(), lr=1e-2)      # parameters / gradients become views of ONE flat buffer
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    opt.all_reduce_grad()                           # one collective for the whole network
    gathered = [None] * w
    dist.all_gather_object(gathered, [g.tolist() for g in local])
    avg_ok = all(torch.allclose(p.grad, sum(torch.tensor(gathered[k][i]) for k in range(w)) / w)
                 for i, p in enumerate(net.parameters()))
    # unequal shards (10 envs over 2 ranks is equal; 11 is not): every rank must run the same number of minibatches
    s11, c11 = shard_envs(11, r, w)
    t = torch.tensor([c11 * 8], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    q.put((r, mean_r, mean_l, n, int(t.item()), avg_ok))
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
    for r, mean_r, mean_l, n, n_min, avg_ok in res:
        assert n == 10 and mean_r == pytest.approx(4.5) and mean_l == pytest.approx(242.0)
        assert n_min == 5 * 8 and avg_ok


def _worker_trainer(rank, world, port, q):
    """The trainer-side collectives on CPU tensors: GatedAdam's flat-gradient all-reduce and the reward
    normaliser's cross-rank moments (the observation normaliser's kernels need a GPU; its exchange is the same
    all-gather + ordered merge and is checked on 2 and 8 GPUs by scripts/mappo_multi_gpu.py)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import numpy as np
    from marl_gym_pybullet_drones_b200.dist import init_distributed
    from marl_gym_pybullet_drones_b200.normalization import RewardStdNormalizer
    from marl_gym_pybullet_drones_b200.optim import GatedAdam
    r, lr, w = init_distributed("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    opt = GatedAdam(net.parameters(), lr=1e-2)
    xs = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4) / 10.0        # rank r trains on xs[r]
    for it in range(3):
        opt.zero_grad()
        net(xs[r]).pow(2).mean().backward()
        opt.all_reduce_grad()
        kl = torch.tensor(0.01 * (r + 1))                                      # ranks disagree about the gate ...
        dist.all_reduce(kl)
        opt.step((kl / w) <= 0.02 if it != 1 else torch.tensor(False))         # ... until the KL is averaged
    rng = np.random.default_rng(5)
    rew = rng.uniform(-1, 2, (6, 2, 16))                                       # (steps, rank, envs)
    done = rng.uniform(0, 1, (6, 2, 16)) < 0.2
    rn = RewardStdNormalizer(gamma=0.99, device="cpu")
    for t in range(6):
        rn(torch.as_tensor(rew[t, r]), torch.as_tensor(done[t, r]))
    q.put((r, opt.flat.tolist(), float(opt.step_t), float(rn.var), rew, done))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gated_adam_and_reward_normaliser_match_single_process():
    import numpy as np
    from oracle.normalization import RewardStdNormalizerOracle
    from marl_gym_pybullet_drones_b200.optim import GatedAdam
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_trainer, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2] == 2.0            # identical replicas, one step gated off
    # single process on the union of the data: mean of the two ranks' gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    opt = GatedAdam(net.parameters(), lr=1e-2)
    xs = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4) / 10.0
    for it in range(3):
        opt.zero_grad()
        (0.5 * (net(xs[0]).pow(2).mean() + net(xs[1]).pow(2).mean())).backward()
        opt.step(torch.tensor(it != 1))
    assert np.allclose(opt.flat.tolist(), res[0][1], rtol=1e-5, atol=1e-7)
    # reward normaliser: the oracle fed with both ranks' envs side by side
    rew, done = res[0][4], res[0][5]
    o = RewardStdNormalizerOracle(gamma=0.99)
    for t in range(6):
        o(rew[t].reshape(-1), done[t].reshape(-1))
    assert res[0][3] == res[1][3] and abs(res[0][3] - float(o.rms.var)) <= 1e-12 * float(o.rms.var)


def _peer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from marl_gym_pybullet_drones_b200.dist import init_distributed
    from marl_gym_pybullet_drones_b200.peer import PeerAllReduce
    init_distributed("gloo")
    # no GPU here: bd_peer_create fails on every rank, the handle exchange still runs (over gloo, host tensors) and every rank
    # takes the same decision — None, i.e. the trainer keeps NCCL — instead of one rank waiting for the other
    p = PeerAllReduce.create(1000, torch.device("cuda", 0))
    q.put((rank, p is None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.skipif(torch.cuda.is_available(), reason="the CPU-only decision path (GPU boxes run tests/test_gpu_peer_allreduce.py)")
def test_peer_allreduce_setup_falls_back_consistently_without_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
