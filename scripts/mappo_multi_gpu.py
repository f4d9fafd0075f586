"""MAPPO with the envs sharded over several GPUs (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/mappo_multi_gpu.py [--envs-per-gpu 4096]

Checks what has to hold across ranks: identical weights and identical normaliser statistics after
training (gradient all-reduce, shared KL-gate decision, all-gathered batch moments), different rollouts
per rank, and prints the aggregate env-steps/s."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_gym_pybullet_drones_b200 import BatchAviary, DeviceMAPPO  # noqa: E402
from marl_gym_pybullet_drones_b200.dist import init_distributed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs-per-gpu", type=int, default=4096)
ap.add_argument("--drones", type=int, default=2)
ap.add_argument("--iters", type=int, default=4)
args = ap.parse_args()
rank, local_rank, world = init_distributed()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
M = args.drones
xyz = np.array([[float(i), 0.0, 0.5] for i in range(M)])
env = BatchAviary(task="multihover", num_envs=args.envs_per_gpu, num_drones=M, initial_xyzs=xyz, seed=100 + rank,
                  track_episode_stats=True, device=dev)
algo = DeviceMAPPO(env, rollout_steps=32, mini_batch_size=8192, opt_epochs=2, norm_obs=True, norm_reward=True, seed=0)
t0 = time.perf_counter()
for it in range(args.iters):
    res = algo.train_step()
    if rank == 0:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0


def spread(t):
    """max |x - x_rank0| over ranks"""
    ref = t.detach().clone().double()
    if world > 1:
        dist.broadcast(ref, src=0)
    d = (t.detach().double() - ref).abs().max()
    if world > 1:
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
    return d.item()


w = torch.cat([p.reshape(-1) for p in algo.ac.parameters()])
dw = spread(w)
dm, dv = spread(algo.obs_normalizer.rms.mean), spread(algo.obs_normalizer.rms.var)
dr = spread(algo.reward_normalizer.var)
dobs = spread(algo.obs[1])
if rank == 0:
    print(f"world={world}  weight spread {dw:.3e}  obs-normaliser mean/var spread {dm:.3e}/{dv:.3e}  reward-var spread {dr:.3e}  "
          f"rollout spread {dobs:.3e} (must be > 0 for world > 1)")
    print(f"{args.iters * 32 * args.envs_per_gpu * world / dt:.3e} env-steps/s over {world} GPUs "
          f"(count {algo.obs_normalizer.rms.count:.1f} = {(args.iters * 32 + 1) * args.envs_per_gpu * world} rows + 1e-4)")
    assert dw == 0.0 and dm == 0.0 and dv == 0.0 and dr == 0.0, "replicas diverged"
    assert world == 1 or dobs > 0.0
algo.close()           # the captured epoch graph contains NCCL collectives: release it before the process group goes
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
