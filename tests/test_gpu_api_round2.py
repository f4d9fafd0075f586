"""Round-2 additions to the C-ABI, each against the path it replaces:
`bd_step_host_compact` vs `bd_step_host` with a full terminal-observation buffer, `bd_step_many` vs k `bd_step`
calls, `bd_get/set_rng_state`, relaxed pointer alignment for the scalar shapes, the bounded tile-epoch wait."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest
import torch

from _util import batch_from_cfg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mh(M):
    return dict(task="multihover", drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30, act="rpm")


GRID = np.array([[0.0, 0.0, 0.5], [1.0, 0.0, 0.5], [0.0, 1.0, 0.5], [1.0, 1.0, 0.5]])


@pytest.mark.parametrize("M,N", [(3, 37), (4, 5000), (4, 40001)])
def test_compact_terminal_obs_equals_full_buffer(M, N):
    """Same envs, same actions: the compact call returns exactly the rows of the finished envs, ascending."""
    envs = [batch_from_cfg(_mh(M), GRID[:M], None, num_envs=N, precision="fp32", auto_reset=True,
                           reset_mode="jitter_philox", seed=5) for _ in range(2)]
    for e in envs:
        e.reset_device()
    rng = np.random.default_rng(0)
    total = 0
    for t in range(30):
        a = (rng.uniform(-1, 1, (N, M, 4)) - 0.6).astype(np.float32)
        full = envs[0].step_host(a, want_terminal_obs=True)
        comp = envs[1].step_host(a, compact_terminal_obs=True, buffer_set=t & 1)
        done = full["terminated"] | full["truncated"]
        assert np.array_equal(full["obs"], comp["obs"]) and np.array_equal(full["reward"], comp["reward"])
        assert np.array_equal(done, comp["terminated"] | comp["truncated"])
        idx = np.flatnonzero(done)
        assert np.array_equal(comp["done_idx"], idx)
        assert np.array_equal(comp["terminal_rows"], full["terminal_obs"][idx])
        total += len(idx)
    assert total > min(N, 200)          # includes steps with more finished envs than the initial staging capacity
    for e in envs:
        e.close()


def test_vec_env_step_keeps_previous_result_valid_for_one_step():
    from marl_gym_pybullet_drones_b200.vec_env import BatchVecEnv
    env = BatchVecEnv(batch_from_cfg(_mh(2), GRID[:2], None, num_envs=64, precision="fp32", auto_reset=True,
                                     reset_mode="jitter_philox", seed=1))
    env.reset()
    rng = np.random.default_rng(1)
    o1, r1, d1, i1 = env.step(rng.uniform(-1, 1, (64, 2, 4)).astype(np.float32))
    keep = o1.copy()
    o2, r2, d2, i2 = env.step(rng.uniform(-1, 1, (64, 2, 4)).astype(np.float32))
    assert np.array_equal(o1, keep) and not np.array_equal(o1, o2)
    assert len(i2['n']) == 64 and i2['n'][5]["answer"] == 42 and i2['n'][-1] is i2['n'][63]
    env.close()


@pytest.mark.parametrize("M,N,precision", [(4, 8192, "fp32"), (3, 500, "fp32"), (2, 300, "fp64")])
def test_step_many_equals_k_steps(M, N, precision):
    from marl_gym_pybullet_drones_b200.batch_aviary import StepResult
    K = 40
    envs = [batch_from_cfg(_mh(M), GRID[:M], None, num_envs=N, precision=precision, auto_reset=True,
                           reset_mode="jitter_philox", seed=9, action_dtype=torch.float32) for _ in range(2)]
    for e in envs:
        e.reset_device()
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = (torch.rand((K, N, M, 4), device="cuda", generator=gen) * 2 - 1.4).contiguous()
    D, rd = envs[0].OBS_DIM, envs[0].real_dtype
    obs = torch.empty((K, N, M, D), device="cuda")
    rew = torch.empty((K, N), device="cuda", dtype=rd)
    term = torch.empty((K, N), device="cuda", dtype=torch.bool)
    trunc = torch.empty((K, N), device="cuda", dtype=torch.bool)
    envs[0].step_many(acts, obs, rew, term, trunc)
    for k in range(K):
        r = envs[1].step_device(acts[k])
        assert torch.equal(r.obs, obs[k]) and torch.equal(r.reward, rew[k]), k
        assert torch.equal(r.terminated, term[k]) and torch.equal(r.truncated, trunc[k]), k
    assert bool(term.any())
    # and the two handles stay interchangeable afterwards
    a = acts[0]
    assert torch.equal(envs[0].step_device(a).obs, envs[1].step_device(a).obs)
    for e in envs:
        e.close()


@pytest.mark.parametrize("cfg,M,N,physics,A", [
    (dict(task="spiral", drone_model="cf2x", num_drones=5, pyb_freq=240, ctrl_freq=48, act="rpm"), 5, 700, "dyn_gnd_drag_dw", 4),
    (dict(task="multihover", drone_model="cf2x", num_drones=16, pyb_freq=240, ctrl_freq=30, act="rpm"), 16, 300, "dyn_dw", 4),
    (dict(task="multihover", drone_model="racer", num_drones=2, pyb_freq=240, ctrl_freq=48, act="one_d_rpm"), 2, 1000, "dyn", 1),
    (dict(task="flock", drone_model="cf2x", num_drones=4, pyb_freq=240, ctrl_freq=30, act="rpm"), 4, 900, "dyn", 4),
    (dict(task="hover", drone_model="cf2x", num_drones=1, pyb_freq=240, ctrl_freq=30, act="rpm"), 1, 5000, "dyn", 4),
])
def test_step_many_one_launch_equals_k_launches_all_flavours(cfg, M, N, physics, A):
    """The K-steps-in-one-launch kernel (states in registers, history in shared memory across steps) against K single
    launches: bit-identical observations, rewards, flags, device state and episode statistics for every kernel flavour
    (aero terms, ONE_D_RPM rows, swarm rewards, Spiral extras), with re-spawns inside the K steps, ragged last tiles,
    and single steps before / after so that ring head and step counters are in the middle of their ranges."""
    from marl_gym_pybullet_drones_b200.batch_aviary import StepResult
    K = 37
    side = int(np.ceil(np.sqrt(M)))
    xyz = np.array([[0.8 * (i % side), 0.8 * (i // side), 0.3 + 0.05 * i] for i in range(M)])
    envs = [batch_from_cfg(cfg, xyz, None, num_envs=N, precision="fp32", physics=physics, auto_reset=True,
                           reset_mode="jitter_philox" if cfg["task"] == "multihover" else "fixed", seed=11,
                           track_episode_stats=True) for _ in range(2)]
    for e in envs:
        e.reset_device()
    gen = torch.Generator(device="cuda").manual_seed(5)
    acts = (torch.rand((K + 7, N, M, A), device="cuda", generator=gen) * 2 - 1.3).contiguous()
    for k in range(4):                       # a few single steps first: ring head != 0, counters running
        ra, rb = envs[0].step_device(acts[K + k]), envs[1].step_device(acts[K + k])
        assert torch.equal(ra.obs, rb.obs)
    D = envs[0].OBS_DIM
    obs = torch.empty((K, N, M, D), device="cuda")
    rew = torch.empty((K, N), device="cuda")
    term = torch.empty((K, N), device="cuda", dtype=torch.bool)
    trunc = torch.empty((K, N), device="cuda", dtype=torch.bool)
    l0 = envs[0].launch_count
    envs[0].step_many(acts[:K], obs, rew, term, trunc)
    assert envs[0].launch_count - l0 == 1            # ONE launch, not K
    for k in range(K):
        r = envs[1].step_device(acts[k])
        assert torch.equal(r.obs, obs[k]), (k, float((r.obs - obs[k]).abs().max()))
        assert torch.equal(r.reward, rew[k]), k
        assert torch.equal(r.terminated, term[k]) and torch.equal(r.truncated, trunc[k]), k
    assert bool((term | trunc).any())
    sa, sb = envs[0].get_state(), envs[1].get_state()
    assert torch.equal(torch.nan_to_num(sa), torch.nan_to_num(sb))   # (the fast kernel does not keep world angular velocities: NaN)
    ea, eb = envs[0].episode_stats(), envs[1].episode_stats()
    assert abs(float(ea[2]) - float(eb[2])) == 0 and abs(float(ea[0]) - float(eb[0])) <= 1e-6 * max(1.0, abs(float(eb[0])))
    for k in range(3):                       # and the handles stay interchangeable (epochs, counters, ring)
        ra, rb = envs[0].step_device(acts[K + 4 + k]), envs[1].step_device(acts[K + 4 + k])
        assert torch.equal(ra.obs, rb.obs) and torch.equal(ra.reward, rb.reward)
    # a second K-step launch right after single steps
    envs[0].step_many(acts[:K], obs, rew, term, trunc)
    for k in range(K):
        r = envs[1].step_device(acts[k])
        assert torch.equal(r.obs, obs[k]) and torch.equal(r.reward, rew[k]), k
    for e in envs:
        e.close()


@pytest.mark.parametrize("task,M,physics,act", [
    ("multihover", 1, "dyn", "rpm"), ("multihover", 3, "dyn_dw", "rpm"), ("multihover", 6, "dyn_gnd_drag_dw", "rpm"),
    ("multihover", 7, "dyn", "one_d_rpm"), ("multihover", 8, "dyn_drag", "rpm"), ("multihover", 32, "dyn_dw", "rpm"),
    ("hover", 1, "dyn_gnd", "one_d_rpm"), ("spiral", 4, "dyn", "rpm"), ("spiral", 3, "dyn_dw", "rpm"),
    ("meetup", 4, "dyn", "rpm"), ("meetup", 3, "dyn", "rpm"), ("leaderfollower", 2, "dyn", "rpm"), ("leaderfollower", 5, "dyn", "rpm"),
    ("flock", 8, "dyn", "rpm"), ("flock", 16, "dyn_dw", "rpm"),
])
def test_step_many_equals_k_steps_across_team_sizes_tasks_and_physics(task, M, physics, act):
    """Differential sweep: `step_many` (one launch where the fast kernel applies, K launches where the generic kernel has
    to run — e.g. swarm tasks whose team size is not a power of two) against K `step_device` calls, bit for bit, with a
    ragged number of envs and re-spawns inside the window."""
    if task == "hover" and M != 1:
        pytest.skip("hover is single-drone")
    N, K = 517, 23
    cfg = dict(task=task, drone_model="cf2x", num_drones=M, pyb_freq=240, ctrl_freq=30 if task != "spiral" else 48, act=act)
    side = int(np.ceil(np.sqrt(M)))
    xyz = np.array([[0.7 * (i % side), 0.7 * (i // side), 0.25 + 0.04 * (i % 7)] for i in range(M)])
    envs = [batch_from_cfg(cfg, xyz, None, num_envs=N, precision="fp32", physics=physics, auto_reset=True,
                           reset_mode="jitter_philox" if task == "multihover" else "fixed", seed=2) for _ in range(2)]
    for e in envs:
        e.reset_device()
    A = envs[0].ACTION_DIM
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + len(task))
    acts = (torch.rand((K, N, M, A), device="cuda", generator=gen) * 2 - 1.25).contiguous()
    D = envs[0].OBS_DIM
    obs = torch.empty((K, N, M, D), device="cuda")
    rew = torch.empty((K, N), device="cuda")
    term = torch.empty((K, N), device="cuda", dtype=torch.bool)
    trunc = torch.empty((K, N), device="cuda", dtype=torch.bool)
    for rep in range(2):
        envs[0].step_many(acts, obs, rew, term, trunc)
        for k in range(K):
            r = envs[1].step_device(acts[k])
            assert torch.equal(r.obs, obs[k]), (rep, k, float((r.obs - obs[k]).abs().max()))
            assert torch.equal(r.reward, rew[k]) and torch.equal(r.terminated, term[k]) and torch.equal(r.truncated, trunc[k]), (rep, k)
    for e in envs:
        e.close()


def test_step_many_one_launch_in_a_cuda_graph():
    """The K-step kernel captured into a CUDA graph: replays read the ring head from the device-resident counter and the
    last CTA out advances it by K; eager single steps before, between and after keep working on the same handle."""
    M, N, K = 4, 3000, 6
    cfg = _mh(M)
    eager = batch_from_cfg(cfg, GRID, None, num_envs=N, precision="fp32", auto_reset=True, reset_mode="jitter_philox", seed=3)
    graphed = batch_from_cfg(cfg, GRID, None, num_envs=N, precision="fp32", auto_reset=True, reset_mode="jitter_philox", seed=3)
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = (torch.rand((K, N, M, 4), generator=gen, device="cuda") * 2 - 1.3).contiguous()
    obs = torch.empty((K, N, M, 72), device="cuda")
    rew = torch.empty((K, N), device="cuda")
    term = torch.empty((K, N), dtype=torch.bool, device="cuda")
    trunc = torch.empty((K, N), dtype=torch.bool, device="cuda")
    assert torch.equal(eager.reset_device(), graphed.reset_device())
    for k in range(2):
        assert torch.equal(eager.step_device(acts[k]).obs, graphed.step_device(acts[k]).obs)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            graphed.step_many(acts, obs, rew, term, trunc)
    torch.cuda.synchronize()
    for rep in range(4):          # 24 steps: more than one trip around the 15-slot ring
        g.replay()
        torch.cuda.synchronize()
        for k in range(K):
            r = eager.step_device(acts[k])
            assert torch.equal(r.obs, obs[k]), (rep, k)
            assert torch.equal(r.reward, rew[k]) and torch.equal(r.terminated, term[k]) and torch.equal(r.truncated, trunc[k])
    ra, rb = eager.step_device(acts[0]), graphed.step_device(acts[0])     # eager step on the captured handle
    assert torch.equal(ra.obs, rb.obs)
    g.replay()
    torch.cuda.synchronize()
    for k in range(K):
        assert torch.equal(eager.step_device(acts[k]).obs, obs[k]), k
    eager.close()
    graphed.close()


def test_rng_state_roundtrip_continues_the_respawn_stream():
    """A fresh handle that is given the state of a running one draws the SAME re-spawn positions from then on;
    without it, it would replay the stream from the start."""
    M, N = 4, 600
    mk = lambda: batch_from_cfg(_mh(M), GRID, None, num_envs=N, precision="fp32", auto_reset=True,   # noqa: E731
                                reset_mode="jitter_philox", seed=21)
    a_env, b_env, c_env = mk(), mk(), mk()
    gen = torch.Generator(device="cuda").manual_seed(0)
    down = torch.rand((N, M, 4), device="cuda", generator=gen) * 2 - 1.5
    for e in (a_env,):
        e.reset_device()
        for _ in range(37):
            e.step_device(down)
    st = a_env.get_rng_state()
    assert st[0] == 21 and st[1] == 37
    b_env.set_rng_state(st)
    assert b_env.get_rng_state() == st
    # drive all three from the same physical state: explicit masked reset of every env draws jitter from the stream
    oa, ob, oc = a_env.reset_device(), b_env.reset_device(), c_env.reset_device()
    assert torch.equal(oa[..., :3], ob[..., :3])          # restored stream: same draws
    assert not torch.equal(oa[..., :3], oc[..., :3])      # fresh stream: different draws
    for e in (a_env, b_env, c_env):
        e.close()


def test_scalar_shapes_accept_unaligned_slices():
    """ONE_D_RPM (A = 1, D = 27): a rollout slot obs[t] / act[t] is 16-byte aligned only when N*M % 4 == 0
    (ADVICE r1): the kernels take any 4-byte aligned pointer for the scalar shapes."""
    from marl_gym_pybullet_drones_b200.batch_aviary import StepResult
    cfg = dict(task="multihover", drone_model="cf2x", num_drones=3, pyb_freq=240, ctrl_freq=30, act="one_d_rpm")
    for M, N in ((3, 7), (1, 131), (2, 129)):          # generic kernel (M = 3) and fast tile kernel (M = 1, 2)
        cfg["num_drones"] = M
        xyz = GRID[:M]
        envs = [batch_from_cfg(cfg, xyz, None, num_envs=N, precision="fp32") for _ in range(2)]
        T = 4
        obs = torch.zeros((T + 1, N, M, 27), device="cuda")
        act = (torch.rand((T, N, M, 1), device="cuda") - 0.5).contiguous()
        rew = torch.zeros((T, N), device="cuda")
        fl = torch.zeros((2, T, N), dtype=torch.bool, device="cuda")
        for e in envs:
            e.reset_device()
        assert any(obs[t + 1].data_ptr() % 16 for t in range(T))     # some slots really are unaligned
        for t in range(T):
            envs[0].step_device(act[t], out=StepResult(obs[t + 1], rew[t], fl[0, t], fl[1, t], None))
            r = envs[1].step_device(act[t].clone())
            assert torch.equal(r.obs, obs[t + 1]), (M, N, t)
        for e in envs:
            e.close()


def test_epoch_wait_is_bounded():
    """A tile epoch that never arrives (protocol bug, here injected with bd_debug_set_tile_epoch) must make the launch
    FAIL after about a second, not hang the GPU.  Runs in a child process: the trap poisons the CUDA context."""
    code = r"""
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
from marl_gym_pybullet_drones_b200 import _native
N, M = 65536, 4
env = BatchAviary(task="multihover", num_envs=N, num_drones=M, precision="fp32",
                  initial_xyzs=np.array([[0, 0, .5], [1, 0, .5], [0, 1, .5], [1, 1, .5]], dtype=float))
env.reset_device()
a = torch.zeros((N, M, 4), device="cuda")
for _ in range(5):
    env.step_device(a)
torch.cuda.synchronize()
lib = _native.load()
assert lib.bd_debug_set_tile_epoch(env._h, 3, -1000, None) == 0
t0 = time.time()
try:
    for _ in range(3):                    # back-to-back: the 2nd / 3rd launch take the tile-epoch wait
        env.step_device(a)
    torch.cuda.synchronize()
    print("NOERROR")
except Exception as ex:
    print("TRAPPED %%.2f %%s" %% (time.time() - t0, type(ex).__name__))
""" % ROOT
    t0 = time.time()
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    out = res.stdout + res.stderr
    assert "TRAPPED" in out, out[-2000:]
    assert time.time() - t0 < 90
    # the device is usable again for a new context
    x = torch.ones(4, device="cuda") * 2
    assert float(x.sum()) == 8.0
