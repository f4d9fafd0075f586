"""GPU tests of the reference-facing protocols: Gymnasium single-env classes and the VecEnv 4-tuple."""
import functools

import numpy as np
import pytest
import torch

from _util import load_golden, rel_err

pytestmark = pytest.mark.gpu


def test_hover_env_follows_reference_golden():
    from marl_gym_pybullet_drones_b200 import ActionType, HoverAviary, Physics
    cfg, g = load_golden("hover_one_d_rpm")
    env = HoverAviary(physics=Physics.DYN, pyb_freq=240, ctrl_freq=30, act=ActionType.ONE_D_RPM)
    assert env.observation_space.shape == (1, 27) and env.action_space.shape == (1, 1)
    obs, info = env.reset(seed=3, options={})
    assert obs.shape == (1, 27) and obs.dtype == np.float32 and info == {"answer": 42}
    for t in range(244):
        o, r, te, tr, info = env.step(g["actions"][t])
        assert isinstance(r, float) and isinstance(te, bool) and isinstance(tr, bool)
        assert rel_err(o, g["obs"][t]) <= 2.5e-7 and rel_err(r, g["reward"][t]) <= 1e-9
        assert te == bool(g["terminated"][t]) and tr == bool(g["truncated"][t])
    sv = env._getDroneStateVector(0)
    assert sv.shape == (20,) and rel_err(sv[:16], g["states"][243][0, :16]) <= 1e-9
    env.close()


def test_multihover_env_reset_uses_numpy_global_rng_like_reference():
    from marl_gym_pybullet_drones_b200 import MultiHoverAviary
    cfg, g = load_golden("multihover2_gauss_f32")       # generated with np.random.seed(1) before reset
    env = MultiHoverAviary(num_drones=2, pyb_freq=240, ctrl_freq=30)
    np.random.seed(1)
    obs, info = env.reset()
    assert np.array_equal(env.INIT_XYZS, g["init_xyzs"])           # same MT19937 draws, same rejection rule
    assert np.allclose(env.TARGET_POS, g["target_pos"], atol=0)
    assert rel_err(obs, g["obs0"]) <= 2.5e-7 and info["termination_reasons"] == []
    for t in range(30):
        o, r, te, tr, info = env.step(g["actions"][t])
        assert rel_err(o, g["obs"][t]) <= 2.5e-7 and rel_err(r, g["reward"][t]) <= 1e-9
        assert te == bool(g["terminated"][t])
        if te:
            assert info["termination_reasons"] and info["termination_reasons"][0].startswith("Drone ")
    env.close()


def test_spiral_env_info_and_shapes():
    from marl_gym_pybullet_drones_b200 import SpiralFormationAviary
    cfg, g = load_golden("spiral5_gauss_f32")
    env = SpiralFormationAviary(num_drones=5, act="rpm")
    assert env.observation_space.shape == (5, 119) and env.CTRL_FREQ == 48 and env.EPISODE_LEN_SEC == 12
    obs, info = env.reset()
    assert rel_err(obs, g["obs0"]) <= 2.5e-7 and info["time"] == 0.0
    for t in range(10):
        o, r, te, tr, info = env.step(g["actions"][t])
        assert rel_err(o, g["obs"][t]) <= 2.5e-7 and rel_err(r, g["reward"][t]) <= 1e-9
        assert info["time"] == pytest.approx(5 * t / 240) and info["radius"] == 0.4
    env.close()


@pytest.mark.parametrize("cls,name", [("MeetupAviary", "meetup2_default"), ("FlockAviary", "flock5"),
                                      ("LeaderFollowerAviary", "leaderfollower2_climb")])
def test_swarm_envs_follow_reference_golden(cls, name):
    """Gymnasium views of the swarm tasks: constructor defaults (2 drones, default spawn) and 5-tuples."""
    import marl_gym_pybullet_drones_b200 as pkg
    cfg, g = load_golden(name)
    kw = {}
    if "initial_xyzs" in cfg:
        kw = dict(initial_xyzs=np.array(cfg["initial_xyzs"]), initial_rpys=np.array(cfg["initial_rpys"]))
    env = getattr(pkg, cls)(num_drones=cfg["num_drones"], **kw) if kw else getattr(pkg, cls)()
    assert env.NUM_DRONES == cfg["num_drones"] and env.EPISODE_LEN_SEC == 8
    assert np.allclose(env.INIT_XYZS, g["init_xyzs"], atol=0)
    obs, info = env.reset()
    assert rel_err(obs, g["obs0"]) <= 2.5e-7 and info == {"answer": 42}
    for t in range(min(80, g["actions"].shape[0])):
        o, r, te, tr, info = env.step(g["actions"][t])
        assert rel_err(o, g["obs"][t]) <= 2.5e-7 and rel_err(r, g["reward"][t]) <= 1e-9, (name, t)
        assert te == bool(g["terminated"][t]) and tr == bool(g["truncated"][t]), (name, t)
    env.close()


def test_pyb_physics_is_rejected():
    from marl_gym_pybullet_drones_b200 import HoverAviary, Physics
    with pytest.raises(NotImplementedError):
        HoverAviary(physics=Physics.PYB)


def test_vec_env_protocol_and_episode_statistics():
    """mappo.py:47-54 usage: make_vec_envs(...) wrapped in VecRecordEpisodeStatistics, 4-tuple step."""
    from marl_gym_pybullet_drones_b200 import MultiHoverAviary, VecRecordEpisodeStatistics, make_vec_envs
    grid = np.array([[0.0, 0.0, 0.2], [1.0, 0.0, 0.2]])
    env_func = functools.partial(MultiHoverAviary, num_drones=2, initial_xyzs=grid, pyb_freq=240, ctrl_freq=30)
    N = 8
    env = VecRecordEpisodeStatistics(make_vec_envs(env_func, batch_size=N, n_processes=4, seed=3), deque_size=100)
    assert env.num_envs == N and env.observation_space.shape == (2, 72) and env.action_space.shape == (2, 4)
    obs, info = env.reset()
    assert obs.shape == (N, 2, 72) and len(info["n"]) == N
    rng = np.random.default_rng(0)
    finished = 0
    for t in range(80):
        a = (rng.uniform(-1, 1, (N, 2, 4)) - 0.7).astype(np.float32)
        obs, rew, done, info = env.step(a)
        assert obs.shape == (N, 2, 72) and rew.shape == (N,) and done.shape == (N,) and done.dtype == np.bool_
        for e in range(N):
            inf = info["n"][e]
            if done[e]:
                finished += 1
                assert inf["terminal_observation"].shape == (2, 72)
                assert "termination_reasons" in inf["terminal_info"]
                assert inf["episode"]["l"] >= 1 and np.isfinite(inf["episode"]["r"])
                assert obs[e, 0, 2] >= 0.1 - 1e-6                  # reset obs, not the crashed one
                assert inf["terminal_observation"][:, 2].min() < 0.03 or \
                    np.abs(inf["terminal_observation"][:, 3:5]).max() > 1.2 or \
                    np.abs(inf["terminal_observation"][:, 0:2]).max() > 3.0
            else:
                assert "terminal_observation" not in inf
    assert finished >= N and len(env.return_queue) == finished
    env.close()


def test_vec_env_protocol_for_swarm_tasks():
    """`make_vec_envs(partial(<swarm aviary>))`: 4-tuples, auto-reset with terminal observation, `info == {"answer": 42}`
    semantics of the three swarm tasks behind the reference's VecEnv protocol."""
    from marl_gym_pybullet_drones_b200 import FlockAviary, LeaderFollowerAviary, MeetupAviary, make_vec_envs
    for cls, M in ((MeetupAviary, 2), (FlockAviary, 3), (LeaderFollowerAviary, 2)):
        env = make_vec_envs(functools.partial(cls, num_drones=M), batch_size=6, n_processes=2, seed=1)
        obs, info = env.reset()
        assert obs.shape == (6, M, 72) and env.action_space.shape == (M, 4)
        rng = np.random.default_rng(2)
        dones = 0
        for t in range(40):
            obs, rew, done, info = env.step((rng.uniform(-1, 1, (6, M, 4)) - 0.8).astype(np.float32))
            assert rew.shape == (6,) and np.isfinite(rew).all()
            for e in np.nonzero(done)[0]:
                dones += 1
                assert info["n"][e]["terminal_observation"].shape == (M, 72)
                assert obs[e, 0, 2] == pytest.approx(0.1125, abs=1e-6)      # default spawn height after the reset
        assert dones >= 6, cls.__name__
        env.close()


def test_multi_device_handles_are_independent():
    """Two aviaries (two handles) on the same GPU do not share state."""
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    a = BatchAviary(task="hover", num_envs=4)
    b = BatchAviary(task="hover", num_envs=4)
    a.reset_device()
    b.reset_device()
    act = torch.full((4, 1, 4), 0.5, device="cuda")
    ra = a.step_device(act)
    rb = b.step_device(torch.zeros_like(act))
    assert not torch.equal(ra.obs, rb.obs)
    rb2 = b.step_device(torch.zeros_like(act))
    assert float(rb2.obs[0, 0, 2]) == pytest.approx(0.1125, abs=1e-6)
    a.close()
    b.close()
    with pytest.raises(RuntimeError):
        a.step_device(act)


@pytest.mark.parametrize("M,physics", [(4, "dyn"), (3, "dyn")])   # fast tile kernel / generic kernel
def test_device_episode_statistics_match_host_bookkeeping(M, physics):
    """bd_episode_stats == VecRecordEpisodeStatistics semantics (record_episode_statistics.py:144-171)."""
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    N, T = 300, 120
    side = int(np.ceil(np.sqrt(M)))
    xyz = np.array([[float(i % side), float(i // side), 0.3] for i in range(M)])
    env = BatchAviary(task="multihover", num_envs=N, num_drones=M, initial_xyzs=xyz, physics=physics, seed=4,
                      track_episode_stats=True)
    env.reset_device()
    g = torch.Generator(device="cuda").manual_seed(0)
    ep_r, ep_l = np.zeros(N), np.zeros(N)
    rets, lens = [], []
    for t in range(T):
        a = torch.rand((N, M, 4), generator=g, device="cuda") * 2 - 1.4
        r = env.step_device(a)
        rew, done = r.reward.cpu().numpy().astype(np.float64), r.done.cpu().numpy()
        ep_r += rew
        ep_l += 1
        for e in np.nonzero(done)[0]:
            rets.append(ep_r[e]); lens.append(ep_l[e])
            ep_r[e] = 0; ep_l[e] = 0
        if t == 59:
            s = env.episode_stats(reset=True).cpu().numpy()
            assert s[2] == len(rets) and s[1] == sum(lens) and abs(s[0] - sum(rets)) <= 1e-3 * max(1.0, abs(sum(rets)))
            rets, lens = [], []
    s = env.episode_stats(reset=False).cpu().numpy()
    assert len(rets) > 50 and s[2] == len(rets) and s[1] == sum(lens)
    assert abs(s[0] - sum(rets)) <= 1e-3 * max(1.0, abs(sum(rets)))
    assert np.array_equal(env.episode_stats(reset=True).cpu().numpy(), s)      # reset=False left them in place
    assert env.episode_stats().cpu().numpy().tolist() == [0.0, 0.0, 0.0]
    env.close()
