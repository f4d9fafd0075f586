"""Host-side logic that needs no GPU: spaces shim, loud failure without CUDA, env_func parsing,
episode statistics wrapper (against a scripted fake venv), env sharding."""
import functools

import numpy as np
import pytest
import torch

from marl_gym_pybullet_drones_b200 import Physics
from marl_gym_pybullet_drones_b200.dist import shard_envs
from marl_gym_pybullet_drones_b200.spaces import Box


def test_box_shim_surface():
    b = Box(low=-np.ones((2, 4), dtype=np.float32), high=np.ones((2, 4), dtype=np.float32), dtype=np.float32)
    assert b.shape == (2, 4) and b.dtype == np.float32
    s = b.sample()
    assert s.shape == (2, 4) and b.contains(s) and not b.contains(np.full((2, 4), 3.0, dtype=np.float32))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_batch_aviary_fails_loudly_without_cuda():
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchAviary(task="multihover", num_envs=2, num_drones=2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_trainer_kernels_fail_loudly_without_cuda():
    """The PPO-update front end and the fused actor have no CPU path either: they raise, they do not fall back."""
    from marl_gym_pybullet_drones_b200.ppo_native import PpoNet
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PpoNet(72, 1, 4, True, 128)
    from marl_gym_pybullet_drones_b200.actor import FusedActor
    with pytest.raises(Exception):
        FusedActor(72, 256, 4)


def test_unsupported_modes_raise_before_touching_the_device():
    from marl_gym_pybullet_drones_b200.batch_aviary import BatchAviary
    with pytest.raises(ValueError, match="pyb_freq is not divisible"):
        BatchAviary(pyb_freq=240, ctrl_freq=50)
    with pytest.raises(ValueError, match="no controller is available"):      # BaseRLAviary.py:73-78
        BatchAviary(act="vel", drone_model="racer")
    with pytest.raises(NotImplementedError):
        BatchAviary(obs="rgb")
    with pytest.raises(NotImplementedError):
        BatchAviary(gui=True)
    if not torch.cuda.is_available():
        with pytest.raises((NotImplementedError, RuntimeError)):
            BatchAviary(physics=Physics.PYB)


def test_env_func_spec_extraction():
    from marl_gym_pybullet_drones_b200.envs import MultiHoverAviary, SpiralFormationAviary
    from marl_gym_pybullet_drones_b200.vec_env import _spec_from_env_func
    f = functools.partial(functools.partial(MultiHoverAviary, num_drones=3), ctrl_freq=48, num_drones=2)
    assert _spec_from_env_func(f) == ("multihover", {"num_drones": 2, "ctrl_freq": 48})
    assert _spec_from_env_func(SpiralFormationAviary) == ("spiral", {})
    with pytest.raises(TypeError):
        _spec_from_env_func(lambda: None)


class _FakeVenv:
    """Scripted venv: env 0 finishes every 3rd step, env 1 never."""

    num_envs, observation_space, action_space = 2, None, None

    def __init__(self):
        self.t = 0

    def reset(self):
        self.t = 0
        return np.zeros((2, 1, 3)), {'n': ({}, {})}

    def step_async(self, a):
        pass

    def step_wait(self):
        self.t += 1
        done = np.array([self.t % 3 == 0, False])
        infos = [{"cost": 1.0}, {"cost": 2.0}]
        if done[0]:
            infos[0] = {"terminal_info": {"cost": 5.0}, "terminal_observation": np.ones((1, 3))}
        return np.zeros((2, 1, 3)), np.array([1.0, 0.5]), done, {'n': tuple(infos)}


def test_vec_record_episode_statistics_matches_reference_semantics():
    from marl_gym_pybullet_drones_b200.vec_env import VecRecordEpisodeStatistics
    env = VecRecordEpisodeStatistics(_FakeVenv(), deque_size=10)
    env.add_tracker("cost", 0, mode="queue")
    env.reset()
    for t in range(1, 7):
        obs, rew, done, info = env.step(None)
        if t % 3 == 0:
            ep = info['n'][0]['episode']
            assert ep['r'] == 3.0 and ep['l'] == 3 and ep['cost'] == 1.0 + 1.0 + 5.0
        else:
            assert 'episode' not in info['n'][0]
    assert list(env.return_queue) == [3.0, 3.0] and list(env.length_queue) == [3, 3]
    assert list(env.queued_stats["cost"]) == [7.0, 7.0]
    assert env.episode_return[1] == 3.0 and env.episode_length[1] == 6


def test_shard_envs_is_array_split():
    for total in (176, 65536, 7, 3):
        for world in (1, 2, 3, 8):
            parts = [shard_envs(total, r, world) for r in range(world)]
            want = np.array_split(np.arange(total), world)
            for (start, count), w in zip(parts, want):
                assert count == len(w) and (count == 0 or start == w[0])
