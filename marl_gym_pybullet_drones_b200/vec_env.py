"""VecEnv protocol on top of `BatchAviary` — what `MAPPO.train_step` calls.

Mirrors the working (SubprocVecEnv) route of the reference's vectorisation layer:
`make_vec_envs` (`safe_control_gym/envs/env_wrappers/vectorized_env/__init__.py:42-66`),
`VecEnv` (`vec_env.py:12-160`), `SubprocVecEnv.step/reset` and the worker's
reset-on-done (`subproc_vec_env.py:51-73,186-207`), and
`VecRecordEpisodeStatistics` (`record_episode_statistics.py:100-171`).

`step(actions)` takes/returns numpy arrays and the reference's 4-tuple
`(obs (N,M,D), rew (N,), done (N,), {'n': infos})`; one `bd_step_host` call does
the work.  For large N use `BatchAviary.step_device` (device tensors, no infos).
"""
from __future__ import annotations

import functools
from collections import deque
from copy import deepcopy

import numpy as np

from .batch_aviary import BatchAviary

_TASK_OF_CLASS = {"HoverAviary": "hover", "MultiHoverAviary": "multihover", "SpiralFormationAviary": "spiral",
                  "MeetupAviary": "meetup", "FlockAviary": "flock", "LeaderFollowerAviary": "leaderfollower"}


class BatchVecEnv:
    """N aviaries behind the reference's VecEnv interface, stepped by one kernel launch."""

    closed = False
    viewer = None
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, batch: BatchAviary):
        self.batch = batch
        self.num_envs = batch.num_envs
        self.observation_space = batch.observation_space
        self.action_space = batch.action_space
        self.waiting = False
        self._pending = None
        self._steps = np.zeros(self.num_envs, dtype=np.int64)   # step_counter mirror for info dicts

    # -- info dicts (HoverAviary.py:119-131, MultiHoverAviary.py:274-285, SpiralAviary.py:200-205)
    def _info(self, e, kin=None, terminated=False, step_counter=0):
        b = self.batch
        if b.task == "hover":
            return {"answer": 42}
        if b.task == "spiral":
            return {"time": step_counter / b.PYB_FREQ, "omega": b.OMEGA, "radius": b.R}
        reasons = []
        if terminated and kin is not None:
            for i in range(b.NUM_DRONES):
                x, y, z, roll, pitch = kin[i, 0], kin[i, 1], kin[i, 2], kin[i, 3], kin[i, 4]
                if z < 0.03:
                    reasons.append(f"Drone {i} crashed (z={z:.2f})")
                if abs(roll) > 1.2 or abs(pitch) > 1.2:
                    reasons.append(f"Drone {i} flipped (roll={roll:.2f}, pitch={pitch:.2f})")
                if abs(x) > 3.0 or abs(y) > 3.0:
                    reasons.append(f"Drone {i} out of bounds (pos=[{x:.2f}, {y:.2f}, {z:.2f}])")
        return {"answer": 42, "termination_reasons": reasons}

    def reset(self):
        """-> (obs (N,M,D), {'n': infos}) (subproc_vec_env.py:66-73)."""
        self._assert_not_closed()
        obs = self.batch.reset_device().cpu().numpy()
        self._steps[:] = 0
        return obs, {'n': tuple(self._info(e) for e in range(self.num_envs))}

    def step_async(self, actions):
        self._assert_not_closed()
        self._pending = np.asarray(actions)
        self.waiting = True

    def step_wait(self):
        self._assert_not_closed()
        b = self.batch
        res = b.step_host(self._pending, want_terminal_obs=b.auto_reset)
        self.waiting = False
        obs = res["obs"].copy()
        rews = res["reward"].astype(np.float64)
        term, trunc = res["terminated"].copy(), res["truncated"].copy()
        dones = np.logical_or(term, trunc)
        infos = []
        S = b.PYB_STEPS_PER_CTRL
        for e in range(self.num_envs):
            if dones[e] and b.auto_reset:
                end_obs = res["terminal_obs"][e].copy()
                end_info = self._info(e, end_obs, bool(term[e]), int(self._steps[e]))
                info = self._info(e)   # info of the fresh episode (reset)
                info['terminal_observation'] = end_obs
                info['terminal_info'] = end_info
                self._steps[e] = 0
            else:
                info = self._info(e, obs[e], bool(term[e]), int(self._steps[e]))
                self._steps[e] += S
            infos.append(info)
        return obs, rews, dones, {'n': tuple(infos)}

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # -- the rest of the VecEnv surface (subproc_vec_env.py:84-158) ---------------------
    def get_attr(self, attr_name, indices=None):
        idx = self._get_indices(indices)
        return [getattr(self.batch, attr_name) for _ in idx]

    def set_attr(self, attr_name, values, indices=None):
        raise NotImplementedError("per-env attributes cannot be set on a batched aviary")

    def env_method(self, method_name, method_args=None, method_kwargs=None, indices=None):
        raise NotImplementedError("per-env methods are not available on a batched aviary; use .batch")

    def get_env_random_state(self):
        """Stand-in for the workers' RNG states (`mappo.py:203-229` checkpoints them)."""
        return [{"philox_seed": int(self.batch._cfg.seed)}]

    def set_env_random_state(self, worker_random_states):
        return None

    def _get_indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    def get_images(self):
        raise NotImplementedError("rendering needs PyBullet (out of scope)")

    def render(self, mode='human'):
        raise NotImplementedError("rendering needs PyBullet (out of scope)")

    @property
    def unwrapped(self):
        return self

    def close(self):
        if self.closed:
            return
        self.batch.close()
        self.closed = True

    def _assert_not_closed(self):
        assert not self.closed, 'Trying to operate on a BatchVecEnv after calling close()'


class VecRecordEpisodeStatistics:
    """Episode returns / lengths per env (record_episode_statistics.py:100-171)."""

    def __init__(self, venv, deque_size=None, **kwargs):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space = venv.observation_space
        self.action_space = venv.action_space
        self.deque_size = deque_size
        self.episode_return = np.zeros(self.num_envs)
        self.episode_length = np.zeros(self.num_envs)
        self.return_queue = deque(maxlen=deque_size)
        self.length_queue = deque(maxlen=deque_size)
        self.episode_stats = {}
        self.accumulated_stats = {}
        self.queued_stats = {}

    def add_tracker(self, name, init_value, mode='accumulate'):
        self.episode_stats[name] = [init_value for _ in range(self.num_envs)]
        if mode == 'accumulate':
            self.accumulated_stats[name] = init_value
        elif mode == 'queue':
            self.queued_stats[name] = deque(maxlen=self.deque_size)
        else:
            raise Exception('Tracker mode not implemented.')

    def reset(self, **kwargs):
        self.episode_return = np.zeros(self.num_envs)
        self.episode_length = np.zeros(self.num_envs)
        for key in self.episode_stats:
            for i in range(self.num_envs):
                self.episode_stats[key][i] *= 0
        return self.venv.reset(**kwargs)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, reward, done, info = self.venv.step_wait()
        self.episode_return += np.asarray([float(np.mean(r)) for r in reward])   # :148
        self.episode_length += 1
        for i in np.nonzero(np.ones(self.num_envs, dtype=bool) if self.episode_stats else done)[0]:
            d = bool(done[i])
            inf = info['n'][i]['terminal_info'] if (d and 'terminal_info' in info['n'][i]) else info['n'][i]
            for key in self.episode_stats:
                if key in inf:
                    self.episode_stats[key][i] += inf[key]
            if d:
                info['n'][i]['episode'] = {'r': self.episode_return[i], 'l': self.episode_length[i]}
                self.return_queue.append(deepcopy(self.episode_return[i]))
                self.length_queue.append(deepcopy(self.episode_length[i]))
                self.episode_return[i] = 0
                self.episode_length[i] = 0
                for key in self.episode_stats:
                    info['n'][i]['episode'][key] = deepcopy(self.episode_stats[key][i])
                    if key in self.accumulated_stats:
                        self.accumulated_stats[key] += deepcopy(self.episode_stats[key][i])
                    if key in self.queued_stats:
                        self.queued_stats[key].append(deepcopy(self.episode_stats[key][i]))
                    self.episode_stats[key][i] *= 0
        return obs, reward, done, info

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        return self.venv.close()

    @property
    def unwrapped(self):
        return self.venv.unwrapped

    def __getattr__(self, name):
        return getattr(self.venv, name)


def _spec_from_env_func(env_func):
    """Extract (task, kwargs) from `functools.partial(<aviary class>, **kwargs)` or a class."""
    kwargs = {}
    f = env_func
    while isinstance(f, functools.partial):
        kwargs = {**f.keywords, **kwargs}
        f = f.func
    name = getattr(f, "__name__", "")
    if name in _TASK_OF_CLASS:
        return _TASK_OF_CLASS[name], kwargs
    spec = getattr(env_func, "batch_spec", None)
    if spec is not None:
        return spec["task"], {k: v for k, v in spec.items() if k != "task"}
    raise TypeError(
        "make_vec_envs needs env_func to be one of the aviary classes of this package (optionally "
        "wrapped in functools.partial), or a callable with a `batch_spec` dict attribute: the batched simulator "
        "constructs all envs at once instead of calling env_func N times")


def make_vec_envs(env_func, env_configs=None, batch_size=1, n_processes=1, seed=None, **batch_kwargs):
    """Reference signature (`vectorized_env/__init__.py:42-66`); `n_processes` is accepted and
    ignored (there are no worker processes).  Returns a `BatchVecEnv`."""
    task, kwargs = _spec_from_env_func(env_func)
    kwargs = dict(kwargs)
    kwargs.pop("seed", None)
    kwargs.pop("gui", None)
    kwargs.pop("record", None)
    opts = dict(precision="fp32", auto_reset=True, seed=0 if seed is None else int(seed))
    opts.update(batch_kwargs)
    batch = BatchAviary(task=task, num_envs=batch_size, **kwargs, **opts)
    return BatchVecEnv(batch)
