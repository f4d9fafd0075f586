"""VecEnv protocol on top of `BatchAviary` — what `MAPPO.train_step` calls.

Mirrors the working (SubprocVecEnv) route of the reference's vectorisation layer:
`make_vec_envs` (`safe_control_gym/envs/env_wrappers/vectorized_env/__init__.py:42-66`),
`VecEnv` (`vec_env.py:12-160`), `SubprocVecEnv.step/reset` and the worker's
reset-on-done (`subproc_vec_env.py:51-73,186-207`), and
`VecRecordEpisodeStatistics` (`record_episode_statistics.py:100-171`).

`step(actions)` takes/returns numpy arrays and the reference's 4-tuple
`(obs (N,M,D), rew (N,), done (N,), {'n': infos})`; one `bd_step_host_compact` call does
the work (terminal observations only for the envs that finished, info dicts built lazily).
For rollouts that stay on the GPU use `BatchAviary.step_device` (device tensors, no infos).
"""
from __future__ import annotations

import functools
from collections import deque
from copy import deepcopy

import numpy as np

from .batch_aviary import BatchAviary

_TASK_OF_CLASS = {"HoverAviary": "hover", "MultiHoverAviary": "multihover", "SpiralFormationAviary": "spiral",
                  "MeetupAviary": "meetup", "FlockAviary": "flock", "LeaderFollowerAviary": "leaderfollower"}


class LazyInfos:
    """`info['n']`: a sequence of N per-env info dicts that are built when somebody looks at them.

    `SubprocVecEnv.step` returns a tuple of N dicts (`subproc_vec_env.py:58-64`); at 65 536 envs building them costs
    more than the simulation step.  Here only the envs that finished get their dict up front (it carries
    `terminal_observation` / `terminal_info`); the others are created on first access and then cached, so writes
    such as `info['n'][i]['episode'] = ...` (`record_episode_statistics.py:158`) stick."""

    def __init__(self, n, make, ready=None):
        self._n, self._make = int(n), make
        self._cache = dict(ready) if ready else {}
        self._patches = []          # (env indices as a set, key, value factory): applied when a dict is built

    def patch(self, indices, key, factory):
        """`info['n'][i][key] = factory(i)` for every i in `indices`, without building the dicts now."""
        members = set(int(i) for i in indices)
        for i in members & self._cache.keys():
            self._cache[i][key] = factory(i)
        self._patches.append((members, key, factory))

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        d = self._cache.get(i)
        if d is None:
            d = self._cache[i] = self._make(i)
            for members, key, factory in self._patches:
                if i in members:
                    d[key] = factory(i)
        return d

    def __iter__(self):
        return (self[i] for i in range(self._n))


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
        self._flip = 0
        self._pin_act = None

    def action_buffer(self):
        """A page-locked (N,M,A) array.  `step(buf)` with exactly this array hands it to the copy engine as it is;
        any other array is first copied into pinned staging (what a caller that cannot change its allocation pays)."""
        if self._pin_act is None:
            b = self.batch
            np_act = np.float32 if str(b.action_dtype).endswith("float32") else np.float64
            self._pin_act = b.pinned_array((self.num_envs, b.NUM_DRONES, b.ACTION_DIM), np_act)
        return self._pin_act

    # -- info dicts (HoverAviary.py:119-131, MultiHoverAviary.py:274-285, SpiralAviary.py:200-205)
    def _info(self, kin=None, terminated=False, step_counter=0):
        b = self.batch
        if b.task == "spiral":
            return {"time": step_counter / b.PYB_FREQ, "omega": b.OMEGA, "radius": b.R}
        if b.task != "multihover":
            return {"answer": 42}
        reasons = []
        if terminated and kin is not None:
            k5 = np.asarray(kin[:, :5], dtype=np.float64)
            crashed = k5[:, 2] < 0.03
            flipped = (np.abs(k5[:, 3]) > 1.2) | (np.abs(k5[:, 4]) > 1.2)
            out = (np.abs(k5[:, 0]) > 3.0) | (np.abs(k5[:, 1]) > 3.0)
            for i in np.flatnonzero(crashed | flipped | out):     # same order as MultiHoverAviary.py:221-239
                x, y, z, roll, pitch = k5[i].tolist()
                if crashed[i]:
                    reasons.append(f"Drone {i} crashed (z={z:.2f})")
                if flipped[i]:
                    reasons.append(f"Drone {i} flipped (roll={roll:.2f}, pitch={pitch:.2f})")
                if out[i]:
                    reasons.append(f"Drone {i} out of bounds (pos=[{x:.2f}, {y:.2f}, {z:.2f}])")
        return {"answer": 42, "termination_reasons": reasons}

    def reset(self):
        """-> (obs (N,M,D), {'n': infos}) (subproc_vec_env.py:66-73)."""
        self._assert_not_closed()
        obs = self.batch.reset_device().cpu().numpy()
        self._steps[:] = 0
        return obs, {'n': LazyInfos(self.num_envs, lambda e: self._info())}

    def step_async(self, actions):
        self._assert_not_closed()
        self._pending = actions if actions is self._pin_act else np.asarray(actions)
        self.waiting = True

    def step_wait(self):
        """One `bd_step_host_compact` call: observations, rewards and flags come back in page-locked buffers (two
        sets, used alternately: what a call returns stays valid until the next-but-one call), terminal observations
        only for the envs that finished; per-env Python work is done for those envs only."""
        self._assert_not_closed()
        b = self.batch
        self._flip ^= 1
        res = b.step_host(self._pending, compact_terminal_obs=b.auto_reset, buffer_set=self._flip,
                          actions_pinned=self._pending is self._pin_act)
        self.waiting = False
        obs, term, trunc = res["obs"], res["terminated"], res["truncated"]
        rews = res["reward"].astype(np.float64)
        dones = np.logical_or(term, trunc)
        S = b.PYB_STEPS_PER_CTRL
        steps_before = self._steps
        track = b.task == "spiral"          # only the Spiral info dict reports the step counter ("time")
        if not track:
            idx, rows = (res["done_idx"], res["terminal_rows"]) if b.auto_reset else (None, None)
        elif b.auto_reset:
            # compact terminal rows: handle-owned page-locked memory, two sets used alternately, so — like the
            # observations — they stay valid until the next-but-one step; a dict built from them copies its row
            idx, rows = res["done_idx"], res["terminal_rows"]
            self._steps = np.where(dones, 0, steps_before + S)
        else:
            idx, rows = None, None
            self._steps = steps_before + S

        def make(e, obs=obs, term=term, steps=steps_before, idx=idx, rows=rows):
            if idx is not None and dones[e]:
                k = int(np.searchsorted(idx, e))
                info = self._info()   # info of the fresh episode (reset)
                info['terminal_observation'] = rows[k].copy()
                info['terminal_info'] = self._info(rows[k], bool(term[e]), int(steps[e]))
                return info
            return self._info(obs[e], bool(term[e]), int(steps[e]))
        return obs, rews, dones, {'n': LazyInfos(self.num_envs, make)}

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
        """The workers' RNG states (`subproc_vec_env.py:101-106`, checkpointed by `mappo.py:203-229`).  There are no
        workers: ONE state, the batched simulator's Philox key and counters (`bd_get_rng_state`), plus numpy's
        global state, which the single-env views draw their re-spawn jitter from like the reference does."""
        return [{"philox": self.batch.get_rng_state(), "numpy": np.random.get_state()}]

    def set_env_random_state(self, worker_random_states):
        """`subproc_vec_env.py:108-112`: restore what `get_env_random_state` returned."""
        states = list(worker_random_states)
        if len(states) != 1 or "philox" not in states[0]:
            raise ValueError("set_env_random_state expects the one-element list returned by get_env_random_state")
        self.batch.set_rng_state(states[0]["philox"])
        if states[0].get("numpy") is not None:
            np.random.set_state(states[0]["numpy"])

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
    """Running return / length of every env's current episode, queues of the finished ones, and optional named
    trackers fed from the info dicts — the protocol of `record_episode_statistics.py:100-171` (attribute names
    `return_queue`, `length_queue`, `episode_return`, `episode_length`, `add_tracker`, `accumulated_stats`,
    `queued_stats`; `info['n'][i]['episode'] = {'r', 'l', <tracker>...}` on the step an episode ends).

    Bookkeeping is vectorised: per-env Python work happens only for the envs that finished, and — when trackers are
    registered — for the envs whose info dict carries a tracked key."""

    def __init__(self, venv, deque_size=None, **kwargs):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.deque_size = deque_size
        self.return_queue, self.length_queue = deque(maxlen=deque_size), deque(maxlen=deque_size)
        self.episode_return = np.zeros(self.num_envs, dtype=np.float64)
        self.episode_length = np.zeros(self.num_envs, dtype=np.float64)
        self._trackers = {}            # name -> (mode, initial value, per-env running values)
        self.accumulated_stats, self.queued_stats = {}, {}

    # the reference exposes the per-env tracker values as `episode_stats[name][i]`
    @property
    def episode_stats(self):
        return {name: t[2] for name, t in self._trackers.items()}

    def add_tracker(self, name, init_value, mode='accumulate'):
        if mode == 'accumulate':
            self.accumulated_stats[name] = init_value
        elif mode == 'queue':
            self.queued_stats[name] = deque(maxlen=self.deque_size)
        else:
            raise Exception('Tracker mode not implemented.')
        self._trackers[name] = (mode, init_value, [deepcopy(init_value) for _ in range(self.num_envs)])

    def _clear_tracker(self, name, i):
        mode, init, values = self._trackers[name]
        values[i] = values[i] * 0      # keeps the value's type / shape, like the reference's `*= 0`

    def reset(self, **kwargs):
        self.episode_return[:] = 0.0
        self.episode_length[:] = 0.0
        for name in self._trackers:
            for i in range(self.num_envs):
                self._clear_tracker(name, i)
        return self.venv.reset(**kwargs)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, reward, done, info = self.venv.step_wait()
        r = np.asarray(reward, dtype=np.float64)
        self.episode_return += r if r.ndim == 1 else r.reshape(self.num_envs, -1).mean(axis=1)   # mean over agents (:148)
        self.episode_length += 1
        done = np.asarray(done, dtype=bool)
        infos = info['n']
        finished = np.flatnonzero(done)
        if self._trackers:             # tracked keys may sit in any env's info (or in its terminal_info when it ended)
            for i in range(self.num_envs):
                src = infos[i]
                if done[i] and 'terminal_info' in src:
                    src = src['terminal_info']
                for name, (_, _, values) in self._trackers.items():
                    if name in src:
                        values[i] = values[i] + src[name]
        if finished.size:
            ep_r, ep_l = self.episode_return[finished].tolist(), self.episode_length[finished].tolist()
            self.return_queue.extend(ep_r)
            self.length_queue.extend(ep_l)
            if not self._trackers and hasattr(infos, "patch"):
                # lazily built info dicts: the episode record is attached when somebody reads the dict
                pos = dict(zip(finished.tolist(), range(finished.size)))
                infos.patch(pos.keys(), 'episode', lambda i: {'r': ep_r[pos[i]], 'l': ep_l[pos[i]]})
            else:
                for k, i in enumerate(finished.tolist()):
                    ep = {'r': ep_r[k], 'l': ep_l[k]}
                    for name, (mode, _, values) in self._trackers.items():
                        ep[name] = deepcopy(values[i])
                        if mode == 'accumulate':
                            self.accumulated_stats[name] = self.accumulated_stats[name] + deepcopy(values[i])
                        else:
                            self.queued_stats[name].append(deepcopy(values[i]))
                        self._clear_tracker(name, i)
                    infos[i]['episode'] = ep
            self.episode_return[finished] = 0.0
            self.episode_length[finished] = 0.0
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
