"""Gymnasium-protocol single-environment views of the GPU simulator.

`HoverAviary`, `MultiHoverAviary` and `SpiralFormationAviary` keep the
reference's constructor keywords, `reset(seed, options) -> (obs, info)` and
`step(action) -> (obs, reward, terminated, truncated, info)` 5-tuple
(`BaseAviary.py:220-255, 259-383`), returning numpy arrays, so scripts written
against one reference env (`examples/learn.py`, `MAPPO.run` evaluation,
`mappo/mappo.py:534-581`) run unchanged.  Each is a `BatchAviary` with N = 1 and
`auto_reset=False`; every number comes from the CUDA kernels.

Differences that are visible on purpose:
* `physics` defaults to `Physics.DYN` (the reference default `Physics.PYB`
  needs PyBullet's solver and raises `NotImplementedError` here);
* `ActionType.PID / VEL / ONE_D_PID` run the reference's DSL PID controller inside the step
  kernel (one controller per drone, never reset by `reset()`, as in `BaseRLAviary.py:73-78`);
* `obs` is always float32 (the reference returns float64 until the action
  buffer has filled with float32 actions, `BaseRLAviary.py:315-318`).
"""
from __future__ import annotations

import numpy as np
import torch

from .batch_aviary import BatchAviary
from .enums import ActionType, DroneModel, ObservationType, Physics
from .spaces import Env


class _SingleAviary(Env):
    TASK = "multihover"

    def __init__(self, drone_model=DroneModel.CF2X, num_drones=1, neighbourhood_radius=np.inf,
                 initial_xyzs=None, initial_rpys=None, physics=Physics.DYN, pyb_freq=240, ctrl_freq=30,
                 gui=False, record=False, obs=ObservationType.KIN, act=ActionType.RPM,
                 precision="fp64", device=None, **task_kwargs):
        self._batch = BatchAviary(task=self.TASK, num_envs=1, drone_model=drone_model, num_drones=num_drones,
                                  neighbourhood_radius=neighbourhood_radius, initial_xyzs=initial_xyzs,
                                  initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq,
                                  ctrl_freq=ctrl_freq, gui=gui, record=record, obs=obs, act=act,
                                  precision=precision, device=device, auto_reset=False, reset_mode="fixed",
                                  keep_ang_vel=True, **task_kwargs)
        b = self._batch
        for name in ("NUM_DRONES", "CTRL_FREQ", "PYB_FREQ", "PYB_STEPS_PER_CTRL", "CTRL_TIMESTEP", "PYB_TIMESTEP",
                     "EPISODE_LEN_SEC", "ACTION_BUFFER_SIZE", "DRONE_MODEL", "PHYSICS", "OBS_TYPE", "ACT_TYPE",
                     "M", "L", "KF", "KM", "J", "J_INV", "G", "GRAVITY", "HOVER_RPM", "MAX_RPM", "MAX_THRUST",
                     "MAX_XY_TORQUE", "MAX_Z_TORQUE", "GND_EFF_COEFF", "PROP_RADIUS", "GND_EFF_H_CLIP",
                     "DRAG_COEFF", "DW_COEFF_1", "DW_COEFF_2", "DW_COEFF_3", "THRUST2WEIGHT_RATIO",
                     "COLLISION_H", "COLLISION_R", "COLLISION_Z_OFFSET", "MAX_SPEED_KMH",
                     "NEIGHBOURHOOD_RADIUS", "observation_space", "action_space"):
            setattr(self, name, getattr(b, name))
        self.INIT_XYZS = np.array(b.INIT_XYZS, dtype=np.float64).reshape(self.NUM_DRONES, 3)
        self.INIT_RPYS = np.array(b.INIT_RPYS, dtype=np.float64).reshape(self.NUM_DRONES, 3)
        self.step_counter = 0
        self.termination_reasons = []

    # -------------------------------------------------------------- protocol
    def reset(self, seed: int = None, options: dict = None):
        """BaseAviary.reset (:220-255); `seed` is ignored exactly like the reference (:243)."""
        self._pre_reset()
        obs = self._batch.reset_device()
        self.step_counter = 0
        return obs[0].cpu().numpy(), self._computeInfo()

    def step(self, action):
        """BaseAviary.step (:259-383): 5-tuple, no auto-reset."""
        action = np.asarray(action)
        if action.dtype not in (np.float32, np.float64):
            action = action.astype(np.float64)
        if self._batch.precision == "fp32":
            action = action.astype(np.float32)
        # the action's own dtype is kept: numpy computes the rpm partly in float32 for float32 actions
        a = torch.as_tensor(np.ascontiguousarray(action.reshape(1, self.NUM_DRONES, -1))).to(self._batch.device)
        res = self._batch.step_device(a)
        obs = res.obs[0].cpu().numpy()
        reward = float(res.reward[0].item())
        terminated = bool(res.terminated[0].item())
        truncated = bool(res.truncated[0].item())
        info = self._computeInfo(terminated)
        self.step_counter += self.PYB_STEPS_PER_CTRL
        return obs, reward, terminated, truncated, info

    def _pre_reset(self):
        pass

    def _computeInfo(self, terminated=False):
        return {"answer": 42}

    # ---------------------------------------------------------- state access
    def _states(self) -> np.ndarray:
        return self._batch.get_state()[0].cpu().numpy().astype(np.float64)

    def _getDroneStateVector(self, nth_drone: int) -> np.ndarray:
        """(20,) [pos quat rpy vel ang_v last_rpm] (BaseAviary.py:541-561)."""
        return self._states()[nth_drone].reshape(20,)

    @property
    def pos(self):
        return self._states()[:, 0:3]

    @property
    def quat(self):
        return self._states()[:, 3:7]

    @property
    def rpy(self):
        return self._states()[:, 7:10]

    @property
    def vel(self):
        return self._states()[:, 10:13]

    @property
    def ang_v(self):
        return self._states()[:, 13:16]

    def render(self, mode="human", close=False):
        s = self._states()
        for i in range(self.NUM_DRONES):   # same content as BaseAviary.render's per-drone line (:407-412)
            print("[INFO] BaseAviary.render() ——— it {:04d}".format(self.step_counter),
                  "——— drone {:d}".format(i),
                  "——— x {:+06.2f}, y {:+06.2f}, z {:+06.2f}".format(*s[i, 0:3]),
                  "——— velocity {:+06.2f}, {:+06.2f}, {:+06.2f}".format(*s[i, 10:13]))

    def getPyBulletClient(self):
        raise NotImplementedError("there is no PyBullet client behind the GPU simulator")

    def close(self):
        self._batch.close()


class HoverAviary(_SingleAviary):
    """Single-drone hover task (reference `envs/HoverAviary.py`)."""

    TASK = "hover"

    def __init__(self, drone_model=DroneModel.CF2X, initial_xyzs=None, initial_rpys=None,
                 physics=Physics.DYN, pyb_freq=240, ctrl_freq=30, gui=False, record=False,
                 obs=ObservationType.KIN, act=ActionType.RPM, precision="fp64", device=None):
        self.TARGET_POS = np.array([0, 0, 1])
        super().__init__(drone_model=drone_model, num_drones=1, initial_xyzs=initial_xyzs,
                         initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq,
                         gui=gui, record=record, obs=obs, act=act, precision=precision, device=device)


class MultiHoverAviary(_SingleAviary):
    """Multi-drone hover task (reference `envs/MultiHoverAviary.py`)."""

    TASK = "multihover"

    def __init__(self, drone_model=DroneModel.CF2X, num_drones=2, neighbourhood_radius=np.inf,
                 initial_xyzs=None, initial_rpys=None, physics=Physics.DYN, pyb_freq=240, ctrl_freq=30,
                 gui=False, record=False, obs=ObservationType.KIN, act=ActionType.RPM,
                 precision="fp64", device=None):
        super().__init__(drone_model=drone_model, num_drones=num_drones,
                         neighbourhood_radius=neighbourhood_radius, initial_xyzs=initial_xyzs,
                         initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq,
                         gui=gui, record=record, obs=obs, act=act, precision=precision, device=device)
        self.TARGET_POS = self.INIT_XYZS + np.array([[0, 0, 1 / (i + 1)] for i in range(self.NUM_DRONES)])

    def _pre_reset(self):
        """MultiHoverAviary.reset (:75-110): host-side jitter with the process-global np.random,
        so a script that seeds numpy sees the reference's own draw sequence."""
        if not hasattr(self, "ORIGINAL_INIT_XYZS"):
            self.ORIGINAL_INIT_XYZS = self.INIT_XYZS.copy()
        M = self.NUM_DRONES
        while True:
            cand = self.ORIGINAL_INIT_XYZS.copy() + np.random.uniform(-0.25, 0.25, (M, 3))
            cand[:, 2] = np.clip(cand[:, 2], 0.1, 1.0)
            dists = np.linalg.norm(cand[:, np.newaxis, :] - cand[np.newaxis, :, :], axis=2)
            np.fill_diagonal(dists, np.inf)
            if not np.any(dists < 0.5) and not np.any(cand[:, 2] < 0.1):
                break
        self.INIT_XYZS = cand
        self.TARGET_POS = self.INIT_XYZS + np.array([[0, 0, 1 / (i + 1)] for i in range(M)])
        self.termination_reasons = []
        self._batch.set_initial_poses(self.INIT_XYZS, self.INIT_RPYS)

    def _computeInfo(self, terminated=False):
        """MultiHoverAviary._computeTerminated's reason strings (:220-240) + _computeInfo (:274-285)."""
        reasons = []
        if terminated:
            s = self._states()
            for i in range(self.NUM_DRONES):
                x, y, z, roll, pitch = s[i, 0], s[i, 1], s[i, 2], s[i, 7], s[i, 8]
                if z < 0.03:
                    reasons.append(f"Drone {i} crashed (z={z:.2f})")
                if abs(roll) > 1.2 or abs(pitch) > 1.2:
                    reasons.append(f"Drone {i} flipped (roll={roll:.2f}, pitch={pitch:.2f})")
                if abs(x) > 3.0 or abs(y) > 3.0:
                    reasons.append(f"Drone {i} out of bounds (pos=[{x:.2f}, {y:.2f}, {z:.2f}])")
        self.termination_reasons = reasons
        return {"answer": 42, "termination_reasons": self.termination_reasons}


class SpiralFormationAviary(_SingleAviary):
    """Analytic spiral-formation tracking task (reference `envs/SpiralAviary.py`)."""

    TASK = "spiral"

    def __init__(self, drone_model=DroneModel.CF2X, num_drones=3, neighbourhood_radius=np.inf,
                 initial_xyzs=None, initial_rpys=None, physics=Physics.DYN, pyb_freq=240, ctrl_freq=48,
                 gui=False, record=False, obs=ObservationType.KIN, act=ActionType.VEL,
                 spiral_radius=0.4, spiral_period=10.0, height_rate=0.05,
                 target_center=np.array([0.0, 0.0, 0.0]), precision="fp64", device=None):
        super().__init__(drone_model=drone_model, num_drones=num_drones,
                         neighbourhood_radius=neighbourhood_radius, initial_xyzs=initial_xyzs,
                         initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq,
                         gui=gui, record=record, obs=obs, act=act, precision=precision, device=device,
                         spiral_radius=spiral_radius, spiral_period=spiral_period, height_rate=height_rate,
                         target_center=tuple(float(v) for v in target_center))
        self.R, self.PERIOD, self.VZ = spiral_radius, spiral_period, height_rate
        self.OMEGA = 2 * np.pi / self.PERIOD
        self.CENTER = np.array(target_center, dtype=np.float64)

    def _computeInfo(self, terminated=False):
        return {"time": self.step_counter / self.PYB_FREQ, "omega": self.OMEGA, "radius": self.R}


class _SwarmAviary(_SingleAviary):
    """Shared constructor of the three swarm tasks (`num_drones=2`, RPM action, 8 s episodes:
    reference `envs/MeetupAviary.py:12-69`, `FlockAviary.py:13-69`, `LeaderFollowerAviary.py:11-68`)."""

    def __init__(self, drone_model=DroneModel.CF2X, num_drones=2, neighbourhood_radius=np.inf,
                 initial_xyzs=None, initial_rpys=None, physics=Physics.DYN, pyb_freq=240, ctrl_freq=30,
                 gui=False, record=False, obs=ObservationType.KIN, act=ActionType.RPM,
                 precision="fp64", device=None):
        super().__init__(drone_model=drone_model, num_drones=num_drones,
                         neighbourhood_radius=neighbourhood_radius, initial_xyzs=initial_xyzs,
                         initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq,
                         gui=gui, record=record, obs=obs, act=act, precision=precision, device=device)


class MeetupAviary(_SwarmAviary):
    """Pairs (i, M-1-i) meet mid-flight: reward -2 |p_i - p_partner|^2 per pair, terminated when every
    pair is within 0.1 m (reference `envs/MeetupAviary.py:74-154`)."""

    TASK = "meetup"


class FlockAviary(_SwarmAviary):
    """Flocking: velocity alignment + flock speed - spacing penalty - spacing variance
    (reference `envs/FlockAviary.py:75-189`)."""

    TASK = "flock"


class LeaderFollowerAviary(_SwarmAviary):
    """Drone 0 hovers at (0, 0, 0.5), the others match its height (reference
    `envs/LeaderFollowerAviary.py:72-145`)."""

    TASK = "leaderfollower"
