"""TEST INFRASTRUCTURE — fp64 numpy restatement of the reference's Physics.DYN path.

This is the parity oracle for the CUDA drone-step kernels.  It is NOT a product
code path: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
leg may import it.  It deliberately keeps the reference's structure — one Python
object per environment, per-drone / per-substep loops, tiny numpy operations —
because (a) that makes each line checkable against the file:line it restates
and (b) timing it reproduces the cost profile of the reference's CPU path
(`bench.py --impl reference`).

Pinning status: the DYN path (rows 2-8, 10-13 of SURVEY.md §8a) is PINNED against
golden trajectories produced by importing and running the unmodified reference
classes from `/root/reference` (see `tests/golden/make_golden.py`); the PyBullet
closed-form helpers underneath both (`oracle/bullet_math.py`) are UNPINNED
(restated from Bullet's sources, cross-checked against scipy).  The composition
of ground effect / drag / downwash with DYN is this repo's definition (the
reference only applies those through PyBullet's solver, `BaseAviary.py:354-367`);
the per-term formulas are pinned against the reference's own
`_groundEffect/_drag/_downwash`.

Restated reference locations (all under `gym_pybullet_drones/envs/`):
  BaseAviary.py:74-128     constants                 -> `AirframeParams`
  BaseAviary.py:194-207    default initial poses     -> `OracleAviary.__init__`
  BaseAviary.py:220-255    reset                     -> `OracleAviary.reset`
  BaseAviary.py:259-383    step                      -> `OracleAviary.step`
  BaseAviary.py:451-505    housekeeping              -> `OracleAviary._housekeeping`
  BaseAviary.py:509-519    kinematic refresh         -> `OracleAviary._refresh_kinematics`
  BaseAviary.py:541-561    20-float state vector     -> `OracleAviary.state_vector`
  BaseAviary.py:715-811    ground effect/drag/downwash formulas
  BaseAviary.py:815-892    _dynamics, _integrateQ    -> `_dynamics`, `integrate_q`
  BaseRLAviary.py:66-67,132-156,160-239,284-322  action buffer, RPM maps, KIN obs
  HoverAviary.py:51-131, MultiHoverAviary.py:58-285, SpiralAviary.py:20-205  tasks
  MeetupAviary.py:56-154, FlockAviary.py:56-189, LeaderFollowerAviary.py:55-145  swarm tasks
  safe_control_gym/envs/gym_pybullet_drones/base_aviary.py:462-511  Euler-angle integrator variant
  safe_control_gym/envs/env_wrappers/vectorized_env/subproc_vec_env.py:186-207  auto-reset
"""
import math
from collections import deque

import numpy as np

from . import bullet_math as bm
from .dsl_pid import DSLPIDOracle

AERO_GND, AERO_DRAG, AERO_DW = 1, 2, 4

# URDF facts (cf2x.urdf:5,11-12,42-78; cf2p.urdf:12,42-78; racer.urdf:5,11-12,36-72)
_AIRFRAMES = {
    "cf2x": dict(M=0.027, L=0.0397, KF=3.16e-10, KM=7.94e-12, T2W=2.25,
                 J=(1.4e-5, 1.4e-5, 2.17e-5), PROP_RADIUS=2.31348e-2, MAX_SPEED_KMH=30.0,
                 PROPS=((0.028, -0.028, 0), (-0.028, -0.028, 0), (-0.028, 0.028, 0), (0.028, 0.028, 0))),
    "cf2p": dict(M=0.027, L=0.0397, KF=3.16e-10, KM=7.94e-12, T2W=2.25,
                 J=(2.3951e-5, 2.3951e-5, 3.2347e-5), PROP_RADIUS=2.31348e-2, MAX_SPEED_KMH=30.0,
                 PROPS=((0.0397, 0, 0), (0, 0.0397, 0), (-0.0397, 0, 0), (0, -0.0397, 0))),
    "racer": dict(M=0.830, L=0.109, KF=8.47e-9, KM=2.13e-11, T2W=4.17,
                  J=(.003113, .003113, .003113), PROP_RADIUS=12.7e-2, MAX_SPEED_KMH=200.0,
                  PROPS=((0.0850, 0.0675, 0), (-0.0850, 0.0675, 0), (-0.085, -0.0675, 0), (0.085, -0.0675, 0))),
}


class AirframeParams:
    """BaseAviary.py:74-128 (constants and derived quantities), fp64."""

    def __init__(self, model="cf2x"):
        a = _AIRFRAMES[model]
        self.model = model
        self.G = 9.8
        self.M, self.L, self.KF, self.KM = a["M"], a["L"], a["KF"], a["KM"]
        self.THRUST2WEIGHT_RATIO = a["T2W"]
        self.J = np.diag(a["J"])
        self.J_INV = np.linalg.inv(self.J)
        self.COLLISION_H, self.COLLISION_R, self.COLLISION_Z_OFFSET = 0.025, 0.06, 0.0
        self.MAX_SPEED_KMH = a["MAX_SPEED_KMH"]
        self.GND_EFF_COEFF = 11.36859
        self.PROP_RADIUS = a["PROP_RADIUS"]
        self.DRAG_COEFF = np.array([9.1785e-7, 9.1785e-7, 10.311e-7])
        self.DW_COEFF_1, self.DW_COEFF_2, self.DW_COEFF_3 = 2267.18, .16, -.11
        self.PROPS = np.array(a["PROPS"], dtype=np.float64)
        self.GRAVITY = self.G * self.M                                        # :117
        self.HOVER_RPM = np.sqrt(self.GRAVITY / (4 * self.KF))                # :118
        self.MAX_RPM = np.sqrt((self.THRUST2WEIGHT_RATIO * self.GRAVITY) / (4 * self.KF))  # :119
        self.MAX_THRUST = (4 * self.KF * self.MAX_RPM ** 2)                   # :120
        self.GND_EFF_H_CLIP = 0.25 * self.PROP_RADIUS * np.sqrt(
            (15 * self.MAX_RPM ** 2 * self.KF * self.GND_EFF_COEFF) / self.MAX_THRUST)  # :128


def integrate_q(quat, omega, dt):
    """BaseAviary.py:879-892: q <- (I cos(th) + (2/|w|) (Lambda/2) sin(th)) q."""
    omega_norm = np.linalg.norm(omega)
    p, q, r = omega
    if np.isclose(omega_norm, 0):
        return quat
    lam = np.array([[0, r, -q, p],
                    [-r, 0, p, q],
                    [q, -p, 0, r],
                    [-p, -q, -r, 0]]) * .5
    theta = omega_norm * dt / 2
    return np.dot(np.eye(4) * np.cos(theta) + 2 / omega_norm * lam * np.sin(theta), quat)


class OracleAviary:
    """One environment: M drones, explicit dynamics, KIN observation, RPM/ONE_D_RPM action.

    task: 'hover' (HoverAviary), 'multihover' (MultiHoverAviary), 'spiral'
    (SpiralFormationAviary), 'meetup' (MeetupAviary), 'flock' (FlockAviary),
    'leaderfollower' (LeaderFollowerAviary).  `act`: 'rpm' | 'one_d_rpm' | 'pid' | 'vel' | 'one_d_pid'
    (the last three through `oracle/dsl_pid.py`, BaseRLAviary.py:73-78,193-235).
    """

    def __init__(self, task="multihover", drone_model="cf2x", num_drones=1,
                 initial_xyzs=None, initial_rpys=None, pyb_freq=240, ctrl_freq=30,
                 act="rpm", aero=0, integrator="quat",
                 spiral_radius=0.4, spiral_period=10.0, height_rate=0.05,
                 target_center=(0.0, 0.0, 0.0), jitter_source=None):
        if pyb_freq % ctrl_freq != 0:                                         # BaseAviary.py:79-80
            raise ValueError("[ERROR] pyb_freq is not divisible by ctrl_freq.")
        self.task = task
        self.P = AirframeParams(drone_model)
        self.NUM_DRONES = 1 if task == "hover" else num_drones
        self.PYB_FREQ, self.CTRL_FREQ = pyb_freq, ctrl_freq
        self.PYB_STEPS_PER_CTRL = int(pyb_freq / ctrl_freq)                   # :81
        self.CTRL_TIMESTEP = 1. / ctrl_freq
        self.PYB_TIMESTEP = 1. / pyb_freq
        self.ACT = act
        self.AERO = aero
        self.INTEGRATOR = integrator
        self.ACTION_BUFFER_SIZE = int(ctrl_freq // 2)                         # BaseRLAviary.py:66
        self.jitter_source = jitter_source
        M = self.NUM_DRONES
        if task == "spiral":                                                  # SpiralAviary.py:39-56
            self.EPISODE_LEN_SEC = 12
            self.R, self.PERIOD = spiral_radius, spiral_period
            self.OMEGA = 2 * np.pi / self.PERIOD
            self.VZ = height_rate
            self.CENTER = np.array(target_center, dtype=np.float64)
            if initial_xyzs is None:
                initial_xyzs = np.array([[self.R * np.cos(2 * np.pi * i / M),
                                          self.R * np.sin(2 * np.pi * i / M), 0.3] for i in range(M)])
        else:
            self.EPISODE_LEN_SEC = 8                                          # HoverAviary.py:52, MultiHover:58
        if initial_xyzs is None:                                              # BaseAviary.py:194-197
            L = self.P.L
            initial_xyzs = np.vstack([np.array([x * 4 * L for x in range(M)]),
                                      np.array([y * 4 * L for y in range(M)]),
                                      np.ones(M) * (self.P.COLLISION_H / 2 - self.P.COLLISION_Z_OFFSET + .1)]
                                     ).transpose().reshape(M, 3)
        self.INIT_XYZS = np.array(initial_xyzs, dtype=np.float64).reshape(M, 3)
        self.INIT_RPYS = (np.zeros((M, 3)) if initial_rpys is None
                          else np.array(initial_rpys, dtype=np.float64).reshape(M, 3))
        self.A = {"rpm": 4, "vel": 4, "pid": 3, "one_d_rpm": 1, "one_d_pid": 1}[act]   # BaseRLAviary.py:141-146
        if act in ("pid", "vel", "one_d_pid"):                                # BaseRLAviary.py:73-78
            if drone_model not in ("cf2x", "cf2p"):
                raise ValueError("[ERROR] in BaseRLAviary.__init()__, no controller is available for the specified drone_model")
            self.ctrl = [DSLPIDOracle() for _ in range(M)]
        if act == "vel":                                                      # BaseRLAviary.py:94-95
            self.SPEED_LIMIT = 0.03 * self.P.MAX_SPEED_KMH * (1000 / 3600)
        self.action_buffer = deque(maxlen=self.ACTION_BUFFER_SIZE)
        for _ in range(self.ACTION_BUFFER_SIZE):                              # BaseRLAviary.py:153-154
            self.action_buffer.append(np.zeros((M, self.A)))
        if task == "hover":
            self.TARGET_POS = np.array([0, 0, 1])                             # HoverAviary.py:51
        elif task == "multihover":
            self.TARGET_POS = self.INIT_XYZS + np.array([[0, 0, 1 / (i + 1)] for i in range(M)])  # :72
        self.termination_reasons = []
        self._housekeeping()
        self._refresh_kinematics()

    # ------------------------------------------------------------------ reset
    def _housekeeping(self):
        """BaseAviary.py:451-505; the 'PyBullet store' is `self._store_*`."""
        M = self.NUM_DRONES
        self.step_counter = 0
        self.last_clipped_action = np.zeros((M, 4))
        self.pos = np.zeros((M, 3))
        self.quat = np.zeros((M, 4))
        self.rpy = np.zeros((M, 3))
        self.vel = np.zeros((M, 3))
        self.ang_v = np.zeros((M, 3))
        self.rpy_rates = np.zeros((M, 3))
        self._store_pos = [tuple(self.INIT_XYZS[i, :]) for i in range(M)]
        self._store_quat = [bm.quaternion_from_euler(self.INIT_RPYS[i, :]) for i in range(M)]  # :488
        self._store_vel = [(0.0, 0.0, 0.0)] * M
        self._store_angv = [(0.0, 0.0, 0.0)] * M

    def _refresh_kinematics(self):
        """BaseAviary.py:509-519 (quaternion passes through Bullet's matrix round trip)."""
        for i in range(self.NUM_DRONES):
            self.pos[i] = self._store_pos[i]
            self.quat[i] = bm.pose_roundtrip(self._store_quat[i])
            self.rpy[i] = bm.euler_from_quaternion(self.quat[i])
            self.vel[i] = self._store_vel[i]
            self.ang_v[i] = self._store_angv[i]

    def reset(self, jitter=None, fixed=False):
        """MultiHoverAviary.py:75-110 then BaseAviary.py:220-255.

        `jitter`: optional iterable of (M,3) arrays standing in for the successive
        `np.random.uniform(-0.25, 0.25, (M, 3))` draws (MultiHoverAviary.py:83,101);
        default = process-global `np.random`, as in the reference.
        `fixed=True` (test injection, the counterpart of BD_RESET_FIXED): skip the
        jitter / clip / rejection block and reset at INIT_XYZS as they are.
        """
        if self.task == "multihover" and fixed:
            self.TARGET_POS = self.INIT_XYZS + np.array([[0, 0, 1 / (i + 1)] for i in range(self.NUM_DRONES)])
            self.termination_reasons = []
        elif self.task == "multihover":
            if not hasattr(self, "ORIGINAL_INIT_XYZS"):
                self.ORIGINAL_INIT_XYZS = self.INIT_XYZS.copy()
            draws = iter(jitter) if jitter is not None else (
                iter(self.jitter_source()) if self.jitter_source is not None else None)

            def draw():
                if draws is None:
                    return np.random.uniform(-0.25, 0.25, (self.NUM_DRONES, 3))
                return np.asarray(next(draws), dtype=np.float64)
            while True:
                cand = self.ORIGINAL_INIT_XYZS.copy() + draw()
                cand[:, 2] = np.clip(cand[:, 2], 0.1, 1.0)
                dists = np.linalg.norm(cand[:, np.newaxis, :] - cand[np.newaxis, :, :], axis=2)
                np.fill_diagonal(dists, np.inf)
                if not np.any(dists < 0.5) and not np.any(cand[:, 2] < 0.1):
                    break
            self.INIT_XYZS = cand
            self.TARGET_POS = self.INIT_XYZS + np.array([[0, 0, 1 / (i + 1)] for i in range(self.NUM_DRONES)])
            self.termination_reasons = []
        self._housekeeping()
        self._refresh_kinematics()
        return self._compute_obs(), self._compute_info()

    # ------------------------------------------------------------------- step
    def _calculate_next_step(self, current_position, destination, step_size=1):
        """BaseAviary.py:1108-1150."""
        direction = destination - current_position
        distance = np.linalg.norm(direction)
        if distance <= step_size:
            return destination
        return current_position + (direction / distance) * step_size

    def _preprocess_action(self, action):
        """BaseRLAviary.py:160-239 (no clipping in the RPM branches)."""
        self.action_buffer.append(action)
        rpm = np.zeros((self.NUM_DRONES, 4))
        for k in range(action.shape[0]):
            target = action[k, :]
            if self.ACT == "rpm":
                rpm[k, :] = np.array(self.P.HOVER_RPM * (1 + 0.05 * target))           # :192
            elif self.ACT == "one_d_rpm":
                rpm[k, :] = np.repeat(self.P.HOVER_RPM * (1 + 0.05 * target), 4)        # :225
            elif self.ACT == "pid":                                                     # :193-206
                state = self.state_vector(k)
                next_pos = self._calculate_next_step(state[0:3], target, 1)
                rpm[k, :] = self.ctrl[k].compute_control(self.CTRL_TIMESTEP, state[0:3], state[3:7],
                                                         state[10:13], target_pos=next_pos)
            elif self.ACT == "vel":                                                     # :207-222
                state = self.state_vector(k)
                if np.linalg.norm(target[0:3]) != 0:
                    v_unit_vector = target[0:3] / np.linalg.norm(target[0:3])
                else:
                    v_unit_vector = np.zeros(3)
                rpm[k, :] = self.ctrl[k].compute_control(
                    self.CTRL_TIMESTEP, state[0:3], state[3:7], state[10:13], target_pos=state[0:3],
                    target_rpy=np.array([0, 0, state[9]]),
                    target_vel=self.SPEED_LIMIT * np.abs(target[3]) * v_unit_vector)
            else:                                                                       # one_d_pid, :226-235
                state = self.state_vector(k)
                rpm[k, :] = self.ctrl[k].compute_control(
                    self.CTRL_TIMESTEP, state[0:3], state[3:7], state[10:13],
                    target_pos=state[0:3] + 0.1 * np.array([0, 0, target[0]]))
        return rpm

    def step(self, action):
        """BaseAviary.py:259-383."""
        action = np.asarray(action)
        clipped_action = np.reshape(self._preprocess_action(action), (self.NUM_DRONES, 4))  # :341
        for _ in range(self.PYB_STEPS_PER_CTRL):                              # :343
            if self.PYB_STEPS_PER_CTRL > 1:                                   # :346-347
                self._refresh_kinematics()
            staged = [self._dynamics(clipped_action[i, :], i) for i in range(self.NUM_DRONES)]
            # the reference writes each drone into PyBullet right away; because every
            # read in _dynamics/_downwash uses self.pos/quat/vel (refreshed only at
            # :347/:374), staging and committing afterwards is equivalent (Jacobi).
            for i, (p_, q_, v_, w_) in enumerate(staged):
                self._store_pos[i], self._store_quat[i] = tuple(p_), tuple(q_)
                self._store_vel[i], self._store_angv[i] = tuple(v_), tuple(w_)
            self.last_clipped_action = clipped_action                         # :372
        self._refresh_kinematics()                                            # :374
        obs = self._compute_obs()
        reward = self._compute_reward()
        terminated = self._compute_terminated()
        truncated = self._compute_truncated()
        info = self._compute_info()
        self.step_counter = self.step_counter + (1 * self.PYB_STEPS_PER_CTRL)  # :382
        return obs, reward, terminated, truncated, info

    # --------------------------------------------------------- aero formulas
    def gnd_effects(self, rpm, n):
        """Per-propeller ground-effect thrust, BaseAviary.py:739-742 (0 if tilted >= pi/2)."""
        P = self.P
        rot = np.array(bm.matrix_from_quaternion(self.quat[n, :])).reshape(3, 3)
        prop_heights = np.array([self.pos[n, 2] + np.dot(rot, P.PROPS[i])[2] for i in range(4)])
        prop_heights = np.clip(prop_heights, P.GND_EFF_H_CLIP, np.inf)
        g = np.array(rpm ** 2) * P.KF * P.GND_EFF_COEFF * (P.PROP_RADIUS / (4 * prop_heights)) ** 2
        if np.abs(self.rpy[n, 0]) < np.pi / 2 and np.abs(self.rpy[n, 1]) < np.pi / 2:
            return g
        return np.zeros(4)

    def drag_world(self, rpm_last, n):
        """World-frame drag: R (R^T (k*v)) = k*v, BaseAviary.py:773-774."""
        drag_factors = -1 * self.P.DRAG_COEFF * np.sum(np.array(2 * np.pi * rpm_last / 60))
        return drag_factors * np.array(self.vel[n, :])

    def downwash_body_z(self, n):
        """Sum of the body-z downwash forces on drone n, BaseAviary.py:798-804."""
        P = self.P
        total = 0.0
        for i in range(self.NUM_DRONES):
            delta_z = self.pos[i, 2] - self.pos[n, 2]
            delta_xy = np.linalg.norm(np.array(self.pos[i, 0:2]) - np.array(self.pos[n, 0:2]))
            if delta_z > 0 and delta_xy < 10:
                alpha = P.DW_COEFF_1 * (P.PROP_RADIUS / (4 * delta_z)) ** 2
                beta = P.DW_COEFF_2 * delta_z + P.DW_COEFF_3
                with np.errstate(divide="ignore", invalid="ignore"):
                    total = total + (-alpha * np.exp(-.5 * (delta_xy / beta) ** 2))
        return total

    # --------------------------------------------------------------- dynamics
    def _dynamics(self, rpm, n):
        """BaseAviary.py:815-877 (+ this repo's aero composition, see module docstring)."""
        P = self.P
        pos, quat, vel = self.pos[n, :], self.quat[n, :], self.vel[n, :]
        rpy_rates = self.rpy_rates[n, :]
        rotation = np.array(bm.matrix_from_quaternion(quat)).reshape(3, 3)    # :836
        forces = np.array(rpm ** 2) * P.KF                                    # :838
        if self.AERO & AERO_GND:
            forces = forces + self.gnd_effects(rpm, n)
        thrust = np.array([0, 0, np.sum(forces)])
        thrust_world_frame = np.dot(rotation, thrust)
        force_world_frame = thrust_world_frame - np.array([0, 0, P.GRAVITY])  # :841
        if self.AERO & AERO_DRAG:
            force_world_frame = force_world_frame + self.drag_world(self.last_clipped_action[n, :], n)
        if self.AERO & AERO_DW:
            force_world_frame = force_world_frame + np.dot(rotation, np.array([0, 0, self.downwash_body_z(n)]))
        z_torques = np.array(rpm ** 2) * P.KM
        if P.model == "racer":
            z_torques = -z_torques
        z_torque = (-z_torques[0] + z_torques[1] - z_torques[2] + z_torques[3])  # :845
        if P.model == "racer":
            x_torque = (forces[0] + forces[1] - forces[2] - forces[3]) * (P.L / np.sqrt(2))
            y_torque = (- forces[0] + forces[1] + forces[2] - forces[3]) * (P.L / np.sqrt(2))
        elif P.model == "cf2x":
            x_torque = - (forces[0] + forces[1] - forces[2] - forces[3]) * (P.L / np.sqrt(2))
            y_torque = (- forces[0] + forces[1] + forces[2] - forces[3]) * (P.L / np.sqrt(2))
        else:
            x_torque = (forces[1] - forces[3]) * P.L
            y_torque = (-forces[0] + forces[2]) * P.L
        torques = np.array([x_torque, y_torque, z_torque])
        torques = torques - np.cross(rpy_rates, np.dot(P.J, rpy_rates))       # :856
        rpy_rates_deriv = np.dot(P.J_INV, torques)
        accs = force_world_frame / P.M
        vel = vel + self.PYB_TIMESTEP * accs                                  # :860-863
        rpy_rates = rpy_rates + self.PYB_TIMESTEP * rpy_rates_deriv
        pos = pos + self.PYB_TIMESTEP * vel
        if self.INTEGRATOR == "quat":
            quat = integrate_q(quat, rpy_rates, self.PYB_TIMESTEP)
            ang_v = np.dot(rotation, rpy_rates)                               # :873
        else:   # scg/base_aviary.py:499-508: Euler-angle integration, rates stored un-rotated
            rpy = self.rpy[n, :] + self.PYB_TIMESTEP * rpy_rates
            quat = np.array(bm.quaternion_from_euler(rpy))
            ang_v = rpy_rates
        self.rpy_rates[n, :] = rpy_rates                                      # :877
        return pos, quat, vel, ang_v

    # --------------------------------------------------------------- outputs
    def state_vector(self, n):
        """BaseAviary.py:559-561."""
        return np.hstack([self.pos[n, :], self.quat[n, :], self.rpy[n, :], self.vel[n, :],
                          self.ang_v[n, :], self.last_clipped_action[n, :]]).reshape(20,)

    def _spiral_reference(self, i):
        """SpiralAviary.py:82-99."""
        t = self.step_counter / self.PYB_FREQ
        phase = self.OMEGA * t + 2 * np.pi * i / self.NUM_DRONES
        pos_ref = np.array([self.CENTER[0] + self.R * np.cos(phase),
                            self.CENTER[1] + self.R * np.sin(phase), 0.3 + self.VZ * t])
        vel_ref = np.array([-self.R * self.OMEGA * np.sin(phase), self.R * self.OMEGA * np.cos(phase), self.VZ])
        return pos_ref, vel_ref, phase

    def _compute_obs(self):
        """BaseRLAviary.py:307-319 (+ SpiralAviary.py:120-146)."""
        M = self.NUM_DRONES
        obs_12 = np.zeros((M, 12))
        for i in range(M):
            s = self.state_vector(i)
            obs_12[i, :] = np.hstack([s[0:3], s[7:10], s[10:13], s[13:16]]).reshape(12,)
        ret = np.array([obs_12[i, :] for i in range(M)]).astype('float32')
        for i in range(self.ACTION_BUFFER_SIZE):
            ret = np.hstack([ret, np.array([self.action_buffer[i][j, :] for j in range(M)])])
        if self.task != "spiral":
            return ret
        augmented = []
        for i in range(M):
            s = self.state_vector(i)
            pos, vel = s[0:3], s[3:6]            # SpiralAviary.py:129-130 ("vel" is quat xyz)
            pos_ref, vel_ref, phase = self._spiral_reference(i)
            extra = np.concatenate([pos_ref - pos, vel_ref - vel,
                                    np.array([np.sin(phase), np.cos(phase)]), vel_ref])
            augmented.append(np.concatenate([ret[i], extra]))
        return np.array(augmented, dtype=np.float32)

    def _compute_reward(self):
        M = self.NUM_DRONES
        if self.task == "hover":                                              # HoverAviary.py:77-79
            s = self.state_vector(0)
            return max(0, 2 - np.linalg.norm(self.TARGET_POS - s[0:3]) ** 4)
        if self.task == "meetup":                                             # MeetupAviary.py:74-95
            total_reward = 0
            states = np.array([self.state_vector(i) for i in range(M)])
            for i in range(int(M / 2)):
                pair_reward = -1 * np.linalg.norm(states[i, 0:3] - states[M - 1 - i, 0:3]) ** 2
                total_reward += pair_reward * 2
            return total_reward
        if self.task == "leaderfollower":                                     # LeaderFollowerAviary.py:72-99
            total_reward = 0
            states = np.array([self.state_vector(i) for i in range(M)])
            total_reward += -1 * np.linalg.norm(np.array([0, 0, 0.5]) - states[0, 0:3]) ** 2
            for i in range(1, M):
                target_pos = np.array([states[i, 0], states[i, 1], states[0, 2]])
                total_reward += -(1 / M) * np.linalg.norm(target_pos - states[i, 0:3]) ** 2
            return total_reward
        if self.task == "flock":                                              # FlockAviary.py:75-150
            states = np.array([self.state_vector(i) for i in range(M)])
            pos, vel = states[None, :, 0:3].copy(), states[None, :, 10:13].copy()
            ali = 0
            EPSILON = 1e-3
            linear_vel_norm = np.linalg.norm(vel, axis=2)
            for i in range(M):
                for j in range(M):
                    if j != i:
                        d = np.einsum('ij,ij->i', vel[:, i, :], vel[:, j, :])
                        ali += (d / (linear_vel_norm[:, i] + EPSILON) / (linear_vel_norm[:, j] + EPSILON))
            if M > 1:
                ali /= (M * (M - 1))
            else:
                ali = np.array([0.0])
            cof_v = np.mean(vel, axis=1)
            avg_flock_linear_speed = np.linalg.norm(cof_v, axis=-1)
            avg_flock_spac_rew = 0.0
            var_flock_spacing = np.array([0.0])
            if M > 1:
                whole_flock_spacing = []
                for i in range(M):
                    flck_neighbor_pos = np.delete(pos, [i], 1)
                    diff = flck_neighbor_pos - np.reshape(pos[:, i, :], (pos[:, i, :].shape[0], 1, -1))
                    whole_flock_spacing.append(np.amin(np.linalg.norm(diff, axis=-1), axis=-1))
                whole_flock_spacing = np.stack(whole_flock_spacing, axis=-1)
                avg_flock_spacing = np.mean(whole_flock_spacing, axis=-1)
                var_flock_spacing = np.var(whole_flock_spacing, axis=-1)
                if 1.0 < avg_flock_spacing[0] < 3.0:                          # FLOCK_SPACING_MIN / MAX
                    avg_flock_spac_rew = 0.0
                else:
                    avg_flock_spac_rew = min(math.fabs(avg_flock_spacing[0] - 1.0), math.fabs(avg_flock_spacing[0] - 3.0))
            return ali[0] + avg_flock_linear_speed[0] - avg_flock_spac_rew - var_flock_spacing[0]
        if self.task == "multihover":                                         # MultiHoverAviary.py:128-186
            reward = 0.0
            for i in range(M):
                s = self.state_vector(i)
                pos, vel, target = s[0:3], s[10:13], self.TARGET_POS[i]
                err_xy = np.linalg.norm(pos[0:2] - target[0:2])
                err_z = pos[2] - target[2]
                vel_z = vel[2]
                r_xy = 1.0 / (1 + err_xy)
                r_z = np.exp(-7.5 * abs(err_z))
                r_vel = -1.5 * vel_z ** 2 if abs(err_z) < 0.2 else 0.0
                hover_bonus = 0.5 if (err_xy < 0.03 and abs(err_z) < 0.03 and abs(vel_z) < 0.03) else 0.0
                reward += (r_xy + r_z + r_vel + hover_bonus)
            reward /= M
            return float(reward)
        reward = 0.0                                                          # SpiralAviary.py:150-181
        for i in range(M):
            s = self.state_vector(i)
            pos, vel = s[0:3], s[3:6]
            pos_ref, vel_ref, _ = self._spiral_reference(i)
            r_pos = np.exp(-4.0 * np.linalg.norm(pos - pos_ref) ** 2)
            r_vel = np.exp(-2.0 * np.linalg.norm(vel - vel_ref) ** 2)
            r_xy = pos[0:2] - self.CENTER[0:2]
            if np.linalg.norm(r_xy) > 1e-3:
                radial = r_xy / np.linalg.norm(r_xy)
                tangent = np.array([-radial[1], radial[0]])
                v_xy = vel[0:2]
                if np.linalg.norm(v_xy) > 1e-3:
                    r_tan = max(0.0, np.dot(v_xy / np.linalg.norm(v_xy), tangent))
                else:
                    r_tan = 0.0
            else:
                r_tan = 0.0
            reward += 1.0 * r_pos + 2.0 * r_vel + 1.0 * r_tan
        return reward / M

    def _compute_terminated(self):
        if self.task == "hover":                                              # HoverAviary.py:92-96
            s = self.state_vector(0)
            return bool(np.linalg.norm(self.TARGET_POS - s[0:3]) < .0001)
        if self.task == "meetup":                                             # MeetupAviary.py:99-121
            states = np.array([self.state_vector(i) for i in range(self.NUM_DRONES)])
            for i in range(int(self.NUM_DRONES / 2)):
                if np.linalg.norm(states[i, 0:3] - states[self.NUM_DRONES - 1 - i, 0:3]) > 0.1:
                    return False
            return True
        if self.task in ("flock", "leaderfollower"):                          # FlockAviary.py:154-165, LeaderFollower:103-115
            return False
        if self.task == "multihover":                                         # MultiHoverAviary.py:216-241
            terminated, reasons = False, []
            for i in range(self.NUM_DRONES):
                s = self.state_vector(i)
                x, y, z, roll, pitch = s[0], s[1], s[2], s[7], s[8]
                if z < 0.03:
                    terminated = True
                    reasons.append(f"Drone {i} crashed (z={z:.2f})")
                if abs(roll) > 1.2 or abs(pitch) > 1.2:
                    terminated = True
                    reasons.append(f"Drone {i} flipped (roll={roll:.2f}, pitch={pitch:.2f})")
                if abs(x) > 3.0 or abs(y) > 3.0:
                    terminated = True
                    reasons.append(f"Drone {i} out of bounds (pos=[{x:.2f}, {y:.2f}, {z:.2f}])")
            self.termination_reasons = reasons
            return terminated
        for i in range(self.NUM_DRONES):                                      # SpiralAviary.py:185-191
            z = self.state_vector(i)[2]
            if z < 0.05 or z > 3.0:
                return True
        return False

    def _compute_truncated(self):
        if self.task == "hover":                                              # HoverAviary.py:108-117
            s = self.state_vector(0)
            if (abs(s[0]) > 1.5 or abs(s[1]) > 1.5 or s[2] > 2.0 or abs(s[7]) > .4 or abs(s[8]) > .4):
                return True
        if self.task in ("meetup", "flock", "leaderfollower"):
            # MeetupAviary.py:125-154 (also z < 0.1), FlockAviary.py:169-189, LeaderFollowerAviary.py:119-145
            xy, zmax = {"meetup": (5.0, 3.0), "flock": (10.0, 10.0), "leaderfollower": (2.0, 2.0)}[self.task]
            for i in range(self.NUM_DRONES):
                s = self.state_vector(i)
                if (abs(s[0]) > xy or abs(s[1]) > xy or s[2] > zmax
                        or (self.task == "meetup" and s[2] < 0.1)
                        or abs(s[7]) > .4 or abs(s[8]) > .4):
                    return True
        return bool(self.step_counter / self.PYB_FREQ > self.EPISODE_LEN_SEC)  # MultiHover:268, Spiral:196

    def _compute_info(self):
        if self.task in ("hover", "meetup", "flock", "leaderfollower"):
            return {"answer": 42}
        if self.task == "multihover":
            return {"answer": 42, "termination_reasons": self.termination_reasons}
        return {"time": self.step_counter / self.PYB_FREQ, "omega": self.OMEGA, "radius": self.R}

    def close(self):
        pass


def step_env_autoreset(env, action, **reset_kwargs):
    """subproc_vec_env.py:188-207: 5-tuple -> (ob, reward, done, info) with auto-reset."""
    ob, reward, terminated, truncated, info = env.step(action)
    done = terminated or truncated
    if done:
        end_obs, end_info = np.array(ob, copy=True), dict(info)
        ob, info = env.reset(**reset_kwargs)
        info = dict(info)
        info["terminal_observation"] = end_obs
        info["terminal_info"] = end_info
    return ob, reward, done, info


class OracleVecEnv:
    """Sequential stand-in for SubprocVecEnv.step/reset (subproc_vec_env.py:51-73)."""

    def __init__(self, envs, **reset_kwargs):
        self.envs = list(envs)
        self.num_envs = len(self.envs)
        self.reset_kwargs = reset_kwargs

    def reset(self):
        res = [e.reset(**self.reset_kwargs) for e in self.envs]
        return np.stack([r[0] for r in res]), {"n": [r[1] for r in res]}

    def step(self, actions):
        res = [step_env_autoreset(e, a, **self.reset_kwargs) for e, a in zip(self.envs, actions)]
        obs, rews, dones, infos = zip(*res)
        return np.stack(obs), np.stack(rews), np.stack(dones), {"n": infos}
