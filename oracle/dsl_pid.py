"""TEST INFRASTRUCTURE — numpy restatement of the reference's DSL PID controller.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may
import this module.  It restates `gym_pybullet_drones/control/DSLPIDControl.py`
(:20-262) and `control/BaseControl.py` (:18-52) as they are used from
`BaseRLAviary._preprocessAction` (`envs/BaseRLAviary.py:193-235`):

* `BaseRLAviary.__init__` (:73-78) always builds `DSLPIDControl(DroneModel.CF2X)`,
  whatever the env's airframe, so the gains, mixer, mass and kf are CF2X's;
* the controllers are created once and NEVER reset by `env.reset()`: integral
  errors and `last_rpy` carry over episode boundaries (`ctrl.reset()` is only
  called from the constructor, DSLPIDControl.py:60);
* the scipy round trip in `_dslPIDPositionControl/_dslPIDAttitudeControl`
  (:159-160, :201-203: matrix -> intrinsic 'XYZ' Euler angles -> quaternion ->
  `w,x,y,z = (x,y,z,w)` relabelling -> `from_quat([w,x,y,z])`, which undoes the
  relabelling -> matrix) is the identity on a rotation matrix up to rounding; it is
  restated with closed forms (`euler_XYZ_from_matrix`, `matrix_from_euler_XYZ`)
  and cross-checked against scipy in `tests/test_oracle_dsl_pid.py`.

Pinned against trajectories of the unmodified reference classes stepped with
`ActionType.PID / VEL / ONE_D_PID` (`tests/golden/make_golden.py`, cases `*_pid`,
`*_vel`, `*_one_d_pid`).
"""
import math

import numpy as np

from . import bullet_math as bm

CF2X_MASS, CF2X_KF, CTRL_G = 0.027, 3.16e-10, 9.8        # cf2x.urdf:5,11; BaseControl.py:20,35-37


def euler_XYZ_from_matrix(R):
    """Intrinsic X-Y-Z angles (a,b,c) with R = Rx(a) Ry(b) Rz(c) (scipy `as_euler('XYZ')`)."""
    b = math.asin(max(-1.0, min(1.0, R[0, 2])))
    a = math.atan2(-R[1, 2], R[2, 2])
    c = math.atan2(-R[0, 1], R[0, 0])
    return np.array([a, b, c])


def matrix_from_euler_XYZ(e):
    """R = Rx(a) Ry(b) Rz(c) (scipy `from_euler('XYZ', e).as_matrix()`)."""
    a, b, c = e
    sa, ca, sb, cb, sc, cc = math.sin(a), math.cos(a), math.sin(b), math.cos(b), math.sin(c), math.cos(c)
    return np.array([[cb * cc, -cb * sc, sb],
                     [ca * sc + sa * sb * cc, ca * cc - sa * sb * sc, -sa * cb],
                     [sa * sc - ca * sb * cc, sa * cc + ca * sb * sc, ca * cb]])


class DSLPIDOracle:
    """DSLPIDControl(DroneModel.CF2X), DSLPIDControl.py:20-262."""

    def __init__(self, g=CTRL_G):
        self.GRAVITY = g * CF2X_MASS                                          # BaseControl.py:35
        self.KF = CF2X_KF                                                     # BaseControl.py:37
        self.P_COEFF_FOR = np.array([.4, .4, 1.25])                           # :37-42
        self.I_COEFF_FOR = np.array([.05, .05, .05])
        self.D_COEFF_FOR = np.array([.2, .2, .5])
        self.P_COEFF_TOR = np.array([70000., 70000., 60000.])
        self.I_COEFF_TOR = np.array([.0, .0, 500.])
        self.D_COEFF_TOR = np.array([20000., 20000., 12000.])
        self.PWM2RPM_SCALE, self.PWM2RPM_CONST = 0.2685, 4070.3               # :43-44
        self.MIN_PWM, self.MAX_PWM = 20000, 65535                             # :45-46
        self.MIXER_MATRIX = np.array([[-.5, -.5, -1], [-.5, .5, 1], [.5, .5, -1], [.5, -.5, 1]])  # :48-53
        self.reset()

    def reset(self):
        """DSLPIDControl.py:64-79 (the fields that are read later)."""
        self.control_counter = 0
        self.last_rpy = np.zeros(3)
        self.integral_pos_e = np.zeros(3)
        self.integral_rpy_e = np.zeros(3)

    def state(self):
        return np.concatenate([self.integral_pos_e, self.integral_rpy_e, self.last_rpy])

    def compute_control(self, control_timestep, cur_pos, cur_quat, cur_vel, target_pos,
                        target_rpy=np.zeros(3), target_vel=np.zeros(3), target_rpy_rates=np.zeros(3)):
        """DSLPIDControl.py:82-139; returns the 4 motor RPMs."""
        self.control_counter += 1
        thrust, target_euler = self._position(control_timestep, cur_pos, cur_quat, cur_vel,
                                              target_pos, target_rpy, target_vel)
        return self._attitude(control_timestep, thrust, cur_quat, target_euler, target_rpy_rates)

    def _position(self, dt, cur_pos, cur_quat, cur_vel, target_pos, target_rpy, target_vel):
        """DSLPIDControl.py:143-197."""
        cur_rotation = np.array(bm.matrix_from_quaternion(cur_quat)).reshape(3, 3)
        pos_e = target_pos - cur_pos
        vel_e = target_vel - cur_vel
        self.integral_pos_e = self.integral_pos_e + pos_e * dt
        self.integral_pos_e = np.clip(self.integral_pos_e, -2., 2.)
        self.integral_pos_e[2] = np.clip(self.integral_pos_e[2], -0.15, .15)
        target_thrust = np.multiply(self.P_COEFF_FOR, pos_e) \
            + np.multiply(self.I_COEFF_FOR, self.integral_pos_e) \
            + np.multiply(self.D_COEFF_FOR, vel_e) + np.array([0, 0, self.GRAVITY])
        scalar_thrust = max(0., np.dot(target_thrust, cur_rotation[:, 2]))
        thrust = (math.sqrt(scalar_thrust / (4 * self.KF)) - self.PWM2RPM_CONST) / self.PWM2RPM_SCALE
        target_z_ax = target_thrust / np.linalg.norm(target_thrust)
        target_x_c = np.array([math.cos(target_rpy[2]), math.sin(target_rpy[2]), 0])
        target_y_ax = np.cross(target_z_ax, target_x_c) / np.linalg.norm(np.cross(target_z_ax, target_x_c))
        target_x_ax = np.cross(target_y_ax, target_z_ax)
        target_rotation = (np.vstack([target_x_ax, target_y_ax, target_z_ax])).transpose()
        return thrust, euler_XYZ_from_matrix(target_rotation)

    def _attitude(self, dt, thrust, cur_quat, target_euler, target_rpy_rates):
        """DSLPIDControl.py:201-246."""
        cur_rotation = np.array(bm.matrix_from_quaternion(cur_quat)).reshape(3, 3)
        cur_rpy = np.array(bm.euler_from_quaternion(cur_quat))
        target_rotation = matrix_from_euler_XYZ(target_euler)
        rot_matrix_e = np.dot(target_rotation.transpose(), cur_rotation) - np.dot(cur_rotation.transpose(), target_rotation)
        rot_e = np.array([rot_matrix_e[2, 1], rot_matrix_e[0, 2], rot_matrix_e[1, 0]])
        rpy_rates_e = target_rpy_rates - (cur_rpy - self.last_rpy) / dt
        self.last_rpy = cur_rpy
        self.integral_rpy_e = self.integral_rpy_e - rot_e * dt
        self.integral_rpy_e = np.clip(self.integral_rpy_e, -1500., 1500.)
        self.integral_rpy_e[0:2] = np.clip(self.integral_rpy_e[0:2], -1., 1.)
        target_torques = - np.multiply(self.P_COEFF_TOR, rot_e) \
            + np.multiply(self.D_COEFF_TOR, rpy_rates_e) \
            + np.multiply(self.I_COEFF_TOR, self.integral_rpy_e)
        target_torques = np.clip(target_torques, -3200, 3200)
        pwm = thrust + np.dot(self.MIXER_MATRIX, target_torques)
        pwm = np.clip(pwm, self.MIN_PWM, self.MAX_PWM)
        return self.PWM2RPM_SCALE * pwm + self.PWM2RPM_CONST
