"""oracle/dsl_pid.py: closed forms vs scipy (the library the reference's controller calls,
DSLPIDControl.py:159-160,201-203) and known answers of the controller itself."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle.dsl_pid import CF2X_KF, CF2X_MASS, DSLPIDOracle, euler_XYZ_from_matrix, matrix_from_euler_XYZ


def test_intrinsic_xyz_closed_forms_match_scipy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        e = rng.uniform(-1.4, 1.4, 3)
        R = Rotation.from_euler('XYZ', e).as_matrix()
        assert np.allclose(matrix_from_euler_XYZ(e), R, atol=1e-15)
        assert np.allclose(euler_XYZ_from_matrix(R), Rotation.from_matrix(R).as_euler('XYZ'), atol=1e-13)
        # the reference's quaternion relabelling (w,x,y,z = as_quat(); from_quat([w,x,y,z])) is a no-op
        w, x, y, z = Rotation.from_euler('XYZ', e).as_quat()
        assert np.allclose(Rotation.from_quat([w, x, y, z]).as_matrix(), R, atol=1e-15)


def test_hover_equilibrium_commands_hover_rpm():
    """Level, at rest, on target: thrust = m g, no torque -> every motor at sqrt(m g / 4 kf)."""
    c = DSLPIDOracle()
    p = np.array([0.3, -0.2, 1.0])
    rpm = c.compute_control(1 / 48, p, np.array([0., 0, 0, 1]), np.zeros(3), target_pos=p)
    assert np.allclose(rpm, np.sqrt(9.8 * CF2X_MASS / (4 * CF2X_KF)), rtol=1e-12)
    assert np.all(c.state() == 0)


def test_integrator_clips_and_pwm_limits():
    c = DSLPIDOracle()
    for _ in range(400):      # 400 * 10 m * (1/30) s >> 2: the position integrator saturates
        rpm = c.compute_control(1 / 30, np.zeros(3), np.array([0., 0, 0, 1]), np.zeros(3),
                                target_pos=np.array([10., -10., 10.]))
    assert np.array_equal(c.integral_pos_e, [2.0, -2.0, 0.15])                 # DSLPIDControl.py:181-182
    assert rpm.min() >= 0.2685 * 20000 + 4070.3 - 1e-9 and rpm.max() <= 0.2685 * 65535 + 4070.3 + 1e-9


def test_racer_has_no_controller():
    from oracle.aviary_oracle import OracleAviary
    with pytest.raises(ValueError):
        OracleAviary(task="multihover", drone_model="racer", act="vel")
