"""Enums keep the reference's names/values; URDF-derived constants match SURVEY.md §8(a) row 1."""
import numpy as np
import pytest

from marl_gym_pybullet_drones_b200 import (ActionType, DroneModel, ImageType, ObservationType, Physics,
                                           drone_constants)
from marl_gym_pybullet_drones_b200.enums import physics_aero_flags
from oracle.aviary_oracle import AirframeParams


def test_enum_values_match_reference():   # utils/enums.py:3-48
    assert [m.value for m in DroneModel] == ["cf2x", "cf2p", "racer"]
    assert Physics("dyn") is Physics.DYN and Physics("pyb_gnd_drag_dw") is Physics.PYB_GND_DRAG_DW
    assert [a.value for a in ActionType] == ["rpm", "pid", "vel", "one_d_rpm", "one_d_pid"]
    assert [o.value for o in ObservationType] == ["kin", "rgb"]
    assert [i.value for i in ImageType] == [0, 1, 2, 3]


def test_cf2x_derived_constants():
    k = drone_constants(DroneModel.CF2X)
    assert (k.M, k.L, k.KF, k.KM, k.THRUST2WEIGHT_RATIO) == (0.027, 0.0397, 3.16e-10, 7.94e-12, 2.25)
    assert np.allclose(np.diag(k.J), [1.4e-5, 1.4e-5, 2.17e-5], rtol=0, atol=0)
    assert k.GRAVITY == pytest.approx(0.2646, rel=1e-15)
    assert k.HOVER_RPM == pytest.approx(14468.429183500699, rel=1e-15)
    assert k.MAX_RPM == pytest.approx(21702.64377525105, rel=1e-15)
    assert k.MAX_THRUST == pytest.approx(0.59535, rel=1e-12)
    assert k.GND_EFF_H_CLIP == pytest.approx(0.03776371349209501, rel=1e-14)
    assert k.ARM_XY == pytest.approx(0.028072139213105935, rel=1e-15)
    assert k.DEFAULT_SPAWN_Z == pytest.approx(0.1125)
    assert k.PROP_OFFSETS[0] == (0.028, -0.028, 0.0) and k.PROP_OFFSETS[3] == (0.028, 0.028, 0.0)


@pytest.mark.parametrize("model", list(DroneModel))
def test_package_constants_equal_oracle_table(model):
    k, o = drone_constants(model), AirframeParams(model.value)
    for name in ("M", "L", "KF", "KM", "THRUST2WEIGHT_RATIO", "GRAVITY", "HOVER_RPM", "MAX_RPM", "MAX_THRUST",
                 "GND_EFF_H_CLIP", "GND_EFF_COEFF", "PROP_RADIUS", "DW_COEFF_1", "DW_COEFF_2", "DW_COEFF_3"):
        assert getattr(k, name) == getattr(o, name), name
    assert np.array_equal(k.J, o.J) and np.array_equal(k.DRAG_COEFF, o.DRAG_COEFF)
    assert np.array_equal(np.array(k.PROP_OFFSETS), o.PROPS)


def test_cf2p_inertia():   # cf2p.urdf:12
    assert np.array_equal(np.diag(drone_constants(DroneModel.CF2P).J), [2.3951e-5, 2.3951e-5, 3.2347e-5])


def test_physics_modes():
    assert physics_aero_flags(Physics.DYN) == 0
    assert physics_aero_flags(Physics.DYN_GND_DRAG_DW) == 7
    for p in (Physics.PYB, Physics.PYB_GND, Physics.PYB_DRAG, Physics.PYB_DW, Physics.PYB_GND_DRAG_DW):
        with pytest.raises(NotImplementedError):
            physics_aero_flags(p)
