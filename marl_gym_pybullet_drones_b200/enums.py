"""API enums of the drone aviaries.

Same member names and string values as the reference
(`gym_pybullet_drones/utils/enums.py:3-48`) so that user code written against
the reference (`DroneModel.CF2X`, `Physics("dyn")`, `ActionType('one_d_rpm')`,
...) keeps working unchanged against `BatchAviary`.
"""
from enum import Enum


class DroneModel(Enum):
    """Airframes (reference enums.py:3-8)."""

    CF2X = "cf2x"
    CF2P = "cf2p"
    RACE = "racer"


class Physics(Enum):
    """Physics updates (reference enums.py:13-21).

    Only `DYN` is stepped by the CUDA kernels.  The aerodynamic add-ons the
    reference applies through PyBullet (`PYB_GND`, `PYB_DRAG`, `PYB_DW`,
    `PYB_GND_DRAG_DW`) are available on top of `DYN` through the `DYN_*`
    members below, which do not exist in the reference (it never composes
    them with DYN; see DESIGN.md "aero composition").
    """

    PYB = "pyb"
    DYN = "dyn"
    PYB_GND = "pyb_gnd"
    PYB_DRAG = "pyb_drag"
    PYB_DW = "pyb_dw"
    PYB_GND_DRAG_DW = "pyb_gnd_drag_dw"
    # extensions (explicit dynamics + the reference's aero formulas)
    DYN_GND = "dyn_gnd"
    DYN_DRAG = "dyn_drag"
    DYN_DW = "dyn_dw"
    DYN_GND_DRAG_DW = "dyn_gnd_drag_dw"


class ImageType(Enum):
    """Camera capture kinds (reference enums.py:25-31); unused by the DYN path."""

    RGB = 0
    DEP = 1
    SEG = 2
    BW = 3


class ActionType(Enum):
    """Action kinds (reference enums.py:35-41)."""

    RPM = "rpm"
    PID = "pid"
    VEL = "vel"
    ONE_D_RPM = "one_d_rpm"
    ONE_D_PID = "one_d_pid"


class ObservationType(Enum):
    """Observation kinds (reference enums.py:45-48)."""

    KIN = "kin"
    RGB = "rgb"


# aero bit flags shared with the C-ABI (include/batch_drones.h)
AERO_GND = 1
AERO_DRAG = 2
AERO_DW = 4

_PHYSICS_AERO = {
    Physics.DYN: 0,
    Physics.DYN_GND: AERO_GND,
    Physics.DYN_DRAG: AERO_DRAG,
    Physics.DYN_DW: AERO_DW,
    Physics.DYN_GND_DRAG_DW: AERO_GND | AERO_DRAG | AERO_DW,
}


def physics_aero_flags(physics: Physics) -> int:
    """Aero bit mask of an explicit-dynamics mode; PYB* modes are not provided."""
    if physics not in _PHYSICS_AERO:
        raise NotImplementedError(
            f"{physics} needs PyBullet's rigid-body solver; only the explicit "
            "dynamics modes (Physics.DYN, DYN_GND, DYN_DRAG, DYN_DW, "
            "DYN_GND_DRAG_DW) are stepped on the GPU")
    return _PHYSICS_AERO[physics]
