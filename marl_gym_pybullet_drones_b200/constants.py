"""Airframe constants: the URDF `<properties>` reader and the derived quantities.

Produces the same numbers as the reference's `BaseAviary._parseURDFParameters`
(`gym_pybullet_drones/envs/BaseAviary.py:985-1017`) and the constant block of
`BaseAviary.__init__` (`BaseAviary.py:74-128`), with the arithmetic written in
the same order so the fp64 values are bit-identical (`HOVER_RPM`,
`MAX_RPM`, `GND_EFF_H_CLIP`, ...).  The reader looks elements up by tag name
rather than by child index.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from .enums import DroneModel

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


@dataclass(frozen=True)
class DroneConstants:
    """Physical constants of one airframe (all fp64, SI units)."""

    model: DroneModel
    M: float
    L: float
    THRUST2WEIGHT_RATIO: float
    IXX: float
    IYY: float
    IZZ: float
    KF: float
    KM: float
    COLLISION_H: float
    COLLISION_R: float
    COLLISION_Z_OFFSET: float
    MAX_SPEED_KMH: float
    GND_EFF_COEFF: float
    PROP_RADIUS: float
    DRAG_COEFF_XY: float
    DRAG_COEFF_Z: float
    DW_COEFF_1: float
    DW_COEFF_2: float
    DW_COEFF_3: float
    PROP_OFFSETS: tuple = field(default=())  # 4 x (x, y, z) in the body frame
    G: float = 9.8

    # ---- derived exactly like BaseAviary.py:117-128 -------------------------
    @property
    def J(self) -> np.ndarray:
        return np.diag([self.IXX, self.IYY, self.IZZ])

    @property
    def J_INV(self) -> np.ndarray:
        return np.linalg.inv(self.J)

    @property
    def DRAG_COEFF(self) -> np.ndarray:
        return np.array([self.DRAG_COEFF_XY, self.DRAG_COEFF_XY, self.DRAG_COEFF_Z])

    @property
    def GRAVITY(self) -> float:
        return self.G * self.M

    @property
    def HOVER_RPM(self) -> float:
        return float(np.sqrt(self.GRAVITY / (4 * self.KF)))

    @property
    def MAX_RPM(self) -> float:
        return float(np.sqrt((self.THRUST2WEIGHT_RATIO * self.GRAVITY) / (4 * self.KF)))

    @property
    def MAX_THRUST(self) -> float:
        return 4 * self.KF * self.MAX_RPM ** 2

    @property
    def MAX_XY_TORQUE(self) -> float:
        if self.model == DroneModel.CF2P:
            return self.L * self.KF * self.MAX_RPM ** 2
        return float((2 * self.L * self.KF * self.MAX_RPM ** 2) / np.sqrt(2))

    @property
    def MAX_Z_TORQUE(self) -> float:
        return 2 * self.KM * self.MAX_RPM ** 2

    @property
    def GND_EFF_H_CLIP(self) -> float:
        return float(0.25 * self.PROP_RADIUS * np.sqrt(
            (15 * self.MAX_RPM ** 2 * self.KF * self.GND_EFF_COEFF) / self.MAX_THRUST))

    @property
    def ARM_XY(self) -> float:
        """Lever arm of the roll/pitch torques: L/sqrt(2) for X frames, L for +."""
        if self.model == DroneModel.CF2P:
            return self.L
        return float(self.L / np.sqrt(2))

    @property
    def DEFAULT_SPAWN_Z(self) -> float:
        """BaseAviary.py:197."""
        return self.COLLISION_H / 2 - self.COLLISION_Z_OFFSET + .1


def _floats(text: str):
    return [float(s) for s in text.split()]


def parse_urdf(model: DroneModel, path: str | None = None) -> DroneConstants:
    """Read `<model>.urdf` (ours, or any file with the reference's schema)."""
    path = path or os.path.join(ASSET_DIR, model.value + ".urdf")
    root = ET.parse(path).getroot()
    props = root.find("properties").attrib
    base = next(l for l in root.findall("link") if l.get("name") == "base_link")
    inertial = base.find("inertial")
    inertia = inertial.find("inertia").attrib
    coll = base.find("collision")
    cyl = coll.find("geometry").find("cylinder").attrib
    offsets = []
    for i in range(4):
        link = next((l for l in root.findall("link") if l.get("name") == f"prop{i}_link"), None)
        if link is not None:
            offsets.append(tuple(_floats(link.find("inertial").find("origin").get("xyz"))))
    return DroneConstants(
        model=model,
        M=float(inertial.find("mass").get("value")),
        L=float(props["arm"]),
        THRUST2WEIGHT_RATIO=float(props["thrust2weight"]),
        IXX=float(inertia["ixx"]), IYY=float(inertia["iyy"]), IZZ=float(inertia["izz"]),
        KF=float(props["kf"]), KM=float(props["km"]),
        COLLISION_H=float(cyl["length"]), COLLISION_R=float(cyl["radius"]),
        COLLISION_Z_OFFSET=_floats(coll.find("origin").get("xyz"))[2],
        MAX_SPEED_KMH=float(props["max_speed_kmh"]),
        GND_EFF_COEFF=float(props["gnd_eff_coeff"]),
        PROP_RADIUS=float(props["prop_radius"]),
        DRAG_COEFF_XY=float(props["drag_coeff_xy"]),
        DRAG_COEFF_Z=float(props["drag_coeff_z"]),
        DW_COEFF_1=float(props["dw_coeff_1"]),
        DW_COEFF_2=float(props["dw_coeff_2"]),
        DW_COEFF_3=float(props["dw_coeff_3"]),
        PROP_OFFSETS=tuple(offsets),
    )


_CACHE: dict = {}


def drone_constants(model: DroneModel) -> DroneConstants:
    if model not in _CACHE:
        _CACHE[model] = parse_urdf(model)
    return _CACHE[model]
