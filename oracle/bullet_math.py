"""TEST INFRASTRUCTURE — CPU restatement of the PyBullet calls on the DYN path.

PARITY UNPINNED for this file: PyBullet (pybullet==3.2.7, `environment.yml:83`
of the reference; an un-vendored C++ dependency, not installable here) cannot
be run in this container, so these functions restate the *published* Bullet
algorithms from Bullet3's sources:

* `matrix_from_quaternion`  — `pybullet_getMatrixFromQuaternion` (pybullet.c),
  identical to `btMatrix3x3::setRotation` (btMatrix3x3.h): s = 2/|q|^2.
* `quaternion_from_matrix`  — `btMatrix3x3::getRotation` (btMatrix3x3.h).
* `euler_from_quaternion`   — `pybullet_getEulerFromQuaternion` (pybullet.c),
  with its hard gimbal branches at |sarg| >= 0.99999.
* `quaternion_from_euler`   — `pybullet_getQuaternionFromEuler` (pybullet.c),
  including its final normalisation.
* `pose_roundtrip`          — what `p.resetBasePositionAndOrientation` followed
  by `p.getBasePositionAndOrientation` does to a quaternion for a multibody
  whose base inertial frame is the identity (`cf2x.urdf:10`):
  quaternion -> btTransform basis -> quaternion.

Reference call sites: `BaseAviary.py:488,517,518,836,865`.
They are cross-checked against `scipy.spatial.transform.Rotation` in
`tests/test_oracle_bullet_math.py`.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline leg may import this module.
"""
import math

import numpy as np

PYBULLET_PI = 3.14159265358979323846


def matrix_from_quaternion(q):
    """(x,y,z,w) -> 9 row-major floats (BaseAviary.py:836 call site)."""
    x, y, z, w = float(q[0]), float(q[1]), float(q[2]), float(q[3])
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return (1.0 - (yy + zz), xy - wz, xz + wy,
            xy + wz, 1.0 - (xx + zz), yz - wx,
            xz - wy, yz + wx, 1.0 - (xx + yy))


def quaternion_from_matrix(m):
    """9 row-major floats -> (x,y,z,w); btMatrix3x3::getRotation."""
    el = ((m[0], m[1], m[2]), (m[3], m[4], m[5]), (m[6], m[7], m[8]))
    trace = el[0][0] + el[1][1] + el[2][2]
    temp = [0.0, 0.0, 0.0, 0.0]
    if trace > 0.0:
        s = math.sqrt(trace + 1.0)
        temp[3] = s * 0.5
        s = 0.5 / s
        temp[0] = (el[2][1] - el[1][2]) * s
        temp[1] = (el[0][2] - el[2][0]) * s
        temp[2] = (el[1][0] - el[0][1]) * s
    else:
        if el[0][0] < el[1][1]:
            i = 2 if el[1][1] < el[2][2] else 1
        else:
            i = 2 if el[0][0] < el[2][2] else 0
        j = (i + 1) % 3
        k = (i + 2) % 3
        s = math.sqrt(el[i][i] - el[j][j] - el[k][k] + 1.0)
        temp[i] = s * 0.5
        s = 0.5 / s
        temp[3] = (el[k][j] - el[j][k]) * s
        temp[j] = (el[j][i] + el[i][j]) * s
        temp[k] = (el[k][i] + el[i][k]) * s
    return (temp[0], temp[1], temp[2], temp[3])


def pose_roundtrip(q):
    """Quaternion as read back by getBasePositionAndOrientation (BaseAviary.py:517)."""
    return quaternion_from_matrix(matrix_from_quaternion(q))


def euler_from_quaternion(q):
    """(x,y,z,w) -> (roll, pitch, yaw); BaseAviary.py:518 call site."""
    x, y, z, w = float(q[0]), float(q[1]), float(q[2]), float(q[3])
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2 * (x * z - w * y)
    if sarg <= -0.99999:
        return (0.0, -0.5 * PYBULLET_PI, 2 * math.atan2(x, -y))
    if sarg >= 0.99999:
        return (0.0, 0.5 * PYBULLET_PI, 2 * math.atan2(-x, y))
    return (math.atan2(2 * (y * z + w * x), squ - sqx - sqy + sqz),
            math.asin(sarg),
            math.atan2(2 * (x * y + w * z), squ + sqx - sqy - sqz))


def quaternion_from_euler(rpy):
    """(roll, pitch, yaw) -> normalised (x,y,z,w); BaseAviary.py:488 call site."""
    phi, the, psi = float(rpy[0]) / 2.0, float(rpy[1]) / 2.0, float(rpy[2]) / 2.0
    q = [math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi),
         math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi),
         math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi),
         math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)]
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    return (q[0] / n, q[1] / n, q[2] / n, q[3] / n)


def as_array(t):
    return np.array(t, dtype=np.float64)
