"""Cross-check the restated Bullet helpers (oracle/bullet_math.py) against scipy."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import bullet_math as bm
from oracle.aviary_oracle import integrate_q

RNG = np.random.default_rng(123)


def random_quats(n):
    q = RNG.standard_normal((n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def test_matrix_from_quaternion_matches_scipy():
    for q in random_quats(200):
        m = np.array(bm.matrix_from_quaternion(q)).reshape(3, 3)
        assert np.allclose(m, Rotation.from_quat(q).as_matrix(), atol=4e-15)
    # non-unit input: s = 2/|q|^2 normalises
    q = np.array([0.2, -0.4, 0.1, 1.7])
    assert np.allclose(np.array(bm.matrix_from_quaternion(q)).reshape(3, 3),
                       Rotation.from_quat(q / np.linalg.norm(q)).as_matrix(), atol=4e-15)


def test_euler_roundtrip_matches_scipy():
    for _ in range(200):
        rpy = RNG.uniform([-np.pi, -1.5, -np.pi], [np.pi, 1.5, np.pi])
        q = np.array(bm.quaternion_from_euler(rpy))
        assert np.allclose(q, Rotation.from_euler('xyz', rpy).as_quat(), atol=4e-15) or \
            np.allclose(-q, Rotation.from_euler('xyz', rpy).as_quat(), atol=4e-15)
        assert np.allclose(bm.euler_from_quaternion(q), Rotation.from_quat(q).as_euler('xyz'), atol=2e-13)


def test_euler_gimbal_branches():
    for sign in (+1, -1):
        q = bm.quaternion_from_euler([0.3, sign * (np.pi / 2 - 1e-7), -0.4])
        roll, pitch, yaw = bm.euler_from_quaternion(q)
        assert roll == 0.0 and pitch == sign * 0.5 * bm.PYBULLET_PI
        # the sum/difference roll -+ yaw is what survives at the singularity
        assert np.isfinite(yaw)


def test_pose_roundtrip_normalises_and_canonicalises():
    for q in random_quats(300):
        for scale in (1.0, 1.0 + 3e-9, 0.7):
            r = np.array(bm.pose_roundtrip(q * scale))
            assert abs(np.linalg.norm(r) - 1.0) < 4e-15
            assert np.allclose(r, q, atol=1e-14) or np.allclose(r, -q, atol=1e-14)
            m = np.array(bm.matrix_from_quaternion(q)).reshape(3, 3)
            if np.trace(m) > 0:
                assert r[3] > 0           # trace > 0 branch returns w > 0
            else:
                i = int(np.argmax(np.diag(m)))
                assert r[i] > 0           # pivot component is the positive square root


def test_pose_roundtrip_vs_scipy_canonical_in_the_trace_nonpositive_region():
    """rotations by more than 120 degrees (trace <= 0, where btMatrix3x3::getRotation pivots on the
    largest diagonal element).  The round trip must be the same ROTATION as scipy's canonical (w >= 0) quaternion; the
    SIGN is Bullet's own convention — pivot component positive — which differs from scipy's exactly when the w that
    follows from it is negative.  Both facts are asserted, so a sign slip in the restatement cannot hide."""
    n_neg_w = 0
    for _ in range(2000):
        axis = RNG.standard_normal(3)
        axis /= np.linalg.norm(axis)
        ang = RNG.uniform(2.2, np.pi)                       # > 120 degrees: trace = 1 + 2 cos(ang) <= 0 mostly
        q = np.concatenate([axis * np.sin(ang / 2), [np.cos(ang / 2)]]) * RNG.choice([-1.0, 1.0])
        m = np.array(bm.matrix_from_quaternion(q)).reshape(3, 3)
        if np.trace(m) > 0:
            continue
        r = np.array(bm.pose_roundtrip(q))
        canon = Rotation.from_matrix(m).as_quat(canonical=True)
        assert np.allclose(Rotation.from_quat(r).as_matrix(), m, atol=1e-14)
        i = int(np.argmax(np.diag(m)))
        assert r[i] > 0
        if r[3] >= 0:
            assert np.allclose(r, canon, atol=1e-13)
        else:
            n_neg_w += 1
            assert np.allclose(r, -canon, atol=1e-13)
    assert n_neg_w > 100            # the region where Bullet's and scipy's conventions differ is really visited


def test_integrate_q_is_body_frame_exponential():
    for q in random_quats(100):
        w = RNG.standard_normal(3) * RNG.choice([1e-3, 1.0, 30.0])
        dt = 1 / 240
        got = integrate_q(q, w, dt)
        want = (Rotation.from_quat(q) * Rotation.from_rotvec(w * dt)).as_quat()
        assert np.allclose(got, want, atol=1e-14) or np.allclose(got, -want, atol=1e-14)
        assert abs(np.linalg.norm(got) - 1) < 1e-14


def test_integrate_q_early_out():
    q = random_quats(1)[0]
    assert integrate_q(q, np.array([0.0, 0.0, 9e-9]), 1 / 240) is q      # np.isclose(norm, 0): atol 1e-8
    assert integrate_q(q, np.array([0.0, 0.0, 2e-8]), 1 / 240) is not q
