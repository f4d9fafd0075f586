"""TEST INFRASTRUCTURE — run the UNMODIFIED reference envs without PyBullet.

Used only by `tests/golden/make_golden.py` (in the build container, where
`/root/reference` is mounted) to generate golden trajectories: the real
`BaseAviary.step/_dynamics/_integrateQ`, `BaseRLAviary._preprocessAction/
_computeObs`, `MultiHoverAviary`, `HoverAviary` and `SpiralFormationAviary`
code is imported from `/root/reference` and executed with `Physics.DYN`; the
three absent third-party modules are replaced in `sys.modules` by shims:

* `pybullet`      -> `_FakeBullet`: a per-client pose/velocity store whose
  quaternion utilities are `oracle/bullet_math.py` (PARITY UNPINNED: restated
  from Bullet's published sources, see that file).  In DYN mode the reference
  never calls `stepSimulation` (`BaseAviary.py:369-370`), so a state store plus
  closed-form helpers is everything the path needs from PyBullet.
  `applyExternalForce/Torque` calls are recorded so the reference's own
  `_groundEffect/_drag/_downwash` (`BaseAviary.py:715-811`) can be sampled.
* `pybullet_data` -> `getDataPath()` only.
* `gymnasium`     -> `Env`, `spaces.Box`, `envs.registration.register`.

Nothing here is imported by the product, the gpu tests, `smoke()` or `bench.py`.
"""
import importlib
import os
import sys
import types
import xml.etree.ElementTree as ET

import numpy as np

from . import bullet_math as bm

REFERENCE_ROOT = os.environ.get("BD_REFERENCE_ROOT", "/root/reference")


class _Body:
    def __init__(self, urdf, pos, quat):
        self.urdf = urdf
        self.pos = tuple(float(v) for v in pos)
        self.quat = tuple(float(v) for v in quat)
        self.lin = (0.0, 0.0, 0.0)
        self.ang = (0.0, 0.0, 0.0)
        self.link_offsets = []
        if urdf and os.path.exists(urdf) and "plane" not in os.path.basename(urdf):
            root = ET.parse(urdf).getroot()
            for link in root.findall("link")[1:]:
                org = link.find("inertial").find("origin")
                self.link_offsets.append(tuple(float(s) for s in org.get("xyz").split()))


class _FakeBullet(types.ModuleType):
    """The slice of the pybullet module API the DYN path touches."""

    DIRECT = 2
    GUI = 1
    LINK_FRAME = 1
    WORLD_FRAME = 2
    URDF_USE_INERTIA_FROM_FILE = 2
    ER_TINY_RENDERER = 0
    ER_SEGMENTATION_MASK_OBJECT_AND_LINKINDEX = 0
    ER_NO_SEGMENTATION_MASK = 0
    COV_ENABLE_RGB_BUFFER_PREVIEW = 0
    COV_ENABLE_DEPTH_BUFFER_PREVIEW = 0
    COV_ENABLE_SEGMENTATION_MARK_PREVIEW = 0
    STATE_LOGGING_VIDEO_MP4 = 0

    def __init__(self):
        super().__init__("pybullet")
        self._clients = {}
        self._next_client = 0
        self.force_log = []  # (client, body, link, force(3), frame) of applyExternalForce

    # -- session -------------------------------------------------------------
    def connect(self, mode, **kw):
        cid = self._next_client
        self._next_client += 1
        self._clients[cid] = {}
        return cid

    def disconnect(self, physicsClientId=0):
        self._clients.pop(physicsClientId, None)

    def resetSimulation(self, physicsClientId=0):
        self._clients[physicsClientId] = {}

    def setGravity(self, *a, **kw):
        pass

    def setRealTimeSimulation(self, *a, **kw):
        pass

    def setTimeStep(self, *a, **kw):
        pass

    def setAdditionalSearchPath(self, *a, **kw):
        pass

    def stepSimulation(self, *a, **kw):
        raise RuntimeError("rigid-body stepping is not part of the DYN path")

    # -- bodies ----------------------------------------------------------------
    def loadURDF(self, fileName, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1),
                 flags=0, physicsClientId=0, **kw):
        bodies = self._clients[physicsClientId]
        bid = len(bodies)
        bodies[bid] = _Body(fileName, basePosition, baseOrientation)
        return bid

    def getBasePositionAndOrientation(self, bodyUniqueId, physicsClientId=0):
        b = self._clients[physicsClientId][int(bodyUniqueId)]
        return b.pos, bm.pose_roundtrip(b.quat)

    def getBaseVelocity(self, bodyUniqueId, physicsClientId=0):
        b = self._clients[physicsClientId][int(bodyUniqueId)]
        return b.lin, b.ang

    def resetBasePositionAndOrientation(self, bodyUniqueId, posObj, ornObj, physicsClientId=0):
        b = self._clients[physicsClientId][int(bodyUniqueId)]
        b.pos = tuple(float(v) for v in posObj)
        b.quat = tuple(float(v) for v in ornObj)

    def resetBaseVelocity(self, objectUniqueId, linearVelocity=None, angularVelocity=None,
                          physicsClientId=0):
        b = self._clients[physicsClientId][int(objectUniqueId)]
        if linearVelocity is not None:
            b.lin = tuple(float(v) for v in linearVelocity)
        if angularVelocity is not None:
            b.ang = tuple(float(v) for v in angularVelocity)

    def getLinkStates(self, bodyUniqueId, linkIndices, computeLinkVelocity=0,
                      computeForwardKinematics=0, physicsClientId=0):
        b = self._clients[physicsClientId][int(bodyUniqueId)]
        pos, quat = self.getBasePositionAndOrientation(bodyUniqueId, physicsClientId)
        m = np.array(bm.matrix_from_quaternion(quat)).reshape(3, 3)
        out = []
        for li in linkIndices:
            off = np.array(b.link_offsets[li])
            wp = tuple(np.array(pos) + m @ off)
            out.append((wp, quat, off, (0, 0, 0, 1), wp, quat, b.lin, b.ang))
        return out

    def applyExternalForce(self, objectUniqueId, linkIndex, forceObj, posObj, flags,
                           physicsClientId=0):
        self.force_log.append((physicsClientId, int(objectUniqueId), int(linkIndex),
                               tuple(float(v) for v in forceObj), flags))

    def applyExternalTorque(self, *a, **kw):
        pass

    # -- closed-form helpers -----------------------------------------------------
    @staticmethod
    def getMatrixFromQuaternion(q):
        return bm.matrix_from_quaternion(q)

    @staticmethod
    def getEulerFromQuaternion(q):
        return bm.euler_from_quaternion(q)

    @staticmethod
    def getQuaternionFromEuler(rpy):
        return bm.quaternion_from_euler(rpy)


class _Box:
    """`gymnasium.spaces.Box` shim (attributes only)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)


class _Env:
    metadata = {}

    def close(self):
        pass


_INSTALLED = None


def install():
    """Install the shims and make `gym_pybullet_drones` importable. Idempotent."""
    global _INSTALLED
    if _INSTALLED is not None:
        return _INSTALLED
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_pybullet_drones")):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    fake = _FakeBullet()
    sys.modules["pybullet"] = fake
    pd = types.ModuleType("pybullet_data")
    pd.getDataPath = lambda: "/nonexistent/pybullet_data"
    sys.modules["pybullet_data"] = pd

    gym = types.ModuleType("gymnasium")
    gym.Env = _Env
    gym.Wrapper = object
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = _Box
    gym.spaces = spaces
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda **kw: None
    envs.registration = reg
    gym.envs = envs
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces,
                        "gymnasium.envs": envs, "gymnasium.envs.registration": reg})

    # Package stubs: the reference's own `envs/__init__.py` imports aviaries that
    # need transforms3d / firmware bindings; bypass the __init__ files and load
    # the individual, unmodified submodules from the read-only tree.
    pkg_dir = os.path.join(REFERENCE_ROOT, "gym_pybullet_drones")
    for name, sub in (("gym_pybullet_drones", ""), ("gym_pybullet_drones.envs", "envs"),
                      ("gym_pybullet_drones.utils", "utils"),
                      ("gym_pybullet_drones.control", "control")):
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(pkg_dir, sub)]
        sys.modules[name] = mod

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pkg_resources

    def _resource_filename(package, resource):
        assert package == "gym_pybullet_drones"
        return os.path.join(pkg_dir, resource)

    pkg_resources.resource_filename = _resource_filename
    _INSTALLED = fake
    return fake


def reference_classes():
    """-> dict of the reference's env classes and enums (unmodified code)."""
    install()
    enums = importlib.import_module("gym_pybullet_drones.utils.enums")
    return {
        "enums": enums,
        "BaseAviary": importlib.import_module("gym_pybullet_drones.envs.BaseAviary").BaseAviary,
        "HoverAviary": importlib.import_module("gym_pybullet_drones.envs.HoverAviary").HoverAviary,
        "MultiHoverAviary": importlib.import_module(
            "gym_pybullet_drones.envs.MultiHoverAviary").MultiHoverAviary,
        "SpiralFormationAviary": importlib.import_module(
            "gym_pybullet_drones.envs.SpiralAviary").SpiralFormationAviary,
        "MeetupAviary": importlib.import_module("gym_pybullet_drones.envs.MeetupAviary").MeetupAviary,
        "FlockAviary": importlib.import_module("gym_pybullet_drones.envs.FlockAviary").FlockAviary,
        "LeaderFollowerAviary": importlib.import_module(
            "gym_pybullet_drones.envs.LeaderFollowerAviary").LeaderFollowerAviary,
    }
