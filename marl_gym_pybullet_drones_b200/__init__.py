"""B200-native batched quadrotor simulator for the Physics.DYN hot path of
marl-gym-pybullet-drones.

Import map (reference name -> here):

* `gym_pybullet_drones.utils.enums`            -> `.enums`
* `BaseAviary` constants / URDF parser          -> `.constants`
* N aviaries inside `SubprocVecEnv`             -> `.batch_aviary.BatchAviary` (device tensors)
* `HoverAviary`, `MultiHoverAviary`, `SpiralFormationAviary`, `MeetupAviary`, `FlockAviary`,
  `LeaderFollowerAviary` (Gymnasium API) -> `.envs`
* `make_vec_envs`, `VecRecordEpisodeStatistics` -> `.vec_env`

Importing the package does not need a GPU or torch; constructing an aviary does
(there is no CPU fallback — the CUDA library must be built, see `.build`).
"""
from .enums import ActionType, DroneModel, ImageType, ObservationType, Physics  # noqa: F401
from .constants import DroneConstants, drone_constants, parse_urdf  # noqa: F401

__all__ = [
    "ActionType", "DroneModel", "ImageType", "ObservationType", "Physics",
    "DroneConstants", "drone_constants", "parse_urdf",
    "BatchAviary", "StepResult", "HoverAviary", "MultiHoverAviary", "SpiralFormationAviary",
    "MeetupAviary", "FlockAviary", "LeaderFollowerAviary",
    "BatchVecEnv", "VecRecordEpisodeStatistics", "make_vec_envs", "DeviceMAPPO",
]

_LAZY = {
    "BatchAviary": ".batch_aviary", "StepResult": ".batch_aviary",
    "HoverAviary": ".envs", "MultiHoverAviary": ".envs", "SpiralFormationAviary": ".envs",
    "MeetupAviary": ".envs", "FlockAviary": ".envs", "LeaderFollowerAviary": ".envs",
    "BatchVecEnv": ".vec_env", "VecRecordEpisodeStatistics": ".vec_env", "make_vec_envs": ".vec_env",
    "DeviceMAPPO": ".mappo",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(_LAZY[name], __name__), name)
    raise AttributeError(name)
