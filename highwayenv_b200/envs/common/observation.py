"""Observation plugins of the batched backend (host-side descriptors).

Mirror of the reference's ``observation_factory``
(highway_env/envs/common/observation.py:772-794).  ``observe()`` itself is the epilogue of
the step kernel / ``hwy_highway_observe``.
"""
from __future__ import annotations

import numpy as np

from ... import _native as N
from ...spaces import Box


class ObservationType:
    def space(self):
        raise NotImplementedError

    def fill_params(self, p: N.HwyHighwayParams) -> None:
        raise NotImplementedError


class KinematicObservation(ObservationType):
    """Kinematics of the ego and its nearest vehicles (reference observation.py:155-276):
    rows (presence, x, y, vx, vy); row 0 is the ego, the others the closest vehicles by
    |longitudinal distance on the ego's lane| within PERCEPTION_DISTANCE."""

    FEATURES = ["presence", "x", "y", "vx", "vy"]

    def __init__(self, features=None, vehicles_count: int = 5, features_range=None,
                 absolute: bool = False, order: str = "sorted", normalize: bool = True,
                 clip: bool = True, see_behind: bool = False, observe_intentions: bool = False,
                 include_obstacles: bool = True, **kwargs):
        self.features = list(features) if features else list(self.FEATURES)
        if self.features != self.FEATURES:
            raise NotImplementedError(f"Kinematics features {self.features} (only {self.FEATURES})")
        if features_range is not None:
            raise NotImplementedError("custom features_range")
        if order != "sorted":
            raise NotImplementedError("order='shuffled' draws from env.np_random on the host")
        if not 1 <= int(vehicles_count) <= N.HWY_MAX_OBS_VEHICLES:
            raise ValueError(f"vehicles_count must be in 1..{N.HWY_MAX_OBS_VEHICLES}")
        self.vehicles_count = int(vehicles_count)
        self.absolute, self.normalize, self.clip = bool(absolute), bool(normalize), bool(clip)
        self.see_behind = bool(see_behind)

    def space(self):
        return Box(low=-np.inf, high=np.inf, shape=(self.vehicles_count, len(self.features)),
                   dtype=np.float32)

    def fill_params(self, p):
        p.obs_vehicles_count = self.vehicles_count
        p.obs_see_behind = int(self.see_behind)
        p.obs_absolute = int(self.absolute)
        p.obs_normalize = int(self.normalize)
        p.obs_clip = int(self.clip)


OBSERVATION_TYPES = {"Kinematics": KinematicObservation}
_KNOWN_UNSUPPORTED = (
    "TimeToCollision", "OccupancyGrid", "KinematicsGoal", "GrayscaleObservation",
    "AttributesObservation", "MultiAgentObservation", "TupleObservation", "LidarObservation",
    "ExitObservation",
)


def observation_factory(env, config: dict) -> ObservationType:
    kind = config["type"]
    if kind in OBSERVATION_TYPES:
        return OBSERVATION_TYPES[kind](**config)
    if kind in _KNOWN_UNSUPPORTED:
        raise NotImplementedError(f"observation type {kind!r} is not on the accelerated path yet")
    raise ValueError("Unknown observation type")
