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
    row 0 is the ego, the others the closest vehicles by |longitudinal distance on the ego's lane|
    within PERCEPTION_DISTANCE; columns are any of `Vehicle.to_dict`'s keys (vehicle/kinematics.py:237-261):
    presence, x, y, vx, vy, heading, cos_h, sin_h, cos_d, sin_d, long_off, lat_off, ang_off, with an
    optional `features_range` (every listed feature present in the table is mapped to [-1, 1])."""

    FEATURES = ["presence", "x", "y", "vx", "vy"]

    def __init__(self, features=None, vehicles_count: int = 5, features_range=None,
                 absolute: bool = False, order: str = "sorted", normalize: bool = True,
                 clip: bool = True, see_behind: bool = False, observe_intentions: bool = False,
                 include_obstacles: bool = True, **kwargs):
        self.features = list(features) if features else list(self.FEATURES)
        unknown = [f for f in self.features if f not in N.FEATURE_CODES]
        if unknown:
            raise KeyError(f"{unknown} not in index")  # what `df[self.features]` raises in the reference
        if len(self.features) > N.HWY_MAX_OBS_FEATURES:
            raise ValueError(f"at most {N.HWY_MAX_OBS_FEATURES} features")
        if observe_intentions and ("cos_d" in self.features or "sin_d" in self.features):
            # vehicles of this road family have no route: destination == position, the direction is (0, 0)
            pass
        self.features_range = None if features_range is None else {k: [float(v[0]), float(v[1])]
                                                                   for k, v in features_range.items()}
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
        if self.features == self.FEATURES and self.features_range is None:
            p.obs_n_features = 0  # default columns and ranges: the specialised epilogue
            return
        ranges = self.features_range
        if ranges is None:  # normalize_obs (observation.py:214-226): side lanes of the straight road = lanes_count
            ranges = {"x": [-5.0 * 40.0, 5.0 * 40.0], "y": [-4.0 * p.lanes_count, 4.0 * p.lanes_count],
                      "vx": [-2 * 40.0, 2 * 40.0], "vy": [-2 * 40.0, 2 * 40.0]}
        p.obs_n_features = len(self.features)
        for c, f in enumerate(self.features):
            p.obs_feature[c] = N.FEATURE_CODES[f]
            p.obs_feature_ranged[c] = int(f in ranges)
            if f in ranges:
                p.obs_feature_lo[c], p.obs_feature_hi[c] = ranges[f]


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
