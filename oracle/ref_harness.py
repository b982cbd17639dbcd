"""Live-reference harness — TEST INFRASTRUCTURE ONLY (never imported by the product).

Imports the UNMODIFIED reference package from ``/root/reference`` (read-only) under the
stub ``gymnasium``/``pygame``/``matplotlib`` in ``oracle/shim`` and exposes helpers to
(a) build a reference env, (b) dump its state in this repo's structure-of-arrays schema,
(c) step it.  It only works in the build container: ``/root/reference`` does not exist on
the GPU box, so nothing under ``tests -m gpu``, ``bench.py`` or ``smoke()`` may use it.
Golden fixtures are produced from it by ``oracle/gen_golden.py`` and committed under
``tests/golden/``.

Reference entry points driven here (file:line in /root/reference):
  * ``AbstractEnv.reset``  highway_env/envs/common/abstract.py:219-249
  * ``AbstractEnv.step``   highway_env/envs/common/abstract.py:259-285
  * ``Road.act/step``      highway_env/road/road.py:464-481
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("HWY_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "highway_env"))


def _ensure_imports() -> None:
    if "highway_env" in sys.modules:
        return
    try:  # a real gymnasium wins if it is ever installed
        importlib.import_module("gymnasium")
    except ModuleNotFoundError:
        sys.path.insert(0, _SHIM)
    for mod in ("pygame", "matplotlib"):
        try:
            importlib.import_module(mod)
        except ModuleNotFoundError:
            if _SHIM not in sys.path:
                sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    importlib.import_module("highway_env")


ENV_CLASSES = {
    "highway-v0": ("highway_env.envs.highway_env", "HighwayEnv"),
    "highway-fast-v0": ("highway_env.envs.highway_env", "HighwayEnvFast"),
    "intersection-v0": ("highway_env.envs.intersection_env", "IntersectionEnv"),
    "roundabout-v0": ("highway_env.envs.roundabout_env", "RoundaboutEnv"),
    "roundabout-v1": ("highway_env.envs.roundabout_env", "ConnectedLaneRoundaboutEnv"),
    "intersection-v1": ("highway_env.envs.intersection_env", "ContinuousIntersectionEnv"),
    "intersection-v2": ("highway_env.envs.intersection_env", "ConnectedLaneIntersectionEnv"),
    "intersection-multi-agent-v0": ("highway_env.envs.intersection_env", "MultiAgentIntersectionEnv"),
    "two-way-v0": ("highway_env.envs.two_way_env", "TwoWayEnv"),
    "u-turn-v0": ("highway_env.envs.u_turn_env", "UTurnEnv"),
    "u-turn-v1": ("highway_env.envs.u_turn_env", "ConnectedLaneUTurnEnv"),
    "exit-v0": ("highway_env.envs.exit_env", "ExitEnv"),
    "merge-v0": ("highway_env.envs.merge_env", "MergeEnv"),
    "merge-v1": ("highway_env.envs.merge_env", "ConnectedLaneMergeEnv"),
}


def make_reference_env(env_id: str, config: dict | None = None):
    """Construct the reference env (``gym.make`` equivalent without wrappers)."""
    _ensure_imports()
    mod, cls = ENV_CLASSES[env_id]
    return getattr(importlib.import_module(mod), cls)(config=config)


def lane_list(env):
    """Lanes in graph-enumeration order (road.py:65-71) with their indices."""
    out = []
    for _from, tos in env.road.network.graph.items():
        for _to, lanes in tos.items():
            for _id, lane in enumerate(lanes):
                out.append(((_from, _to, _id), lane))
    return out


NET_MAX_ROUTE = 16


def node_ids(env) -> dict:
    """Node name -> integer id: graph keys in insertion order, then sink nodes as they appear."""
    ids = {}
    g = env.road.network.graph
    for f in g.keys():
        ids.setdefault(f, len(ids))
    for f in g.keys():
        for t in g[f].keys():
            ids.setdefault(t, len(ids))
    return ids


def dump_network(env) -> dict:
    """Lane table in graph-enumeration order (road/road.py:65-71) + successor lists, as arrays."""
    from highway_env.road.lane import CircularLane, SineLane, StraightLane

    ids = node_ids(env)
    lanes = lane_list(env)
    idx_of = {li: k for k, (li, _) in enumerate(lanes)}
    n = len(lanes)
    f = {k: np.zeros(n, dtype=np.float64) for k in (
        "width", "speed_limit", "length", "sx", "sy", "ex", "ey", "dx", "dy", "lx", "ly", "heading",
        "amplitude", "pulsation", "phase", "cx", "cy", "radius", "start_phase", "end_phase", "direction")}
    i = {k: np.zeros(n, dtype=np.int32) for k in (
        "type", "from_node", "to_node", "lane_id", "road_first", "road_count", "forbidden", "priority",
        "exit_lane")}
    for k, ((fr, to, lid), lane) in enumerate(lanes):
        i["from_node"][k], i["to_node"][k], i["lane_id"][k] = ids[fr], ids[to], lid
        i["road_first"][k] = idx_of[(fr, to, 0)]
        i["road_count"][k] = len(env.road.network.graph[fr][to])
        i["forbidden"][k], i["priority"][k] = int(lane.forbidden), int(lane.priority)
        i["exit_lane"][k] = int("il" in fr and "o" in to)  # intersection_env.py:354-373
        f["width"][k], f["speed_limit"][k], f["length"][k] = lane.width, lane.speed_limit, lane.length
        if isinstance(lane, CircularLane):
            i["type"][k] = 2
            f["cx"][k], f["cy"][k] = lane.center
            f["radius"][k], f["start_phase"][k], f["end_phase"][k] = lane.radius, lane.start_phase, lane.end_phase
            f["direction"][k] = lane.direction
        elif isinstance(lane, StraightLane):
            i["type"][k] = 1 if isinstance(lane, SineLane) else 0
            f["sx"][k], f["sy"][k] = lane.start
            f["ex"][k], f["ey"][k] = lane.end
            f["dx"][k], f["dy"][k] = lane.direction
            f["lx"][k], f["ly"][k] = lane.direction_lateral
            f["heading"][k] = lane.heading
            if isinstance(lane, SineLane):
                f["amplitude"][k], f["pulsation"][k], f["phase"][k] = lane.amplitude, lane.pulsation, lane.phase
        else:
            raise NotImplementedError(type(lane))
    succ = np.full((len(ids), 6), -1, dtype=np.int32)
    succ_count = np.zeros(len(ids), dtype=np.int32)
    for fr, tos in env.road.network.graph.items():
        for to in tos.keys():
            succ[ids[fr], succ_count[ids[fr]]] = idx_of[(fr, to, 0)]
            succ_count[ids[fr]] += 1
    out = {"net_" + k: v for k, v in {**f, **i}.items()}
    out["net_succ"], out["net_succ_count"] = succ, succ_count
    out["net_node_names"] = np.array(list(ids.keys()))
    return out


def encode_route(env, vehicle) -> tuple:
    ids = node_ids(env)
    route = getattr(vehicle, "route", None) or []
    enc = np.zeros(NET_MAX_ROUTE, dtype=np.int32)
    assert len(route) <= NET_MAX_ROUTE, len(route)
    for k, (fr, to, lid) in enumerate(route):
        enc[k] = ids[fr] | (ids[to] << 8) | (((-1 if lid is None else int(lid)) + 1) << 16)
    return enc, len(route)


def dump_state(env, pad: int = 0) -> dict:
    """Snapshot the reference road in the SoA schema used by ``highwayenv_b200``.

    Fields (per vehicle, list order = ``road.vehicles`` order):
      x, y, heading, speed (f64); target_speed, timer, delta (f64, NaN when the vehicle
      class has none); lane, target_lane (index into the graph enumeration, -1 = none);
      crashed (bool); impact (f64[2], NaN when ``None``); speed_index (ego MDPVehicle).
    """
    idx_of = {li: k for k, (li, _) in enumerate(lane_list(env))}
    objects = [o for o in getattr(env.road, "objects", [])]
    vs = list(env.road.vehicles) + objects  # road.objects (Obstacles) occupy the slots after the vehicles
    n = len(vs)
    d = {
        "x": np.array([v.position[0] for v in vs], dtype=np.float64),
        "y": np.array([v.position[1] for v in vs], dtype=np.float64),
        "heading": np.array([float(v.heading) for v in vs], dtype=np.float64),
        "speed": np.array([float(v.speed) for v in vs], dtype=np.float64),
        "target_speed": np.array(
            [float(getattr(v, "target_speed", np.nan)) for v in vs], dtype=np.float64
        ),
        "timer": np.array([float(getattr(v, "timer", np.nan)) for v in vs], dtype=np.float64),
        "delta": np.array(
            [float(v.DELTA) if hasattr(v, "DELTA") else np.nan for v in vs], dtype=np.float64
        ),
        "lane": np.array([idx_of[tuple(v.lane_index)] for v in vs], dtype=np.int32),
        "target_lane": np.array(
            [
                idx_of[tuple(v.target_lane_index)] if hasattr(v, "target_lane_index") else -1
                for v in vs
            ],
            dtype=np.int32,
        ),
        "crashed": np.array([bool(v.crashed) for v in vs], dtype=np.bool_),
        "impact": np.array(
            [
                (np.nan, np.nan) if v.impact is None else (float(v.impact[0]), float(v.impact[1]))
                for v in vs
            ],
            dtype=np.float64,
        ).reshape(n, 2),
        "check_collisions": np.array([bool(v.check_collisions) for v in vs], dtype=np.bool_),
        "speed_index": np.array([int(getattr(v, "speed_index", -1)) for v in vs], dtype=np.int32),
        "time": np.float64(env.time),
        "steps": np.int64(env.steps),
    }
    if any(getattr(v, "enable_lane_change", True) is False for v in vs):  # IDMVehicle(enable_lane_change=False)
        d["no_lane_change"] = np.array([getattr(v, "enable_lane_change", True) is False for v in vs], dtype=np.int32)
    if any(getattr(v, "route", None) for v in vs):
        enc = [encode_route(env, v) for v in vs]
        d["route"] = np.stack([e[0] for e in enc])
        d["route_len"] = np.array([e[1] for e in enc], dtype=np.int32)
    if hasattr(env.road, "steps"):  # RegulatedRoad (road/regulation.py)
        d["road_steps"] = np.int64(env.road.steps)
        d["is_yielding"] = np.array([bool(getattr(v, "is_yielding", False)) for v in vs], dtype=np.bool_)
        # 1 MDPVehicle, 2 plain Vehicle / BicycleVehicle (ContinuousAction ego), 0 traffic
        d["kind"] = np.array([(1 if hasattr(v, "speed_index") else 2) if v in env.controlled_vehicles else 0
                              for v in vs], dtype=np.int32)
        if any(hasattr(v, "lateral_speed") for v in vs):  # BicycleVehicle (vehicle/dynamics.py:52-53)
            d["lat_speed"] = np.array([float(getattr(v, "lateral_speed", 0.0)) for v in vs], dtype=np.float64)
            d["yaw_rate"] = np.array([float(getattr(v, "yaw_rate", 0.0)) for v in vs], dtype=np.float64)
    if objects and "kind" not in d:
        d["kind"] = np.array([3 if v in objects else (1 if v in env.controlled_vehicles else 0) for v in vs], dtype=np.int32)
        d["is_yielding"] = np.zeros(n, dtype=np.bool_)
        d["road_steps"] = np.int64(0)
        d["count"] = np.int32(n)
    if pad:
        assert n <= pad, n
        d["count"] = np.int32(n)
        for k, a in list(d.items()):
            if isinstance(a, np.ndarray) and a.ndim >= 1 and a.shape[0] == n and k not in ("time",):
                fill = np.nan if a.dtype.kind == "f" else 0
                out = np.full((pad,) + a.shape[1:], fill, dtype=a.dtype)
                out[:n] = a
                d[k] = out
    return d


def rollout(env_id: str, config: dict | None, seed: int, actions, record_substeps: bool = False, pad: int = 0,
            mutate=None):
    """Reset with ``seed`` and apply ``actions``; returns dict of stacked arrays.

    Stepping continues after termination (the reference allows it), so trajectories have
    a fixed length ``len(actions)``; ``terminated``/``truncated`` are recorded per step.
    """
    env = make_reference_env(env_id, config)
    obs0, _ = env.reset(seed=seed)
    if mutate is not None:  # scripted edit of the reset state (to reach rare events within a short fixture)
        mutate(env)
        obs0 = env.observation_type.observe()
    states = [dump_state(env, pad)]
    rng_states = [env.np_random.bit_generator.state]
    obs, rew, term, trunc = [np.asarray(obs0)], [], [], []
    agents_rewards, agents_terminated = [], []
    for a in actions:
        if isinstance(a, np.ndarray) and a.ndim == 1 and a.dtype.kind == "i":
            a = tuple(int(x) for x in a)  # MultiAgentAction takes a tuple (action.py:316-321)
        o, r, te, tr, _info = env.step(a)
        if "agents_rewards" in _info:
            agents_rewards.append(np.array(_info["agents_rewards"], dtype=np.float64))
            agents_terminated.append(np.array(_info["agents_terminated"], dtype=np.bool_))
        states.append(dump_state(env, pad))
        rng_states.append(env.np_random.bit_generator.state)
        obs.append(np.asarray(o))
        rew.append(float(r))
        term.append(bool(te))
        trunc.append(bool(tr))
    keys = [k for k in states[0].keys()]
    out = {k: np.stack([s[k] for s in states]) for k in keys}
    out["obs"] = np.stack(obs)
    out["reward"] = np.array(rew, dtype=np.float64)
    out["terminated"] = np.array(term, dtype=np.bool_)
    out["truncated"] = np.array(trunc, dtype=np.bool_)
    out["actions"] = np.asarray(actions)
    if agents_rewards:
        out["agents_rewards"], out["agents_terminated"] = np.stack(agents_rewards), np.stack(agents_terminated)
    m64 = (1 << 64) - 1
    out["rng_words"] = np.array([[st["state"]["state"] >> 64, st["state"]["state"] & m64, st["state"]["inc"] >> 64,
                                  st["state"]["inc"] & m64, (int(st["has_uint32"]) << 32) | int(st["uinteger"])]
                                 for st in rng_states], dtype=np.uint64)
    return out
