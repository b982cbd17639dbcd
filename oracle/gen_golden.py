"""Generate golden fixtures from the UNMODIFIED Python reference (build container only).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Each fixture is a seeded rollout of the reference env (ref_harness.rollout): the full
per-vehicle state after reset and after every env.step, plus obs / reward / terminated /
truncated and the action sequence.  Stepping continues past termination so all
trajectories have fixed length.  tests/test_oracle_golden.py pins the C oracle to these;
the `-m gpu` tests pin the CUDA path to them (teacher-forced and free-running on the
well-conditioned prefix, see DESIGN.md "parity protocol").
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (env_id, config override, seeds, n_steps, action kind)
CASES = {
    # BASELINE.json configs[0]: highway-fast-v0 defaults (V = 21)
    "highway_fast_v20": ("highway-fast-v0", None, list(range(32)), 30, "discrete5"),
    # configs[1]: highway-fast-v0, vehicles_count = 50 (V = 51)
    "highway_fast_v50": ("highway-fast-v0", {"vehicles_count": 50}, list(range(100, 132)), 30, "discrete5"),
    # highway-v0 defaults: all-pairs collisions, 15 substeps, 4 lanes
    "highway_v50": ("highway-v0", None, list(range(200, 232)), 20, "discrete5"),
    # configs[4] shape: highway-v0, vehicles_count = 100, ContinuousAction (V = 101)
    "highway_v100_continuous": (
        "highway-v0",
        {"vehicles_count": 100, "action": {"type": "ContinuousAction"}},
        list(range(300, 316)),
        12,
        "box2",
    ),
    # (f)2 plugins on the same path: DiscreteAction (action.py:165-196) and the full Kinematics feature list
    "highway_discrete_action": ("highway-v0", {"vehicles_count": 20, "action": {"type": "DiscreteAction"}},
                                list(range(800, 832)), 15, "discrete9"),
    "highway_fast_features": (
        "highway-fast-v0",
        {"observation": {"type": "Kinematics", "vehicles_count": 7, "see_behind": True,
                         "features": ["presence", "x", "y", "vx", "vy", "heading", "cos_h", "sin_h", "cos_d", "sin_d",
                                      "long_off", "lat_off", "ang_off"]}},
        list(range(810, 842)), 20, "discrete5"),
    "highway_fast_features_range": (
        "highway-fast-v0",
        {"observation": {"type": "Kinematics", "vehicles_count": 4, "absolute": True, "clip": False,
                         "features": ["x", "lat_off", "presence", "vx", "heading", "long_off"],
                         "features_range": {"x": [-100, 1500], "vx": [0, 45], "long_off": [0, 2000],
                                            "heading": [-1, 1], "vy": [-3, 3]}}},
        list(range(820, 852)), 20, "discrete5"),
    # roundabout-v0 defaults (Kinematics absolute) and BASELINE configs[3] shape (TimeToCollision)
    "roundabout_kin": ("roundabout-v0", None, list(range(400, 432)), 11, "discrete5"),
    "roundabout_ttc": ("roundabout-v0", {"observation": {"type": "TimeToCollision", "horizon": 10}},
                       list(range(500, 532)), 11, "discrete5"),
    # intersection-v0 defaults (Kinematics, 7 features) and BASELINE configs[2] shape (OccupancyGrid)
    "intersection_kin": ("intersection-v0", None, list(range(600, 632)), 13, "discrete3"),
    "intersection_grid": ("intersection-v0", {"observation": {"type": "OccupancyGrid"}},
                          list(range(700, 732)), 13, "discrete3"),
    # (f)3 connected-lane neighbour search (ConnectedLaneNeighboursMixin, abstract.py:26-37; road.py:509-529)
    "roundabout_v1_kin": ("roundabout-v1", None, list(range(900, 932)), 11, "discrete5"),
    "intersection_v2_kin": ("intersection-v2", None, list(range(910, 942)), 13, "discrete3"),
    # (f)2 MultiAgentAction / MultiAgentObservation: two controlled vehicles
    "intersection_multi_agent": ("intersection-multi-agent-v0", None, list(range(920, 952)), 13, "discrete3x2"),
    # (f)3 scenario builders on the same kernels: merge-v0 (straight + sine lanes, an Obstacle at the ramp's end)
    "merge_kin": ("merge-v0", None, list(range(930, 962)), 18, "discrete5"),
    "merge_v1_kin": ("merge-v1", None, list(range(940, 972)), 18, "discrete5"),
    # two-way-v0: oncoming traffic on ("b","a",0), IDM vehicles with enable_lane_change=False, TimeToCollision horizon 5
    "two_way_ttc": ("two-way-v0", None, list(range(960, 992)), 15, "discrete5"),
    # u-turn-v0: circular U-turn, routed traffic, ego with PURSUIT_TAU = TAU_HEADING, TimeToCollision horizon 16
    "u_turn_ttc": ("u-turn-v0", None, list(range(970, 1002)), 10, "discrete5"),
    "u_turn_v1_ttc": ("u-turn-v1", None, list(range(980, 1012)), 10, "discrete5"),
    # intersection-v1 (ContinuousIntersectionEnv): ContinuousAction with the dynamical BicycleVehicle, 8-column Kinematics;
    # and intersection-v0 with a kinematic ContinuousAction / DiscreteAction ego (plain Vehicle under RegulatedRoad)
    "intersection_v1": ("intersection-v1", None, list(range(1200, 1232)), 13, "box2"),
    "intersection_continuous": ("intersection-v0", {"action": {"type": "ContinuousAction", "longitudinal": True,
                                                               "lateral": True}},
                                list(range(1240, 1272)), 13, "box2"),
    # exit-v0: three highway sections (6 / 7 / 6 lanes) with an exit ramp, routed traffic without lane changes,
    # ExitObservation, goal reward on the exit lane
    "exit_obs": ("exit-v0", None, list(range(1100, 1132)), 18, "discrete5"),
    # the merging vehicle is moved onto the end of the ramp at speed: it runs into the Obstacle (objects.py:104-107)
    "merge_obstacle_hit": ("merge-v0", None, list(range(950, 982)), 6, "discrete5"),
}


def _ram_the_obstacle(env):
    lane = env.road.network.get_lane(("b", "c", 2))
    v = env.road.vehicles[4]
    v.position = lane.position(58.0 + 3.0 * (env.np_random.uniform()), 0.0)
    v.heading = lane.heading_at(60.0)
    v.speed, v.target_speed = 28.0, 30.0
    v.lane_index = v.target_lane_index = ("b", "c", 2)
    v.lane = lane


MUTATE = {"merge_obstacle_hit": _ram_the_obstacle}


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, (env_id, over, seeds, T, akind) in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
        per_seed = []
        for seed in seeds:
            if akind == "discrete5":
                actions = rng.integers(0, 5, size=T).astype(np.int64)
            elif akind == "discrete9":
                actions = rng.integers(0, 9, size=T).astype(np.int64)
            elif akind == "discrete3x2":
                actions = rng.integers(0, 3, size=(T, 2)).astype(np.int64)
            elif akind == "discrete3":
                actions = rng.integers(0, 3, size=T).astype(np.int64)
            else:
                actions = rng.uniform(-1, 1, size=(T, 2)).astype(np.float32)
            per_seed.append(rh.rollout(env_id, over, seed, list(actions), pad=32 if env_id.startswith("intersection") else 0,
                                       mutate=MUTATE.get(name)))
        out = {k: np.stack([p[k] for p in per_seed]) for k in per_seed[0].keys()}
        out["seeds"] = np.array(seeds, dtype=np.int64)
        env = rh.make_reference_env(env_id, over)
        import json

        cfg = dict(env.config)
        if not env_id.startswith("highway"):
            env.reset(seed=0)
            out.update(rh.dump_network(env))
        cfg["_others_check_collisions"] = 0 if env_id == "highway-fast-v0" else 1
        cfg["_env_id"] = env_id
        if env_id.startswith("merge"):
            lanes = [li for li, _ in rh.lane_list(env)]
            cfg["_merge_lane"] = lanes.index(("b", "c", 2))
            cfg["_default_side_lanes"] = len(env.road.network.all_side_lanes(env.vehicle.lane_index))
        if env_id.startswith("exit"):
            lanes = [li for li, _ in rh.lane_list(env)]
            n_l = int(cfg["lanes_count"])
            cfg["_exit_lane_a"], cfg["_exit_lane_b"] = lanes.index(("1", "2", n_l)), lanes.index(("2", "exit", 0))
            cfg["_obs_exit_lane"] = lanes.index(("1", "2", n_l))  # get_lane(("1", "2", -1)): the last lane of the road
            cfg["_default_side_lanes"] = n_l  # the controlled vehicle spawns on ("0", "1", 0)
        out["config_json"] = np.array(json.dumps(cfg))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(seeds)} seeds x {T} steps -> {path} "
              f"({os.path.getsize(path)/1e3:.0f} kB, {time.time()-t0:.1f}s)")

    # reset-only fixtures: pins the numpy-PCG64 spawn restatement on many seeds
    for name, (env_id, over) in {} if only else {
        "reset_highway_fast_v50": ("highway-fast-v0", {"vehicles_count": 50}),
        "reset_highway_v100": ("highway-v0", {"vehicles_count": 100, "action": {"type": "ContinuousAction"}}),
    }.items():
        env = rh.make_reference_env(env_id, over)
        seeds = list(range(1000, 1032))
        states, obs = [], []
        for seed in seeds:
            o, _ = env.reset(seed=seed)
            states.append(rh.dump_state(env))
            obs.append(o)
            # a second reset WITHOUT seed continues the stream (gymnasium autoreset)
            o2, _ = env.reset()
            states.append(rh.dump_state(env))
            obs.append(o2)
        out = {k: np.stack([s[k] for s in states]) for k in states[0].keys()}
        out["obs"] = np.stack(obs)
        out["seeds"] = np.array(seeds, dtype=np.int64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: -> {path} ({os.path.getsize(path)/1e3:.0f} kB)")


# ------------------------------------------------------------------ observation plugins on every env family
# One short rollout per env id; on every visited state the reference builds each listed observation type with its own
# observation_factory (envs/common/observation.py:772-794) and observes — the plugin registry is orthogonal to the env.
GRID_RICH = {"type": "OccupancyGrid",
             "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h", "long_off", "lat_off", "ang_off", "heading",
                          "cos_d", "sin_d", "on_road"],
             "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
             "grid_size": [[-32, 32], [-18, 18]], "grid_step": [4, 3], "align_to_vehicle_axes": True, "clip": False}
OBS_PLUGIN_CASES = {
    "obs_plugins_highway": ("highway-v0", {"vehicles_count": 30}, list(range(2000, 2004)), 8, "discrete5", [
        {"type": "TimeToCollision", "horizon": 10},
        {"type": "OccupancyGrid"},
        {"type": "OccupancyGrid", "grid_size": [[-300, 300], [-10, 10]], "grid_step": [2, 2]},  # the reference's own test
        GRID_RICH,
        {"type": "LidarObservation"},
        {"type": "LidarObservation", "cells": 36, "maximum_range": 90, "normalize": False},
    ]),
    "obs_plugins_intersection": ("intersection-v0", None, list(range(2010, 2014)), 8, "discrete3", [
        {"type": "TimeToCollision", "horizon": 5},
        {"type": "OccupancyGrid", "align_to_vehicle_axes": True, "grid_size": [[-32, 32], [-32, 32]], "grid_step": [4, 4]},
        GRID_RICH,
        {"type": "LidarObservation"},
    ]),
    "obs_plugins_roundabout": ("roundabout-v0", None, list(range(2020, 2024)), 8, "discrete5", [
        {"type": "OccupancyGrid"},
        GRID_RICH,
        {"type": "LidarObservation", "cells": 24},
        {"type": "TimeToCollision", "horizon": 7},
    ]),
    "obs_plugins_merge": ("merge-v0", None, list(range(2030, 2034)), 8, "discrete5", [
        {"type": "LidarObservation", "maximum_range": 120},
        {"type": "OccupancyGrid", "grid_size": [[-60, 60], [-12, 12]], "grid_step": [3, 2]},
        {"type": "TimeToCollision", "horizon": 6},
    ]),
}


def gen_obs_plugins(only) -> None:
    import json

    from highway_env.envs.common.observation import observation_factory

    for name, (env_id, over, seeds, T, akind, obs_cfgs) in OBS_PLUGIN_CASES.items():
        if only and name not in only:
            continue
        rng = np.random.default_rng(sum(map(ord, name)))
        pad = 32 if env_id.startswith("intersection") else 0
        states, obs = [], [[] for _ in obs_cfgs]
        for seed in seeds:
            env = rh.make_reference_env(env_id, over)
            env.reset(seed=seed)
            observers = [observation_factory(env, dict(c)) for c in obs_cfgs]
            hi = 3 if akind == "discrete3" else 5
            for t in range(T + 1):
                states.append(rh.dump_state(env, pad))
                for k, ob in enumerate(observers):
                    obs[k].append(np.asarray(ob.observe()).copy())
                if t < T:
                    env.step(int(rng.integers(0, hi)))
        keys = [k for k in states[0].keys() if all(k in s for s in states)]
        out = {k: np.stack([s[k] for s in states]) for k in keys}
        for k in range(len(obs_cfgs)):
            out[f"obs_{k}"] = np.stack(obs[k])
        env = rh.make_reference_env(env_id, over)
        env.reset(seed=0)
        out.update(rh.dump_network(env))
        cfg = dict(env.config)
        cfg["_env_id"] = env_id
        cfg["_target_speeds"] = [float(x) for x in env.vehicle.target_speeds]
        out["config_json"] = np.array(json.dumps(cfg))
        out["obs_cfgs_json"] = np.array(json.dumps(obs_cfgs))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(states)} states x {len(obs_cfgs)} observation types -> {path} ({os.path.getsize(path)/1e3:.0f} kB)")


if __name__ == "__main__":
    if not rh.reference_available():
        raise SystemExit("reference not mounted; golden fixtures can only be generated in the build container")
    args = sys.argv[1:]
    if args and args[0] == "obs_plugins":
        rh._ensure_imports()
        gen_obs_plugins(args[1:])
    else:
        main()
