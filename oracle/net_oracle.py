"""ctypes binding of oracle/net_oracle.c (general road network: roundabout-v0) plus the numpy
restatement of RoundaboutEnv._make_vehicles.  TEST INFRASTRUCTURE ONLY (see hwy_oracle.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnet_oracle.so")

NET_MAX_LANES, NET_MAX_NODES, NET_MAX_SUCC, NET_MAX_ROUTE, NET_MAX_TARGET_SPEEDS = 64, 64, 6, 16, 8
OBS_KINEMATICS, OBS_TTC = 0, 2
KIND_IDM, KIND_MDP = 0, 1

_LANE_I = ("type", "from_node", "to_node", "lane_id", "road_first", "road_count", "forbidden", "priority")
_LANE_F = ("width", "speed_limit", "length", "sx", "sy", "ex", "ey", "dx", "dy", "lx", "ly", "heading",
           "amplitude", "pulsation", "phase", "cx", "cy", "radius", "start_phase", "end_phase", "direction")


class NetLane(C.Structure):
    _fields_ = [(k, C.c_int32) for k in _LANE_I] + [(k, C.c_double) for k in _LANE_F]


class NetGraph(C.Structure):
    _fields_ = [
        ("n_lanes", C.c_int32), ("n_nodes", C.c_int32),
        ("lanes", NetLane * NET_MAX_LANES),
        ("succ_count", C.c_int32 * NET_MAX_NODES),
        ("succ", (C.c_int32 * NET_MAX_SUCC) * NET_MAX_NODES),
    ]


class NetCfg(C.Structure):
    _fields_ = (
        [(k, C.c_int32) for k in (
            "n_vehicles", "simulation_frequency", "policy_frequency", "n_target_speeds", "obs_type",
            "obs_vehicles_count", "obs_see_behind", "obs_absolute", "obs_normalize", "obs_clip",
            "ttc_horizon", "normalize_reward")]
        + [("duration", C.c_double), ("target_speeds", C.c_double * NET_MAX_TARGET_SPEEDS)]
        + [(k, C.c_double) for k in (
            "obs_x_lo", "obs_x_hi", "obs_y_lo", "obs_y_hi", "obs_vx_lo", "obs_vx_hi", "obs_vy_lo", "obs_vy_hi",
            "collision_reward", "high_speed_reward", "lane_change_reward",
            "acc_max", "comfort_acc_max", "comfort_acc_min", "distance_wanted", "time_wanted",
            "politeness", "lane_change_min_acc_gain", "lane_change_max_braking_imposed", "lane_change_delay",
            "perception_distance")]
    )


_SF = ("x", "y", "heading", "speed", "target_speed", "timer", "delta", "impact_x", "impact_y")
_SI = ("lane", "target_lane", "kind", "crashed", "has_impact", "check_collisions")


class NetState(C.Structure):
    _fields_ = ([(k, C.c_void_p) for k in _SF] + [(k, C.c_void_p) for k in _SI]
                + [("route", C.c_void_p), ("route_len", C.c_void_p), ("speed_index", C.c_void_p),
                   ("time", C.c_void_p)])


def build(force: bool = False) -> str:
    src, hdr = os.path.join(_HERE, "net_oracle.c"), os.path.join(_HERE, "net_oracle.h")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=gnu11",
                               "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.net_step.restype = None
        _lib.net_step.argtypes = [C.POINTER(NetGraph), C.POINTER(NetCfg), C.POINTER(NetState), C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.net_observe.restype = None
        _lib.net_observe.argtypes = [C.POINTER(NetGraph), C.POINTER(NetCfg), C.POINTER(NetState), C.c_void_p]
        _lib.net_obs_size.restype = C.c_int
        _lib.net_obs_size.argtypes = [C.POINTER(NetCfg)]
        _lib.net_closest_lane.restype = C.c_int
        _lib.net_closest_lane.argtypes = [C.POINTER(NetGraph), C.c_double, C.c_double, C.c_double]
        for name in ("net_lane_local", "net_lane_position"):
            getattr(_lib, name).restype = None
            getattr(_lib, name).argtypes = [C.POINTER(NetLane), C.c_double, C.c_double,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.net_lane_heading_at.restype = C.c_double
        _lib.net_lane_heading_at.argtypes = [C.POINTER(NetLane), C.c_double]
    return _lib


def graph_from_arrays(d: dict) -> NetGraph:
    """d: the `net_*` arrays of a golden fixture (ref_harness.dump_network) or of the product's
    lane table export."""
    g = NetGraph()
    n = len(d["net_type"])
    g.n_lanes, g.n_nodes = n, len(d["net_succ_count"])
    for k in range(n):
        for f in _LANE_I:
            setattr(g.lanes[k], f, int(d["net_" + f][k]))
        for f in _LANE_F:
            setattr(g.lanes[k], f, float(d["net_" + f][k]))
    for node in range(g.n_nodes):
        g.succ_count[node] = int(d["net_succ_count"][node])
        for j in range(NET_MAX_SUCC):
            g.succ[node][j] = int(d["net_succ"][node][j]) if j < d["net_succ"].shape[1] else -1
    return g


def cfg_from_dict(config: dict, n_vehicles: int = 5) -> NetCfg:
    """roundabout-v0 config dict (envs/roundabout_env.py:13-42 over abstract.py:102-125)."""
    c = NetCfg()
    obs, act = config["observation"], config["action"]
    c.n_vehicles = n_vehicles
    c.simulation_frequency = int(config["simulation_frequency"])
    c.policy_frequency = int(config["policy_frequency"])
    ts = act.get("target_speeds")
    ts = list(np.linspace(20, 30, 3)) if ts is None else [float(t) for t in ts]
    c.n_target_speeds = len(ts)
    for i, t in enumerate(ts):
        c.target_speeds[i] = t
    if obs["type"] == "TimeToCollision":
        c.obs_type = OBS_TTC
        c.ttc_horizon = int(obs.get("horizon", 10))
        c.obs_vehicles_count = 5
    else:
        assert obs["type"] == "Kinematics"
        c.obs_type = OBS_KINEMATICS
        c.obs_vehicles_count = int(obs.get("vehicles_count", 5))
        c.obs_see_behind = int(bool(obs.get("see_behind", False)))
        c.obs_absolute = int(bool(obs.get("absolute", False)))
        c.obs_normalize = int(bool(obs.get("normalize", True)))
        c.obs_clip = int(bool(obs.get("clip", True)))
        fr = obs.get("features_range")
        assert fr is not None, "roundabout gives explicit features_range"
        (c.obs_x_lo, c.obs_x_hi), (c.obs_y_lo, c.obs_y_hi) = map(lambda r: map(float, r), (fr["x"], fr["y"]))
        (c.obs_vx_lo, c.obs_vx_hi), (c.obs_vy_lo, c.obs_vy_hi) = map(lambda r: map(float, r), (fr["vx"], fr["vy"]))
    c.normalize_reward = int(bool(config["normalize_reward"]))
    c.duration = float(config["duration"])
    c.collision_reward = float(config["collision_reward"])
    c.high_speed_reward = float(config["high_speed_reward"])
    c.lane_change_reward = float(config["lane_change_reward"])
    c.acc_max, c.comfort_acc_max, c.comfort_acc_min = 6.0, 3.0, -5.0
    c.distance_wanted, c.time_wanted = 10.0, 1.5
    c.politeness, c.lane_change_min_acc_gain = 0.0, 0.2
    c.lane_change_max_braking_imposed, c.lane_change_delay = 2.0, 1.0
    c.perception_distance = 200.0
    return c


class NetOracleBatch:
    """n roundabout envs stepped one by one through the C oracle."""

    def __init__(self, graph: NetGraph, cfg: NetCfg, n_envs: int):
        self.g, self.cfg, self.n, self.V = graph, cfg, int(n_envs), int(cfg.n_vehicles)
        n, V = self.n, self.V
        self.a = {k: np.zeros((n, V), dtype=np.float64) for k in _SF}
        self.a.update({k: np.zeros((n, V), dtype=np.int32) for k in _SI})
        self.a["route"] = np.zeros((n, V, NET_MAX_ROUTE), dtype=np.int32)
        self.a["route_len"] = np.zeros((n, V), dtype=np.int32)
        self.a["speed_index"] = np.zeros(n, dtype=np.int32)
        self.a["time"] = np.zeros(n, dtype=np.float64)
        self.obs_size = lib().net_obs_size(C.byref(cfg))
        self.obs = np.zeros((n, self.obs_size), dtype=np.float32)
        self.reward = np.zeros(n, dtype=np.float64)
        self.terminated = np.zeros(n, dtype=np.int32)
        self.truncated = np.zeros(n, dtype=np.int32)

    def _state(self, e: int) -> NetState:
        st = NetState()
        for k in list(_SF) + list(_SI) + ["route", "route_len"]:
            setattr(st, k, self.a[k][e].ctypes.data)
        st.speed_index = self.a["speed_index"][e:e + 1].ctypes.data
        st.time = self.a["time"][e:e + 1].ctypes.data
        return st

    def observe(self):
        for e in range(self.n):
            st = self._state(e)
            lib().net_observe(C.byref(self.g), C.byref(self.cfg), C.byref(st), self.obs[e].ctypes.data)
        return self.obs

    def step(self, actions):
        for e in range(self.n):
            st = self._state(e)
            lib().net_step(C.byref(self.g), C.byref(self.cfg), C.byref(st), int(actions[e]),
                           self.obs[e].ctypes.data, self.reward[e:e + 1].ctypes.data,
                           self.terminated[e:e + 1].ctypes.data, self.truncated[e:e + 1].ctypes.data)
        return self.obs, self.reward, self.terminated, self.truncated

    def load_state(self, e: int, st: dict):
        for k in ("x", "y", "heading", "speed"):
            self.a[k][e] = st[k]
        self.a["target_speed"][e] = np.nan_to_num(st["target_speed"])
        self.a["timer"][e] = np.nan_to_num(st["timer"])
        self.a["delta"][e] = np.nan_to_num(st["delta"], nan=4.0)
        self.a["lane"][e], self.a["target_lane"][e] = st["lane"], st["target_lane"]
        self.a["crashed"][e] = st["crashed"]
        has = ~np.isnan(st["impact"][:, 0])
        self.a["has_impact"][e] = has
        self.a["impact_x"][e] = np.where(has, st["impact"][:, 0], 0.0)
        self.a["impact_y"][e] = np.where(has, st["impact"][:, 1], 0.0)
        self.a["check_collisions"][e] = st["check_collisions"]
        kind = np.full(self.V, KIND_IDM, dtype=np.int32)
        kind[0] = KIND_MDP
        self.a["kind"][e] = kind
        self.a["route"][e], self.a["route_len"][e] = st["route"], st["route_len"]
        self.a["speed_index"][e] = st["speed_index"][0]
        self.a["time"][e] = float(st["time"])
