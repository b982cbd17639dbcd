"""Pins the general-network C oracle (oracle/net_oracle.c) to golden roundabout-v0 rollouts of the
unmodified reference (teacher-forced per step + free-running prefix).  CPU-only."""
import numpy as np
import pytest

import net_oracle as no
from parity_utils import compare_state, golden_state, load_golden, well_conditioned

CASES = ["roundabout_kin", "roundabout_ttc", "roundabout_v1_kin"]


def golden_net_state(g, i, t):
    st = golden_state(g, i, t)
    st["route"], st["route_len"] = g["route"][i, t], g["route_len"][i, t]
    return st


def got_state(ob, e):
    got = {k: ob.a[k][e] for k in ob.a if k not in ("speed_index", "time")}
    got["speed_index"] = ob.a["speed_index"][e]
    return got


def check_routes(st, got, ctx):
    assert np.array_equal(st["route_len"], got["route_len"]), f"{ctx} route_len"
    for v in range(len(st["route_len"])):
        n = st["route_len"][v]
        assert np.array_equal(st["route"][v][:n], got["route"][v][:n]), f"{ctx} route of vehicle {v}"


@pytest.mark.parametrize("name", CASES)
def test_teacher_forced(name):
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    cfg = no.cfg_from_dict(g["config"])
    S, T = g["actions"].shape[:2]
    ob = no.NetOracleBatch(graph, cfg, S)
    worst = 0.0
    for i in range(S):  # observation of the reset state
        ob.load_state(i, golden_net_state(g, i, 0))
    obs0 = ob.observe().reshape(g["obs"][:, 0].shape)
    assert np.max(np.abs(obs0 - g["obs"][:, 0])) <= 1e-6
    for t in range(T):
        for i in range(S):
            ob.load_state(i, golden_net_state(g, i, t))
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            ctx = f"{name} seed#{i} t={t}"
            st1 = golden_net_state(g, i, t + 1)
            worst = max(worst, compare_state(st1, got_state(ob, i), ctx=ctx))
            check_routes(st1, got_state(ob, i), ctx)
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-9


@pytest.mark.parametrize("name", CASES)
def test_free_running_prefix(name):
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    cfg = no.cfg_from_dict(g["config"])
    S, T = g["actions"].shape[:2]
    ob = no.NetOracleBatch(graph, cfg, S)
    for i in range(S):
        ob.load_state(i, golden_net_state(g, i, 0))
    alive = np.ones(S, dtype=bool)
    compared = 0
    for t in range(T):
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            st = golden_net_state(g, i, t + 1)
            alive[i] &= well_conditioned(st)
            if not alive[i]:
                continue
            compare_state(st, got_state(ob, i), ctx=f"{name} seed#{i} t={t}")
            check_routes(st, got_state(ob, i), f"{name} seed#{i} t={t}")
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6
            compared += 1
    assert compared >= 3 * S


# ------------------------------------------------------------------ intersection-v0
# intersection_v1 / intersection_continuous: ContinuousAction ego (BicycleVehicle / plain Vehicle, action.py:73-162,
# vehicle/dynamics.py:33-160) under RegulatedRoad, whose conflict prediction forward-simulates such an ego
INTER = ["intersection_kin", "intersection_grid", "intersection_v2_kin", "intersection_v1", "intersection_continuous"]
_IKEYS = ("x", "y", "heading", "speed", "target_speed", "timer", "delta", "lane", "target_lane", "crashed",
          "impact", "check_collisions", "speed_index", "time", "route", "route_len", "kind", "is_yielding",
          "count", "road_steps")


def inter_state(g, i, t):
    st = {k: g[k][i, t] for k in _IKEYS}
    for k in ("lat_speed", "yaw_rate"):  # BicycleVehicle state (fixtures of dynamical egos only)
        if k in g:
            st[k] = g[k][i, t]
    return st


def compare_inter(st, a, e, ctx, tol=1e-9):
    """a: dict of arrays [n_envs, 32, ...] in the oracle's schema."""
    n = int(st["count"])
    assert n == int(a["count"][e]), f"{ctx} count {n} vs {a['count'][e]}"
    for k in ("x", "y", "heading", "speed", "target_speed"):
        ref = st[k][:n]
        has = ~np.isnan(ref)  # a plain Vehicle (ContinuousAction ego) has no target_speed
        d = np.max(np.abs(ref[has] - a[k][e][:n][has]), initial=0.0)
        assert d <= tol, f"{ctx} {k} {d}"
    idm = st["kind"][:n] == 0
    if idm.any():
        assert np.max(np.abs(st["timer"][:n][idm] - a["timer"][e][:n][idm])) <= tol, f"{ctx} timer"
        assert np.max(np.abs(st["delta"][:n][idm] - a["delta"][e][:n][idm])) <= tol, f"{ctx} delta"
    ref = dict(st)
    ref["target_lane"] = np.where(st["target_lane"] < 0, st["lane"], st["target_lane"])  # plain Vehicle: mirrors lane
    for k in ("lane", "target_lane", "kind", "route_len"):
        assert np.array_equal(ref[k][:n], np.asarray(a[k][e][:n]).astype(st[k].dtype)), f"{ctx} {k}"
    assert np.array_equal(st["crashed"][:n].astype(bool), np.asarray(a["crashed"][e][:n]).astype(bool)), f"{ctx} crashed"
    assert np.array_equal(st["is_yielding"][:n].astype(bool), np.asarray(a["is_yielding"][e][:n]).astype(bool)), f"{ctx} yield"
    has = ~np.isnan(st["impact"][:n, 0])
    assert np.array_equal(has, np.asarray(a["has_impact"][e][:n]).astype(bool)), f"{ctx} has_impact"
    for v in range(n):
        m = st["route_len"][v]
        assert np.array_equal(st["route"][v][:m], a["route"][e][v][:m]), f"{ctx} route {v}"
    assert int(st["road_steps"]) == int(a["road_steps"][e]), f"{ctx} road_steps"


@pytest.mark.parametrize("name", INTER)
def test_intersection_reset_and_teacher_forced(name):
    """_make_vehicles (9 spawns, 45 warm-up substeps under RegulatedRoad, challenger, ego, pruning) and
    every step incl. regulation, _clear_vehicles/_spawn_vehicle and the numpy stream, vs the reference."""
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    cfg = no.cfg_from_dict(g["config"])
    S, T = g["actions"].shape[:2]
    ob = no.IntersectionOracle(graph, cfg, S, g, g["config"])
    for i in range(S):
        ob.reset_env(i, seed=int(g["seeds"][i]))
        compare_inter(inter_state(g, i, 0), ob.a, i, f"{name} reset#{i}")
        assert np.array_equal(ob.rng_words(i), g["rng_words"][i, 0])
    obs0 = ob.observe().reshape(g["obs"][:, 0].shape)
    assert np.max(np.abs(obs0 - g["obs"][:, 0])) <= 1e-6
    for t in range(T):
        for i in range(S):
            ob.load_state(i, inter_state(g, i, t))
            ob.set_rng_words(i, g["rng_words"][i, t])
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            ctx = f"{name} #{i} t={t}"
            st1 = inter_state(g, i, t + 1)
            compare_inter(st1, ob.a, i, ctx, tol=1e-7)  # 32 seeds: a few steps pass through a crawling vehicle
            if "lat_speed" in st1:
                n = int(st1["count"])
                for k in ("lat_speed", "yaw_rate"):
                    assert np.max(np.abs(np.nan_to_num(st1[k][:n]) - ob.a[k][i][:n])) <= 1e-7, (ctx, k)
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6, ctx
            assert np.array_equal(ob.rng_words(i), g["rng_words"][i, t + 1]), ctx


def test_intersection_multi_agent_reset_and_teacher_forced():
    """intersection-multi-agent-v0: two controlled vehicles (MultiAgentAction tuple actions, stacked
    observations, mean reward, any-crashed / all-arrived termination, per-agent info) vs the reference."""
    name = "intersection_multi_agent"
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    cfg = no.cfg_from_dict(g["config"])
    S, T = g["actions"].shape[:2]
    ob = no.IntersectionOracle(graph, cfg, S, g, g["config"])
    assert ob.A == 2
    for i in range(S):
        ob.reset_env(i, seed=int(g["seeds"][i]))
        compare_inter(inter_state(g, i, 0), ob.a, i, f"{name} reset#{i}")
        assert np.array_equal(ob.rng_words(i), g["rng_words"][i, 0])
    obs0 = ob.observe().reshape(g["obs"][:, 0].shape)
    assert np.max(np.abs(obs0 - g["obs"][:, 0])) <= 1e-6
    for t in range(T):
        for i in range(S):
            ob.load_state(i, inter_state(g, i, t))
            ob.set_rng_words(i, g["rng_words"][i, t])
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            ctx = f"{name} #{i} t={t}"
            # 32 seeds: a few one-step comparisons pass through a crawling vehicle (1 / not_zero(speed) in the steering
            # law amplifies the last-ulp libm difference to a few 1e-9 m); the north_star tolerance is 1e-5
            compare_inter(inter_state(g, i, t + 1), ob.a, i, ctx, tol=1e-7)
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6, ctx
            assert np.max(np.abs(ob.agents_reward[i] - g["agents_rewards"][i, t])) <= 1e-9, ctx
            assert np.array_equal(ob.agents_terminated[i].astype(bool), g["agents_terminated"][i, t]), ctx
            assert np.array_equal(ob.rng_words(i), g["rng_words"][i, t + 1]), ctx


# ------------------------------------------------------------------ merge-v0 (road objects, sine ramp)
MERGE = ["merge_kin", "merge_v1_kin", "merge_obstacle_hit"]


def merge_state(g, i, t):
    """golden state with the Obstacle in the slot after the vehicles (kind 3)"""
    st = golden_state(g, i, t)
    V = len(st["x"])
    st["target_lane"] = np.where(st["target_lane"] < 0, st["lane"], st["target_lane"])
    st["route"], st["route_len"] = np.zeros((V, no.NET_MAX_ROUTE), dtype=np.int32), np.zeros(V, dtype=np.int32)
    st["kind"], st["count"] = g["kind"][i, t], V
    # RoadObject.__init__ gives objects impact = zeros (objects.py:66); it is inert (objects never step)
    st["impact"] = np.where((st["kind"] == 3)[:, None], np.nan, st["impact"])
    st["is_yielding"], st["road_steps"] = np.zeros(V, dtype=np.int32), 0
    return st


@pytest.mark.parametrize("name", MERGE)
def test_merge_teacher_forced(name):
    """merge-v0 / merge-v1: IDM traffic + a merging vehicle on the sine ramp, an Obstacle at the ramp's end
    (neighbour search, IDM, collisions with the full impact, Kinematics rows), merge reward, x > 370 termination."""
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    V = g["x"].shape[2]
    assert V == 6 and list(g["kind"][0, 0]) == [1, 0, 0, 0, 0, 3]
    cfg = no.cfg_from_dict(g["config"], n_vehicles=V)
    S, T = g["actions"].shape[:2]
    ob = no.NetOracleBatch(graph, cfg, S)
    for i in range(S):
        ob.load_state(i, merge_state(g, i, 0))
    obs0 = ob.observe().reshape(g["obs"][:, 0].shape)
    assert np.max(np.abs(obs0 - g["obs"][:, 0])) <= 1e-6
    worst, crashes = 0.0, 0
    for t in range(T):
        for i in range(S):
            ob.load_state(i, merge_state(g, i, t))
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            ctx = f"{name} seed#{i} t={t}"
            st1 = merge_state(g, i, t + 1)
            worst = max(worst, compare_state(st1, got_state(ob, i), ctx=ctx))
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and not trunc[i] and not g["truncated"][i, t], ctx
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6, ctx
            crashes += int(st1["crashed"].any())
    assert worst < 1e-9


# ------------------------------------------------------------------ two-way-v0 (oncoming lane, no lane changes)
def two_way_state(g, i, t):
    st = golden_state(g, i, t)
    V = len(st["x"])
    st["target_lane"] = np.where(st["target_lane"] < 0, st["lane"], st["target_lane"])
    st["route"], st["route_len"] = np.zeros((V, no.NET_MAX_ROUTE), dtype=np.int32), np.zeros(V, dtype=np.int32)
    st["kind"] = np.array([1] + [0] * (V - 1), dtype=np.int32)
    st["count"], st["is_yielding"], st["road_steps"] = V, np.zeros(V, dtype=np.int32), 0
    st["no_lane_change"] = g["no_lane_change"][i, t]
    return st


def test_two_way_teacher_forced():
    """two-way-v0: IDMVehicle(enable_lane_change=False) traffic incl. two oncoming vehicles on ("b","a",0),
    TimeToCollision horizon 5, reward = 0.8 speed_index / 2 + 0.2 (1 - target lane id), never truncated"""
    name = "two_way_ttc"
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    V = g["x"].shape[2]
    cfg = no.cfg_from_dict(g["config"], n_vehicles=V)
    assert cfg.reward_type == 3 and cfg.obs_type == no.OBS_TTC
    S, T = g["actions"].shape[:2]
    ob = no.NetOracleBatch(graph, cfg, S)
    for i in range(S):
        ob.load_state(i, two_way_state(g, i, 0))
    obs0 = ob.observe().reshape(g["obs"][:, 0].shape)
    assert np.max(np.abs(obs0 - g["obs"][:, 0])) <= 1e-6
    worst = 0.0
    for t in range(T):
        for i in range(S):
            ob.load_state(i, two_way_state(g, i, t))
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            ctx = f"{name} seed#{i} t={t}"
            worst = max(worst, compare_state(two_way_state(g, i, t + 1), got_state(ob, i), ctx=ctx))
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and not trunc[i] and not g["truncated"][i, t], ctx
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-9


# ------------------------------------------------------------------ u-turn-v0 (circular U-turn, routed traffic)
def u_turn_state(g, i, t):
    st = golden_net_state(g, i, t)
    st["target_lane"] = np.where(st["target_lane"] < 0, st["lane"], st["target_lane"])
    return st


@pytest.mark.parametrize("name", ["u_turn_ttc", "u_turn_v1_ttc"])
def test_u_turn_teacher_forced(name):
    """u-turn-v0 / v1 (connected-lane neighbour search): routes to "d" through the circular lanes, ego PURSUIT_TAU = TAU_HEADING, TimeToCollision with a
    16 s horizon, left-lane / speed reward with on_road factor, truncation at 10 s"""
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    V = g["x"].shape[2]
    cfg = no.cfg_from_dict(g["config"], n_vehicles=V)
    assert cfg.reward_type == 4 and cfg.ttc_horizon == 16 and V == 7
    S, T = g["actions"].shape[:2]
    ob = no.NetOracleBatch(graph, cfg, S)
    for i in range(S):
        ob.load_state(i, u_turn_state(g, i, 0))
    obs0 = ob.observe().reshape(g["obs"][:, 0].shape)
    assert np.max(np.abs(obs0 - g["obs"][:, 0])) <= 1e-6
    worst = 0.0
    for t in range(T):
        for i in range(S):
            ob.load_state(i, u_turn_state(g, i, t))
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            ctx = f"{name} seed#{i} t={t}"
            st1 = u_turn_state(g, i, t + 1)
            worst = max(worst, compare_state(st1, got_state(ob, i), ctx=ctx))
            check_routes(st1, got_state(ob, i), ctx)
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-9
