"""Pins the general-network C oracle (oracle/net_oracle.c) to golden roundabout-v0 rollouts of the
unmodified reference (teacher-forced per step + free-running prefix).  CPU-only."""
import numpy as np
import pytest

import net_oracle as no
from parity_utils import compare_state, golden_state, load_golden, well_conditioned

CASES = ["roundabout_kin", "roundabout_ttc"]


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
