"""Pins the C oracle (oracle/hwy_oracle.c) to golden trajectories of the unmodified
reference (tests/golden/*.npz, made by oracle/gen_golden.py).  CPU-only."""
import numpy as np
import pytest

import hwy_oracle as ho
from parity_utils import compare_state, golden_state, load_golden, well_conditioned

CASES = ["highway_fast_v20", "highway_fast_v50", "highway_v50", "highway_v100_continuous",
         "highway_discrete_action", "highway_fast_features", "highway_fast_features_range"]


def _actions(g, t):
    """DiscreteAction (action.py:165-196): index -> float32 (throttle, steering) pair"""
    a = g["actions"][:, t]
    if g["config"]["action"]["type"] == "DiscreteAction":
        return ho.discrete_action_table(g["config"]["action"].get("actions_per_axis", 3))[a]
    return a


def _got(ob, e=0):
    g = {k: ob.a[k][e] for k in ob.a if k not in ("speed_index", "time")}
    g["speed_index"] = ob.a["speed_index"][e]
    return g


@pytest.mark.parametrize("name", CASES)
def test_reset_matches_reference(name):
    g = load_golden(name)
    cfg = ho.cfg_from_dict(g["config"])
    ob = ho.OracleBatch(cfg, len(g["seeds"]), seeds=g["seeds"])
    obs = ob.reset()
    for i in range(len(g["seeds"])):
        w = compare_state(golden_state(g, i, 0), _got(ob, i), tol=0.0, ctx=f"{name} seed#{i} reset")
        assert w == 0.0
        assert np.array_equal(obs[i], g["obs"][i, 0])


@pytest.mark.parametrize("name", CASES)
def test_teacher_forced_steps(name):
    """Load the reference state at t, step once, compare with the reference at t+1."""
    g = load_golden(name)
    cfg = ho.cfg_from_dict(g["config"])
    S, T = g["actions"].shape[:2]
    ob = ho.OracleBatch(cfg, S, seeds=g["seeds"])
    worst = 0.0
    for t in range(T):
        for i in range(S):
            ob.load_state(i, golden_state(g, i, t))
        obs, rew, term, trunc = ob.step(_actions(g, t))
        for i in range(S):
            ctx = f"{name} seed#{i} t={t}"
            worst = max(worst, compare_state(golden_state(g, i, t + 1), _got(ob, i), ctx=ctx))
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]), ctx
            assert bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i] - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-9  # same libm family: expect ~1e-13


@pytest.mark.parametrize("name", CASES)
def test_free_running_prefix(name):
    g = load_golden(name)
    cfg = ho.cfg_from_dict(g["config"])
    S, T = g["actions"].shape[:2]
    ob = ho.OracleBatch(cfg, S, seeds=g["seeds"])
    ob.reset()
    alive = np.ones(S, dtype=bool)
    compared = 0
    for t in range(T):
        obs, rew, term, trunc = ob.step(_actions(g, t))
        for i in range(S):
            st = golden_state(g, i, t + 1)
            alive[i] &= well_conditioned(st)
            if not alive[i]:
                continue
            compare_state(st, _got(ob, i), ctx=f"{name} seed#{i} t={t}")
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9
            assert bool(term[i]) == bool(g["terminated"][i, t])
            assert np.max(np.abs(obs[i] - g["obs"][i, t + 1])) <= 1e-6
            compared += 1
    assert compared >= S * 3


@pytest.mark.parametrize("name", ["reset_highway_fast_v50", "reset_highway_v100"])
def test_reset_stream_continuation(name):
    """reset(seed) then reset() continues the PCG64 stream exactly (autoreset semantics)."""
    g = load_golden(name)
    base = load_golden("highway_fast_v50" if "fast" in name else "highway_v100_continuous")
    cfg = ho.cfg_from_dict(base["config"])
    n = len(g["seeds"])
    ob = ho.OracleBatch(cfg, n, seeds=g["seeds"])
    for r in range(2):
        obs = ob.reset()
        for i in range(n):
            st = {k: g[k][2 * i + r] for k in g if k not in ("obs", "seeds")}
            compare_state(st, _got(ob, i), tol=0.0, ctx=f"{name} seed#{i} reset#{r}")
            assert np.array_equal(obs[i], g["obs"][2 * i + r])
