"""exit-v0 (envs/exit_env.py:15-210; ExitObservation, envs/common/observation.py:624-675): the oracle pinned to golden
rollouts of the unmodified reference — `_create_vehicles` (weighted `choice(p=...)`, create_random, routes to "3",
enable_lane_change=False) incl. the numpy stream, and every step teacher-forced (state, reward with the goal term,
flags, observation)."""
import numpy as np

import net_oracle as no
from parity_utils import compare_state, golden_state, load_golden
from test_net_oracle_golden import got_state


def exit_state(g, i, t):
    st = golden_state(g, i, t)
    V = len(st["x"])
    st["target_lane"] = np.where(st["target_lane"] < 0, st["lane"], st["target_lane"])
    st["route"], st["route_len"] = g["route"][i, t], g["route_len"][i, t]
    st["kind"] = np.array([1] + [0] * (V - 1), dtype=np.int32)
    st["count"], st["is_yielding"], st["road_steps"] = V, np.zeros(V, dtype=np.int32), 0
    st["no_lane_change"] = g["no_lane_change"][i, t]
    return st


def make_oracle(g, n):
    return no.ExitOracle(no.graph_from_arrays(g), no.cfg_from_dict(g["config"]), n, g["config"], g["net_node_names"])


def test_exit_reset_matches_reference():
    g = load_golden("exit_obs")
    S = len(g["seeds"])
    ob = make_oracle(g, S)
    assert ob.V == 21
    for i in range(S):
        ob.reset_env(i, seed=int(g["seeds"][i]))
        assert compare_state(exit_state(g, i, 0), got_state(ob, i), tol=0.0, ctx=f"exit reset#{i}") == 0.0
        st = exit_state(g, i, 0)
        for v in range(ob.V):
            n = st["route_len"][v]
            assert n == ob.a["route_len"][i, v] and np.array_equal(st["route"][v][:n], ob.a["route"][i, v][:n]), (i, v)
        assert np.array_equal(ob.a["no_lane_change"][i], st["no_lane_change"])
        assert np.array_equal(ob.rng_words(i), g["rng_words"][i, 0]), i
    obs0 = ob.observe().reshape(g["obs"][:, 0].shape)
    assert np.max(np.abs(obs0 - g["obs"][:, 0])) <= 1e-6


def test_exit_teacher_forced():
    g = load_golden("exit_obs")
    S, T = g["actions"].shape[:2]
    ob = make_oracle(g, S)
    worst, successes, crashes = 0.0, 0, 0
    for t in range(T):
        for i in range(S):
            ob.load_state(i, exit_state(g, i, t))
        obs, rew, term, trunc = ob.step(g["actions"][:, t])
        for i in range(S):
            ctx = f"exit seed#{i} t={t}"
            st1 = exit_state(g, i, t + 1)
            worst = max(worst, compare_state(st1, got_state(ob, i), tol=1e-7, ctx=ctx))
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i].reshape(g["obs"][i, t + 1].shape) - g["obs"][i, t + 1])) <= 1e-6, ctx
            successes += int(g["reward"][i, t] >= 0.999)
            crashes += int(st1["crashed"][0])
    assert worst < 1e-7
    assert crashes > 0  # the fixture exercises collisions; the goal term is covered by the route through lane 6 when reached
