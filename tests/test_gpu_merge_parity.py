"""GPU parity of merge-v0 / merge-v1 (road objects: the Obstacle at the end of the ramp; sine lane; merge reward)
through the C ABI against golden rollouts of the unmodified reference and against the network oracle."""
import numpy as np
import pytest

import net_oracle as no
from parity_utils import compare_state, load_golden
from test_net_oracle_golden import merge_state

pytestmark = pytest.mark.gpu
CASES = ["merge_kin", "merge_v1_kin", "merge_obstacle_hit"]
KEYS = ("x", "y", "heading", "speed", "target_speed", "timer", "delta", "impact_x", "impact_y", "lane", "target_lane",
        "kind", "crashed", "has_impact", "check_collisions", "route", "route_len", "speed_index", "time")


def make_env(cfg, n, **kw):
    import highwayenv_b200 as hb

    cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
    return hb.make(kw.pop("env_id"), num_envs=n, config=cfg, **kw)


def to_sd(states):
    sd = {}
    for k in ("x", "y", "heading", "speed", "lane", "target_lane", "kind", "crashed", "check_collisions", "route", "route_len"):
        sd[k] = np.stack([s[k] for s in states])
    sd["target_speed"] = np.stack([np.nan_to_num(s["target_speed"]) for s in states])
    sd["timer"] = np.stack([np.nan_to_num(s["timer"]) for s in states])
    sd["delta"] = np.stack([np.nan_to_num(s["delta"], nan=4.0) for s in states])
    has = np.stack([~np.isnan(s["impact"][:, 0]) for s in states])
    sd["has_impact"] = has
    sd["impact_x"] = np.stack([np.nan_to_num(s["impact"][:, 0]) for s in states])
    sd["impact_y"] = np.stack([np.nan_to_num(s["impact"][:, 1]) for s in states])
    sd["speed_index"] = np.array([s["speed_index"][0] for s in states], dtype=np.int32)
    sd["time"] = np.array([float(s["time"]) for s in states])
    return sd


def env_state(sd, e):
    return {k: sd[k][e] for k in sd}


@pytest.mark.parametrize("name", ["merge_kin", "merge_v1_kin"])
def test_reset_matches_reference(name):
    """hwy_merge_reset: integers / uniform draws, lane positions incl. the Obstacle at the ramp's end"""
    g = load_golden(name)
    S = len(g["seeds"])
    env = make_env(g["config"], S, env_id=g["config"]["_env_id"])
    obs, _ = env.reset(seed=[int(s) for s in g["seeds"]])
    sd = env.state_dict()
    for i in range(S):
        assert compare_state(merge_state(g, i, 0), env_state(sd, i), tol=0.0, ctx=f"{name}#{i}") == 0.0
        assert np.array_equal(sd["kind"][i], g["kind"][i, 0])
    assert np.max(np.abs(obs.cpu().numpy() - g["obs"][:, 0])) <= 1e-6


@pytest.mark.parametrize("name", CASES)
def test_teacher_forced_vs_reference(name):
    g = load_golden(name)
    S, T = g["actions"].shape[:2]
    env = make_env(g["config"], S, env_id=g["config"]["_env_id"], autoreset_mode="Disabled")
    env.reset(seed=0)
    worst = 0.0
    for t in range(T):
        env.load_state_dict(to_sd([merge_state(g, i, t) for i in range(S)]))
        obs, rew, term, trunc, info = env.step(g["actions"][:, t].astype(np.int32))
        sd = env.state_dict()
        obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        for i in range(S):
            ctx = f"{name} seed#{i} t={t}"
            worst = max(worst, compare_state(merge_state(g, i, t + 1), env_state(sd, i), ctx=ctx))
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and not trunc[i], ctx
            assert np.max(np.abs(obs[i] - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-7, worst


def test_teacher_forced_vs_oracle_many_envs_and_autoreset():
    g = load_golden("merge_kin")
    n, V = 192, 6
    ob = no.NetOracleBatch(no.graph_from_arrays(g), no.cfg_from_dict(g["config"], n_vehicles=V), n)
    env = make_env(g["config"], n, env_id="merge-v0", autoreset_mode="Disabled")
    env.reset(seed=6100)
    sd0 = env.state_dict()
    for k in ob.a:
        if k in sd0:
            ob.a[k][...] = sd0[k]
    rng = np.random.default_rng(2)
    for t in range(16):
        env.load_state_dict({k: ob.a[k].copy() for k in ob.a})
        act = rng.integers(0, 5, size=n).astype(np.int32)
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, _ = env.step(act)
        sd = env.state_dict()
        for k in ("x", "y", "heading", "speed", "timer", "target_speed"):
            assert np.max(np.abs(sd[k] - ob.a[k])) <= 1e-6, (t, k)
        for k in ("lane", "target_lane", "crashed", "has_impact"):
            assert np.array_equal(sd[k].astype(np.int32), ob.a[k].astype(np.int32)), (t, k)
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool)) and not trunc.any()
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - o_obs)) <= 1e-6
    # SameStep autoreset on the device: finished envs restart at the reference's spawn, the others go on
    env = make_env(g["config"], 64, env_id="merge-v0")
    env.reset(seed=9)
    resets = 0
    for t in range(30):
        obs, rew, term, trunc, info = env.step(rng.integers(0, 5, size=64).astype(np.int32))
        done = (term | trunc).cpu().numpy()
        resets += int(done.sum())
        sd = env.state_dict()
        assert np.all(sd["time"][done] == 0) and np.all(sd["x"][done, 0] == 30.0) and np.all(sd["speed"][done, 0] == 30.0)
        assert np.all(sd["kind"][:, 5] == 3)
    assert resets >= 64
