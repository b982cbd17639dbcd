"""GPU parity of the general-network CUDA path (roundabout-v0) through the C ABI against golden
rollouts of the unmodified reference and against the network oracle on more seeds."""
import numpy as np
import pytest

import net_oracle as no
from parity_utils import compare_state, golden_state, load_golden, well_conditioned

pytestmark = pytest.mark.gpu
CASES = ["roundabout_kin", "roundabout_ttc", "roundabout_v1_kin"]


def make_env(cfg, n, **kw):
    import highwayenv_b200 as hb

    cfg = dict(cfg)
    env_id = cfg.pop("_env_id")
    cfg.pop("_others_check_collisions", None)
    return hb.make(env_id, num_envs=n, config=cfg, **kw)


def gstate(g, i, t):
    st = golden_state(g, i, t)
    st["route"], st["route_len"] = g["route"][i, t], g["route_len"][i, t]
    return st


def to_sd(states):
    n, V = len(states), len(states[0]["x"])
    sd = {k: np.stack([s[k] for s in states]) for k in ("x", "y", "heading", "speed", "lane", "crashed",
                                                         "check_collisions", "route", "route_len")}
    sd["target_speed"] = np.stack([np.nan_to_num(s["target_speed"]) for s in states])
    sd["timer"] = np.stack([np.nan_to_num(s["timer"]) for s in states])
    sd["delta"] = np.stack([np.nan_to_num(s["delta"], nan=4.0) for s in states])
    sd["target_lane"] = np.stack([np.where(s["target_lane"] < 0, s["lane"], s["target_lane"]) for s in states])
    has = np.stack([~np.isnan(s["impact"][:, 0]) for s in states])
    sd["has_impact"] = has
    sd["impact_x"] = np.stack([np.nan_to_num(s["impact"][:, 0]) for s in states])
    sd["impact_y"] = np.stack([np.nan_to_num(s["impact"][:, 1]) for s in states])
    kind = np.zeros((n, V), dtype=np.int32)
    kind[:, 0] = 1
    sd["kind"] = kind
    sd["speed_index"] = np.array([s["speed_index"][0] for s in states], dtype=np.int32)
    sd["time"] = np.array([float(s["time"]) for s in states])
    return sd


def env_state(sd, e):
    return {k: sd[k][e] for k in sd}


def check_routes(st, got, ctx):
    assert np.array_equal(st["route_len"], got["route_len"]), f"{ctx} route_len"
    for v in range(len(st["route_len"])):
        n = st["route_len"][v]
        assert np.array_equal(st["route"][v][:n], got["route"][v][:n]), f"{ctx} route {v}"


@pytest.mark.parametrize("name", CASES)
def test_reset_matches_reference(name):
    g = load_golden(name)
    env = make_env(g["config"], len(g["seeds"]), reset_mode="host")
    obs, _ = env.reset(seed=[int(s) for s in g["seeds"]])
    sd = env.state_dict()
    for i in range(len(g["seeds"])):
        st = gstate(g, i, 0)
        assert compare_state(st, env_state(sd, i), tol=0.0, ctx=f"{name}#{i}") == 0.0
        check_routes(st, env_state(sd, i), f"{name}#{i}")
    assert np.max(np.abs(obs.cpu().numpy() - g["obs"][:, 0])) <= 1e-6


@pytest.mark.parametrize("name", CASES)
def test_teacher_forced_vs_reference(name):
    g = load_golden(name)
    S, T = g["actions"].shape[:2]
    env = make_env(g["config"], S, autoreset_mode="Disabled")
    env.reset(seed=0)
    worst = 0.0
    for t in range(T):
        env.load_state_dict(to_sd([gstate(g, i, t) for i in range(S)]))
        obs, rew, term, trunc, info = env.step(g["actions"][:, t].astype(np.int32))
        sd = env.state_dict()
        obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        for i in range(S):
            ctx = f"{name} seed#{i} t={t}"
            st1 = gstate(g, i, t + 1)
            worst = max(worst, compare_state(st1, env_state(sd, i), ctx=ctx))
            check_routes(st1, env_state(sd, i), ctx)
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i] - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-7, worst


@pytest.mark.parametrize("name", CASES)
def test_free_running_vs_reference(name):
    g = load_golden(name)
    S, T = g["actions"].shape[:2]
    env = make_env(g["config"], S, autoreset_mode="Disabled")
    env.reset(seed=[int(s) for s in g["seeds"]])
    alive = np.ones(S, dtype=bool)
    compared = 0
    for t in range(T):
        obs, rew, term, trunc, _ = env.step(g["actions"][:, t].astype(np.int32))
        sd = env.state_dict()
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        for i in range(S):
            st = gstate(g, i, t + 1)
            alive[i] &= well_conditioned(st)
            if not alive[i]:
                continue
            compare_state(st, env_state(sd, i), ctx=f"{name} seed#{i} t={t}")
            check_routes(st, env_state(sd, i), f"{name} seed#{i} t={t}")
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9
            assert np.max(np.abs(obs[i] - g["obs"][i, t + 1])) <= 1e-6
            compared += 1
    assert compared >= 3 * S


@pytest.mark.parametrize("name,n,T", [("roundabout_ttc", 256, 11), ("roundabout_kin", 128, 11), ("roundabout_v1_kin", 128, 11)])
def test_teacher_forced_vs_oracle_many_envs(name, n, T):
    g = load_golden(name)
    graph = no.graph_from_arrays(g)
    cfg = no.cfg_from_dict(g["config"])
    ob = no.NetOracleBatch(graph, cfg, n)
    env = make_env(g["config"], n, autoreset_mode="Disabled")
    env.reset(seed=31000)
    sd0 = env.state_dict()
    for k in ob.a:
        if k in sd0:  # count / is_yielding / road_steps are intersection-only
            ob.a[k][...] = sd0[k]
    rng = np.random.default_rng(7)
    for t in range(T):
        env.load_state_dict({k: ob.a[k].copy() for k in ob.a})
        act = rng.integers(0, 5, size=n).astype(np.int32)
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, _ = env.step(act)
        sd = env.state_dict()
        for k in ("x", "y", "heading", "speed", "timer", "target_speed"):
            assert np.max(np.abs(sd[k] - ob.a[k])) <= 1e-7, (t, k)
        for k in ("lane", "target_lane", "crashed", "has_impact", "route_len"):
            assert np.array_equal(sd[k].astype(np.int32), ob.a[k].astype(np.int32)), (t, k)
        assert np.array_equal(sd["speed_index"], ob.a["speed_index"])
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool))
        assert np.array_equal(trunc.cpu().numpy(), o_trunc.astype(bool))
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - o_obs)) <= 1e-6


def test_device_reset_matches_host_reset():
    """hwy_roundabout_reset (device PCG64 + ziggurat normal) against the numpy-exact host spawn:
    identical draws / lanes / routes, coordinates within CUDA-vs-numpy sin/cos rounding."""
    g = load_golden("roundabout_ttc")
    n = 512
    dev_env = make_env(g["config"], n, reset_mode="device")
    host_env = make_env(g["config"], n, reset_mode="host")
    o_d, _ = dev_env.reset(seed=777)
    o_h, _ = host_env.reset(seed=777)
    for rep in range(3):  # the streams keep advancing identically over repeated resets
        a, b = dev_env.state_dict(), host_env.state_dict()
        for k in ("x", "y", "heading", "timer"):
            assert np.max(np.abs(a[k] - b[k])) <= 1e-12, (rep, k)
        for k in ("speed", "delta", "target_speed"):
            assert np.array_equal(a[k], b[k]), (rep, k)  # pure RNG arithmetic: bit-exact
        for k in ("lane", "target_lane", "route", "route_len", "kind", "speed_index"):
            assert np.array_equal(a[k], b[k]), (rep, k)
        assert np.max(np.abs(o_d.cpu().numpy() - o_h.cpu().numpy())) <= 1e-6
        # generator state after the spawn equals numpy's
        words = dev_env._rng.cpu().numpy().view(np.uint64)
        for i in (0, 1, n - 1):
            st = host_env._rngs[i].bit_generator.state
            sv = st["state"]["state"]
            assert int(words[0, i]) == sv >> 64 and int(words[1, i]) == sv & ((1 << 64) - 1)
            assert int(words[4, i]) >> 32 == int(st["has_uint32"])
        o_d, _ = dev_env.reset()
        o_h, _ = host_env.reset()


def test_device_autoreset_same_step():
    g = load_golden("roundabout_kin")
    n = 64
    env = make_env(g["config"], n)  # device reset, SameStep
    env.reset(seed=9)
    rng = np.random.default_rng(0)
    resets = 0
    for t in range(25):
        obs, rew, term, trunc, info = env.step(rng.integers(0, 5, size=n).astype(np.int32))
        done = (term | trunc).cpu().numpy()
        resets += int(done.sum())
        sd = env.state_dict()
        assert np.all(sd["time"][done] == 0) and np.all(sd["x"][done, 0] == 2.0) and np.all(sd["y"][done, 0] == 45.0)
        assert np.array_equal(info["final_obs"].cpu().numpy()[~done], obs.cpu().numpy()[~done])
    assert resets >= 2 * n


def test_autoreset_host_path():
    g = load_golden("roundabout_kin")
    env = make_env(g["config"], 16, reset_mode="host")
    env.reset(seed=5)
    for t in range(14):
        obs, rew, term, trunc, info = env.step(np.full(16, 1, dtype=np.int32))
    assert np.all(env.state_dict()["time"] <= 11)
