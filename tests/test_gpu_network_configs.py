"""Configuration coverage of the network kernels beyond the defaults pinned by the golden fixtures: every variant is
checked teacher-forced against the oracle, and the device resets against the numpy-exact host reset / the oracle's
reset (same draws word for word)."""
import numpy as np
import pytest

import net_oracle as no
from parity_utils import load_golden

pytestmark = pytest.mark.gpu
V32 = 32


def _cfg(name, over):
    g = load_golden(name)
    cfg = {k: v for k, v in g["config"].items()}
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(cfg.get(k), dict):
            cfg[k] = {**cfg[k], **v}
        else:
            cfg[k] = v
    return g, cfg


def _make(cfg, n, **kw):
    import highwayenv_b200 as hb

    return hb.make(cfg["_env_id"], num_envs=n, config={k: v for k, v in cfg.items() if not k.startswith("_")}, **kw)


@pytest.mark.parametrize("over", [
    {"observation": {"vehicles_count": 3}},
    {"observation": {"vehicles_count": 8, "see_behind": True}},
    {"observation": {"normalize": False}},
    {"observation": {"clip": False, "absolute": False}},
    {"observation": {"type": "TimeToCollision", "horizon": 5}},
    {"observation": {"type": "TimeToCollision", "horizon": 12}},
    {"simulation_frequency": 10, "policy_frequency": 2},
    {"duration": 4, "normalize_reward": False, "collision_reward": -3, "high_speed_reward": 0.7},
    {"action": {"target_speeds": [0, 5, 10]}},
    {"incoming_vehicle_destination": 1},
], ids=lambda o: "-".join(f"{k}" for k in o))
def test_roundabout_options_vs_oracle(over):
    g, cfg = _cfg("roundabout_kin", over)
    if cfg["observation"].get("type") == "TimeToCollision":
        cfg["observation"] = {"type": "TimeToCollision", "horizon": cfg["observation"]["horizon"]}
    n = 96
    ob = no.NetOracleBatch(no.graph_from_arrays(g), no.cfg_from_dict(cfg), n)
    env = _make(cfg, n, autoreset_mode="Disabled")
    env.reset(seed=2300)
    sd0 = env.state_dict()
    for k in ob.a:
        if k in sd0:
            ob.a[k][...] = sd0[k]
    rng = np.random.default_rng(1)
    for t in range(8):
        env.load_state_dict({k: ob.a[k].copy() for k in ob.a})
        act = rng.integers(0, 5, size=n).astype(np.int32)
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, _ = env.step(act)
        sd = env.state_dict()
        for k in ("x", "y", "heading", "speed", "timer", "target_speed"):
            assert np.max(np.abs(sd[k] - ob.a[k])) <= 1e-6, (t, k)
        for k in ("lane", "target_lane", "crashed", "route_len"):
            assert np.array_equal(sd[k].astype(np.int32), ob.a[k].astype(np.int32)), (t, k)
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool)) and np.array_equal(trunc.cpu().numpy(), o_trunc.astype(bool))
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - o_obs)) <= 1e-6


@pytest.mark.parametrize("over", [
    {"destination": None},                      # "o" + str(np_random.integers(1, 4)) for the controlled vehicle
    {"initial_vehicle_count": 6},
    {"initial_vehicle_count": 14, "spawn_probability": 0.9},
    {"spawn_probability": 0.0},
    {"destination": "o3", "observation": {"vehicles_count": 6, "see_behind": True}},
    {"offroad_terminal": True, "normalize_reward": True, "arrived_reward": 2, "collision_reward": -4},
    {"action": {"type": "DiscreteMetaAction", "longitudinal": True, "lateral": False, "target_speeds": [0, 3, 6]}},
], ids=lambda o: "-".join(f"{k}" for k in o))
def test_intersection_options_reset_and_steps_vs_oracle(over):
    g, cfg = _cfg("intersection_kin", over)
    n = 64
    ob = no.IntersectionOracle(no.graph_from_arrays(g), no.cfg_from_dict(cfg), n, g, cfg)
    for e in range(n):
        ob.reset_env(e, seed=3400 + e)
    for mode in ("device", "host"):
        env = _make(cfg, n, autoreset_mode="Disabled", reset_mode=mode)
        env.reset(seed=3400)
        sd = env.state_dict()
        assert np.array_equal(sd["count"], ob.a["count"]), mode
        live = np.arange(V32)[None, :] < sd["count"][:, None]
        for k in ("lane", "kind", "route_len"):
            assert np.array_equal(np.where(live, sd[k], 0), np.where(live, ob.a[k], 0)), (mode, k)
        for k in ("x", "y", "speed", "delta"):
            assert np.max(np.abs(np.where(live, sd[k] - ob.a[k], 0.0))) <= 1e-6, (mode, k)
        for e in range(n):
            assert np.array_equal(sd["rng"][:, e], ob.rng_words(e)), (mode, e)
    rng = np.random.default_rng(2)
    for t in range(6):
        state = {k: ob.a[k].copy() for k in ob.a}
        state["rng"] = np.stack([ob.rng_words(e) for e in range(n)], axis=1)
        env.load_state_dict(state)
        act = rng.integers(0, 3, size=n).astype(np.int32)
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, _ = env.step(act)
        sd = env.state_dict()
        assert np.array_equal(sd["count"], ob.a["count"]), t
        live = np.arange(V32)[None, :] < sd["count"][:, None]
        for k in ("x", "y", "heading", "speed", "target_speed"):
            assert np.max(np.abs(np.where(live, sd[k] - ob.a[k], 0.0))) <= 1e-5, (t, k)
        for k in ("lane", "crashed", "is_yielding", "route_len"):
            assert np.array_equal(np.where(live, sd[k], 0).astype(np.int32), np.where(live, ob.a[k], 0).astype(np.int32)), (t, k)
        for e in range(n):
            assert np.array_equal(sd["rng"][:, e], ob.rng_words(e)), (t, e)
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool)) and np.array_equal(trunc.cpu().numpy(), o_trunc.astype(bool))
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - o_obs.reshape(n, -1))) <= 1e-6


@pytest.mark.parametrize("over", [
    {"controlled_vehicles": 3},
    {"controlled_vehicles": 4, "destination": None},
], ids=lambda o: "-".join(f"{k}" for k in o))
def test_multi_agent_options_reset_and_steps_vs_oracle(over):
    """more controlled vehicles than the default two (one per access road), random destinations"""
    g, cfg = _cfg("intersection_multi_agent", over)
    n, A = 48, int(cfg["controlled_vehicles"])
    ob = no.IntersectionOracle(no.graph_from_arrays(g), no.cfg_from_dict(cfg), n, g, cfg)
    for e in range(n):
        ob.reset_env(e, seed=4500 + e)
    for mode in ("device", "host"):
        env = _make(cfg, n, autoreset_mode="Disabled", reset_mode=mode)
        obs, _ = env.reset(seed=4500)
        assert obs.shape == (n, A, 15, 7)
        sd = env.state_dict()
        assert np.array_equal(sd["count"], ob.a["count"]), mode
        live = np.arange(V32)[None, :] < sd["count"][:, None]
        for k in ("lane", "kind", "route_len"):
            assert np.array_equal(np.where(live, sd[k], 0), np.where(live, ob.a[k], 0)), (mode, k)
        assert np.all((np.where(live, sd["kind"], 0) == 1).sum(axis=1) == A)
        for k in ("x", "y", "speed"):
            assert np.max(np.abs(np.where(live, sd[k] - ob.a[k], 0.0))) <= 1e-6, (mode, k)
        for e in range(n):
            assert np.array_equal(sd["rng"][:, e], ob.rng_words(e)), (mode, e)
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - ob.observe().reshape(n, -1))) <= 1e-6
    rng = np.random.default_rng(3)
    for t in range(6):
        state = {k: ob.a[k].copy() for k in ob.a}
        state["rng"] = np.stack([ob.rng_words(e) for e in range(n)], axis=1)
        env.load_state_dict(state)
        act = rng.integers(0, 3, size=(n, A)).astype(np.int32)
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, info = env.step(act)
        sd = env.state_dict()
        assert np.array_equal(sd["count"], ob.a["count"]), t
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.max(np.abs(info["agents_rewards"].cpu().numpy() - ob.agents_reward)) <= 1e-9
        assert np.array_equal(info["agents_terminated"].cpu().numpy(), ob.agents_terminated.astype(bool))
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool)) and np.array_equal(trunc.cpu().numpy(), o_trunc.astype(bool))
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - o_obs.reshape(n, -1))) <= 1e-6
        for e in range(n):
            assert np.array_equal(sd["rng"][:, e], ob.rng_words(e)), (t, e)
