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


# ------------------------------------------------------------------ round-2 regressions
@pytest.mark.parametrize("name", ["roundabout_kin", "intersection_kin", "merge_kin"])
def test_reset_with_config_option_keeps_the_numpy_stream(name):
    """reset(options={"config": ...}) re-allocates the device buffers; like the reference (seed=None keeps
    np_random, abstract.py:219-249) the env's generator must go on, not restart from zeroed words."""
    g, cfg = _cfg(name, {})
    n = 16
    a, b = _make(cfg, n), _make(cfg, n)
    a.reset(seed=321)
    b.reset(seed=321)
    a.reset()
    b.reset(options={"config": {"duration": cfg.get("duration", 11)}})
    wa, wb = a._rng.cpu().numpy(), b._rng.cpu().numpy()
    assert np.any(wb != 0) and np.array_equal(wa, wb)
    sa, sb = a.state_dict(), b.state_dict()
    live = np.ones(sa["x"].shape, dtype=bool)
    if "count" in sa:  # dynamic population: slots past `count` hold stale data
        assert np.array_equal(sa["count"], sb["count"])
        live = np.arange(sa["x"].shape[1])[None, :] < sa["count"][:, None]
    for k in ("x", "y", "speed", "lane"):
        assert np.array_equal(np.where(live, sa[k], 0), np.where(live, sb[k], 0)), k


def test_device_int64_actions_are_staged_on_the_device():
    """torch.randint / argmax produce int64 CUDA tensors: they must be converted on the device (no host round trip)
    and give the same step as int32 actions."""
    import torch

    g, cfg = _cfg("roundabout_kin", {})
    a, b = _make(cfg, 32, autoreset_mode="Disabled"), _make(cfg, 32, autoreset_mode="Disabled")
    a.reset(seed=5)
    b.reset(seed=5)
    act = torch.randint(0, 5, (32,), device="cuda")
    assert act.dtype == torch.int64
    oa = a.step(act)[0].clone()
    ob_ = b.step(act.to(torch.int32))[0]
    assert torch.equal(oa, ob_)


def test_intersection_full_slots_spawn_is_reported():
    """A spawn accepted by _spawn_vehicle while all 32 slots are taken cannot be stored (the reference's list is
    unbounded): the drop is counted in info["spawn_overflow"] instead of passing silently."""
    g, cfg = _cfg("intersection_kin", {"spawn_probability": 1.0})
    n = 8
    env = _make(cfg, n, autoreset_mode="Disabled")
    env.reset(seed=77)
    sd = env.state_dict()
    exit_lane = int(np.nonzero(np.asarray(g["net_exit_lane"]))[0][0])
    for e in range(n):
        ego = int(np.nonzero(sd["kind"][e, :sd["count"][e]] == 1)[0][0])
        src = 0 if ego != 0 else 1
        for v in range(V32):
            if v == ego:
                continue
            for k in sd:
                if sd[k].ndim >= 2 and sd[k].shape[:2] == (n, V32) and k != "rng":
                    sd[k][e, v] = sd[k][e, src]
            sd["kind"][e, v], sd["lane"][e, v], sd["target_lane"][e, v] = 0, exit_lane, exit_lane
            sd["route_len"][e, v] = 0
            sd["speed"][e, v] = sd["target_speed"][e, v] = 0.0
        sd["count"][e] = V32
    # park the fillers along the exit lane (before the point where _clear_vehicles removes them), far from the spawn points
    L = {k[4:]: np.asarray(g[k])[exit_lane] for k in g if k.startswith("net_") and np.ndim(g[k]) == 1 and len(g[k]) > exit_lane}
    for e in range(n):
        for v in range(V32):
            if sd["kind"][e, v] == 1:
                continue
            s = 5.0 + 2.0 * v
            sd["x"][e, v] = L["sx"] + s * L["dx"]
            sd["y"][e, v] = L["sy"] + s * L["dy"]
            sd["heading"][e, v] = L["heading"]
    env.load_state_dict(sd)
    total = 0
    for t in range(3):
        _, _, _, _, info = env.step(np.ones(n, dtype=np.int32))
        total = int(info["spawn_overflow"].sum().item())
        assert np.all(env.state_dict()["count"] <= V32)
    assert total > 0
