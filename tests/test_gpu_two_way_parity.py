"""GPU parity of two-way-v0 (oncoming lane, IDMVehicle(enable_lane_change=False), TimeToCollision horizon 5) through
the C ABI against golden rollouts of the unmodified reference and against the network oracle."""
import numpy as np
import pytest

import net_oracle as no
from parity_utils import compare_state, load_golden
from test_gpu_merge_parity import env_state, to_sd
from test_net_oracle_golden import two_way_state

pytestmark = pytest.mark.gpu
NAME = "two_way_ttc"


def make_env(cfg, n, **kw):
    import highwayenv_b200 as hb

    return hb.make("two-way-v0", num_envs=n, config={k: v for k, v in cfg.items() if not k.startswith("_")}, **kw)


def sd_of(states):
    sd = to_sd(states)
    sd["no_lane_change"] = np.stack([s["no_lane_change"] for s in states])
    return sd


def test_reset_matches_reference():
    g = load_golden(NAME)
    S = len(g["seeds"])
    env = make_env(g["config"], S)
    obs, _ = env.reset(seed=[int(s) for s in g["seeds"]])
    sd = env.state_dict()
    for i in range(S):
        # positions go through numpy's vs CUDA's cos/sin of pi on the oncoming lane: a few ulp
        assert compare_state(two_way_state(g, i, 0), env_state(sd, i), tol=1e-12, ctx=f"reset#{i}") <= 1e-12
        assert np.array_equal(sd["no_lane_change"][i].astype(np.int32), g["no_lane_change"][i, 0])
    assert np.max(np.abs(obs.cpu().numpy() - g["obs"][:, 0])) <= 1e-6


def test_teacher_forced_vs_reference():
    g = load_golden(NAME)
    S, T = g["actions"].shape[:2]
    env = make_env(g["config"], S, autoreset_mode="Disabled")
    env.reset(seed=0)
    worst = 0.0
    for t in range(T):
        env.load_state_dict(sd_of([two_way_state(g, i, t) for i in range(S)]))
        obs, rew, term, trunc, _ = env.step(g["actions"][:, t].astype(np.int32))
        sd = env.state_dict()
        obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        for i in range(S):
            ctx = f"{NAME} seed#{i} t={t}"
            worst = max(worst, compare_state(two_way_state(g, i, t + 1), env_state(sd, i), ctx=ctx))
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and not trunc[i], ctx
            assert np.max(np.abs(obs[i] - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-7, worst


def test_teacher_forced_vs_oracle_many_envs_and_autoreset():
    g = load_golden(NAME)
    n, V = 192, 6
    ob = no.NetOracleBatch(no.graph_from_arrays(g), no.cfg_from_dict(g["config"], n_vehicles=V), n)
    env = make_env(g["config"], n, autoreset_mode="Disabled")
    env.reset(seed=7100)
    sd0 = env.state_dict()
    for k in ob.a:
        if k in sd0:
            ob.a[k][...] = sd0[k]
    assert ob.a["no_lane_change"][:, 1:].all() and not ob.a["no_lane_change"][:, 0].any()
    rng = np.random.default_rng(4)
    for t in range(14):
        env.load_state_dict({k: ob.a[k].copy() for k in ob.a})
        act = rng.integers(0, 5, size=n).astype(np.int32)
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, _ = env.step(act)
        sd = env.state_dict()
        for k in ("x", "y", "heading", "speed", "timer", "target_speed"):
            assert np.max(np.abs(sd[k] - ob.a[k])) <= 1e-6, (t, k)
        for k in ("lane", "target_lane", "crashed", "has_impact"):
            assert np.array_equal(sd[k].astype(np.int32), ob.a[k].astype(np.int32)), (t, k)
        assert np.array_equal(sd["speed_index"], ob.a["speed_index"])
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool)) and not trunc.any()
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - o_obs)) <= 1e-6
    env = make_env(g["config"], 64)  # SameStep autoreset on the device
    env.reset(seed=9)
    resets = 0
    for t in range(30):
        obs, rew, term, trunc, info = env.step(rng.integers(0, 5, size=64).astype(np.int32))
        done = (term | trunc).cpu().numpy()
        resets += int(done.sum())
        sd = env.state_dict()
        assert np.all(sd["time"][done] == 0) and np.all(sd["x"][done, 0] == 30.0)
    assert resets >= 32
