"""GPU parity of u-turn-v0 (circular U-turn, routed traffic, TimeToCollision horizon 16) through the C ABI against
golden rollouts of the unmodified reference and against the network oracle."""
import numpy as np
import pytest

import net_oracle as no
from parity_utils import compare_state, load_golden
from test_gpu_merge_parity import env_state, to_sd
from test_net_oracle_golden import check_routes, u_turn_state

pytestmark = pytest.mark.gpu
NAME = "u_turn_ttc"


def make_env(cfg, n, **kw):
    import highwayenv_b200 as hb

    return hb.make(cfg["_env_id"], num_envs=n, config={k: v for k, v in cfg.items() if not k.startswith("_")}, **kw)


def sd_of(g, i, t):
    st = u_turn_state(g, i, t)
    st["kind"] = np.array([1] + [0] * (len(st["x"]) - 1), dtype=np.int32)
    return st


def test_reset_matches_reference():
    g = load_golden(NAME)
    S = len(g["seeds"])
    env = make_env(g["config"], S)
    obs, _ = env.reset(seed=[int(s) for s in g["seeds"]])
    sd = env.state_dict()
    for i in range(S):
        st = u_turn_state(g, i, 0)
        # positions on the circular lanes go through CUDA's vs numpy's sin/cos: a few ulp
        assert compare_state(st, env_state(sd, i), tol=1e-12, ctx=f"reset#{i}") <= 1e-12
        check_routes(st, env_state(sd, i), f"reset#{i}")
        assert np.array_equal(sd["delta"][i][1:], st["delta"][1:])  # incl. the one randomize_behavior() draw
    assert np.max(np.abs(obs.cpu().numpy() - g["obs"][:, 0])) <= 1e-6


@pytest.mark.parametrize("NAME", ["u_turn_ttc", "u_turn_v1_ttc"])
def test_teacher_forced_vs_reference(NAME):
    g = load_golden(NAME)
    S, T = g["actions"].shape[:2]
    env = make_env(g["config"], S, autoreset_mode="Disabled")
    env.reset(seed=0)
    worst = 0.0
    for t in range(T):
        env.load_state_dict(to_sd([sd_of(g, i, t) for i in range(S)]))
        obs, rew, term, trunc, _ = env.step(g["actions"][:, t].astype(np.int32))
        sd = env.state_dict()
        obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        for i in range(S):
            ctx = f"{NAME} seed#{i} t={t}"
            st1 = u_turn_state(g, i, t + 1)
            worst = max(worst, compare_state(st1, env_state(sd, i), tol=1e-5, ctx=ctx))
            check_routes(st1, env_state(sd, i), ctx)
            assert abs(rew[i] - g["reward"][i, t]) <= 1e-9, ctx
            assert bool(term[i]) == bool(g["terminated"][i, t]) and bool(trunc[i]) == bool(g["truncated"][i, t]), ctx
            assert np.max(np.abs(obs[i] - g["obs"][i, t + 1])) <= 1e-6, ctx
    assert worst < 1e-5, worst  # slow traffic (3.5 +- 2 m/s) sits in the ill-conditioned steering regime


def test_teacher_forced_vs_oracle_many_envs_and_autoreset():
    g = load_golden(NAME)
    n, V = 192, 7
    ob = no.NetOracleBatch(no.graph_from_arrays(g), no.cfg_from_dict(g["config"], n_vehicles=V), n)
    env = make_env(g["config"], n, autoreset_mode="Disabled")
    env.reset(seed=8100)
    sd0 = env.state_dict()
    for k in ob.a:
        if k in sd0:
            ob.a[k][...] = sd0[k]
    rng = np.random.default_rng(5)
    for t in range(10):
        env.load_state_dict({k: ob.a[k].copy() for k in ob.a})
        act = rng.integers(0, 5, size=n).astype(np.int32)
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, _ = env.step(act)
        sd = env.state_dict()
        for k in ("x", "y", "heading", "speed", "timer", "target_speed"):
            assert np.max(np.abs(sd[k] - ob.a[k])) <= 1e-5, (t, k)
        for k in ("lane", "target_lane", "crashed", "has_impact", "route_len"):
            assert np.array_equal(sd[k].astype(np.int32), ob.a[k].astype(np.int32)), (t, k)
        assert np.array_equal(sd["speed_index"], ob.a["speed_index"])
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool))
        assert np.array_equal(trunc.cpu().numpy(), o_trunc.astype(bool))
        assert np.max(np.abs(obs.cpu().numpy().reshape(n, -1) - o_obs)) <= 1e-6
    env = make_env(g["config"], 64)  # SameStep autoreset on the device: every env truncates at 10 s
    env.reset(seed=9)
    resets = 0
    for t in range(22):
        obs, rew, term, trunc, info = env.step(rng.integers(0, 5, size=64).astype(np.int32))
        done = (term | trunc).cpu().numpy()
        resets += int(done.sum())
        sd = env.state_dict()
        assert np.all(sd["time"][done] == 0) and np.all(sd["x"][done, 0] == 0.0) and np.all(sd["speed"][done, 0] == 16.0)
    assert resets >= 128
