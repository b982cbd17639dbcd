"""The observation plugin registry on the device (envs/common/observation.py, csrc/hwy_observe.cu): every
observation type on every env family, as the reference's observation_factory allows
(envs/common/observation.py:772-794).  States of reference rollouts are injected and `env.observe()` is compared with
the reference's own observation of that state (fixtures obs_plugins_*.npz) and with the oracle."""
import numpy as np
import pytest

from obs_plugin_utils import FIXTURES, env_state_dict, kinds, load, oracle_batch, oracle_observe

pytestmark = pytest.mark.gpu


def make_env(g, obs_cfg, **kw):
    import highwayenv_b200 as hb

    cfg = {k: v for k, v in g["config"].items() if not k.startswith("_")}
    cfg["observation"] = obs_cfg
    return hb.make(g["config"]["_env_id"], num_envs=g["x"].shape[0], config=cfg, autoreset_mode="Disabled", **kw)


def inject(env, g):
    S, V = g["x"].shape
    sd = env_state_dict(g)
    if g["config"]["_env_id"].startswith("highway"):
        hsd = {k: sd[k] for k in ("x", "y", "heading", "speed", "target_speed", "timer", "delta", "impact_x",
                                  "impact_y", "lane", "target_lane", "crashed", "has_impact", "check_collisions",
                                  "speed_index", "time")}
        hsd["kind"] = kinds(g)
        env.load_state_dict(hsd)
    elif g["config"]["_env_id"].startswith("intersection"):
        env.load_state_dict(sd)
    else:
        sd.pop("count"), sd.pop("road_steps"), sd.pop("is_yielding")
        env.load_state_dict(sd)


@pytest.mark.parametrize("name", FIXTURES)
def test_every_observation_type_on_every_family(name):
    g = load(name)
    ob, ego = oracle_batch(g)
    for k, obs_cfg in enumerate(g["obs_cfgs"]):
        env = make_env(g, obs_cfg)
        env.reset(seed=0)
        inject(env, g)
        got = env.observe().cpu().numpy().astype(np.float64)
        want = np.asarray(g[f"obs_{k}"], dtype=np.float64)
        assert got.shape == want.shape, (name, obs_cfg["type"], got.shape, want.shape)
        assert tuple(env.single_observation_space.shape) == want.shape[1:]
        scale = max(1.0, float(np.abs(want).max()))
        d = np.abs(got - want)
        assert float(d.max()) <= 2e-6 * scale, (name, k, obs_cfg, float(d.max()), np.argwhere(d == d.max())[:3])
        orc = oracle_observe(ob, ego, obs_cfg).astype(np.float64).reshape(want.shape)
        assert float(np.abs(got - orc).max()) <= 2e-6 * scale, (name, k, "vs oracle")


def test_plugin_observation_through_step_and_autoreset():
    """highway-v0 + OccupancyGrid / TimeToCollision through env.step with SameStep autoreset: obs of running envs is the
    plugin's view of the new state, obs of finished envs the view of the freshly reset state, final_obs the view before
    the reset (checked against a twin env without autoreset and against observe() after the fact)."""
    import highwayenv_b200 as hb

    for obs_cfg in ({"type": "OccupancyGrid", "grid_size": [[-60, 60], [-10, 10]], "grid_step": [4, 2]},
                    {"type": "TimeToCollision", "horizon": 8}, {"type": "LidarObservation", "cells": 20}):
        cfg = {"vehicles_count": 20, "duration": 6, "observation": obs_cfg}
        a = hb.make("highway-fast-v0", num_envs=64, config=cfg)
        b = hb.make("highway-fast-v0", num_envs=64, config=cfg, autoreset_mode="Disabled")
        a.reset(seed=3)
        b.reset(seed=3)
        rng = np.random.default_rng(0)
        ever_done = np.zeros(64, dtype=bool)
        for t in range(8):
            act = rng.integers(0, 5, size=64).astype(np.int32)
            oa, _, ta, tra, info = a.step(act)
            ob_, _, tb, trb, _ = b.step(act)
            done = (ta | tra).cpu().numpy()
            fresh = ~ever_done
            assert np.array_equal(info["final_obs"].cpu().numpy()[fresh], ob_.cpu().numpy()[fresh]), (obs_cfg, t)
            assert np.array_equal(oa.cpu().numpy()[fresh & ~done], ob_.cpu().numpy()[fresh & ~done])
            now = a.observe().cpu().numpy()  # the plugin's view of the state as it is now (post-reset for done envs)
            assert np.array_equal(oa.cpu().numpy(), now), (obs_cfg, t)
            ever_done |= done
        assert ever_done.any()
