"""The oracle's observation plugins (net_observe_grid / net_observe_ttc_from / net_observe_lidar, oracle/net_oracle.c)
pinned to the unmodified reference: OccupancyGrid with every constructor option, TimeToCollision and
LidarObservation on highway-v0, intersection-v0, roundabout-v0 and merge-v0 states (the reference's
observation_factory is orthogonal to the env, envs/common/observation.py:772-794)."""
import numpy as np
import pytest

from obs_plugin_utils import FIXTURES, load, oracle_batch, oracle_observe


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_observation_plugins_match_reference(name):
    g = load(name)
    ob, ego = oracle_batch(g)
    for k, cfg in enumerate(g["obs_cfgs"]):
        want = np.asarray(g[f"obs_{k}"], dtype=np.float64)
        got = oracle_observe(ob, ego, cfg).astype(np.float64).reshape(want.shape)
        d = np.abs(got - want)
        assert np.isfinite(got).all(), (name, cfg["type"])
        worst = float(d.max())
        # float32 outputs of identical fp64 arithmetic; libm differences stay below one float32 ulp of the ranges
        assert worst <= 2e-6 * max(1.0, float(np.abs(want).max())), (name, k, cfg, worst, np.argwhere(d == d.max())[:3])
