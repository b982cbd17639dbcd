"""`env.road_substeps(n)` / `hwy_highway_substeps`: the reference's operator seam, n x (Road.act(); Road.step(dt)) with
no action mapping, observation, reward or clock (abstract.py:304-307).  Pinned through `env.step`, which the oracle and
the golden rollouts pin: a DiscreteMetaAction IDLE step and a ContinuousAction step move the vehicles exactly like the
same number of bare substeps, and substeps compose."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE = ("x", "y", "heading", "speed", "target_speed", "timer", "impact_x", "impact_y", "lane", "target_lane", "crashed",
         "has_impact")


def same_state(a, b, tag):
    sa, sb = a.state_dict(), b.state_dict()
    for k in STATE:
        assert np.array_equal(sa[k], sb[k]), (tag, k)


@pytest.mark.parametrize("env_id,over", [("highway-fast-v0", {"vehicles_count": 50}), ("highway-fast-v0", None),
                                         ("highway-v0", {"vehicles_count": 70})])
def test_idle_step_equals_bare_substeps(env_id, over):
    import highwayenv_b200 as hb

    n = 48
    a = hb.make(env_id, num_envs=n, config=over, autoreset_mode="Disabled")
    b = hb.make(env_id, num_envs=n, config=over, autoreset_mode="Disabled")
    a.reset(seed=21)
    b.reset(seed=21)
    frames = int(a.config["simulation_frequency"]) // int(a.config["policy_frequency"])
    idle = np.ones(n, dtype=np.int32)  # DiscreteMetaAction IDLE
    for t in range(12):
        a.step(idle)
        if t % 2:
            b.road_substeps(frames)
        else:  # substeps compose (impacts and lane-change timers carry over between launches)
            b.road_substeps(2)
            b.road_substeps(frames - 2)
        same_state(a, b, t)
    sa, sb = a.state_dict(), b.state_dict()
    assert np.array_equal(sa["rng"], sb["rng"]) and np.all(sb["time"] == 0.0) and np.all(sa["time"] > 0.0)


def test_continuous_step_equals_bare_substeps_with_the_action():
    import highwayenv_b200 as hb

    n, over = 32, {"vehicles_count": 40, "action": {"type": "ContinuousAction"}}
    a = hb.make("highway-v0", num_envs=n, config=over, autoreset_mode="Disabled")
    b = hb.make("highway-v0", num_envs=n, config=over, autoreset_mode="Disabled")
    a.reset(seed=3)
    b.reset(seed=3)
    frames = int(a.config["simulation_frequency"]) // int(a.config["policy_frequency"])
    rng = np.random.default_rng(0)
    for t in range(8):
        act = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
        a.step(act)
        b.road_substeps(frames, action=act)
        same_state(a, b, t)


def test_network_road_substeps_runs_on_every_family():
    import highwayenv_b200 as hb

    for env_id in ("roundabout-v0", "intersection-v0", "merge-v0"):
        env = hb.make(env_id, num_envs=16, autoreset_mode="Disabled")
        env.reset(seed=1)
        before = env.state_dict()
        env.road_substeps(15)
        after = env.state_dict()
        assert np.all(np.isfinite(after["x"])) and not np.array_equal(before["x"], after["x"]), env_id
