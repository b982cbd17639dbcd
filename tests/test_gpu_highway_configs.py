"""Wider config coverage of the highway kernels against the C oracle: every thread mapping
(32 / 64 / 128 threads per env) at its boundaries, lane counts, observation options, frequencies,
reward options.  Teacher-forced each step + bit-exact resets."""
import numpy as np
import pytest

import hwy_oracle as ho

pytestmark = pytest.mark.gpu


def run_case(env_id, over, n, T, seed):
    import highwayenv_b200 as hb
    from highwayenv_b200.config import default_config

    cfg = default_config(env_id)
    cfg.update(over)
    env = hb.make(env_id, num_envs=n, config=over, autoreset_mode="Disabled")
    ocfg = dict(cfg)
    ocfg["_others_check_collisions"] = 0 if env_id == "highway-fast-v0" else 1
    oc = ho.cfg_from_dict(ocfg)
    ob = ho.OracleBatch(oc, n, seeds=range(seed, seed + n), threads=8)
    obs, _ = env.reset(seed=seed)
    assert np.array_equal(obs.cpu().numpy(), ob.reset()), "reset observation"
    sd = env.state_dict()
    for k in ("x", "y", "heading", "speed", "timer", "delta"):
        assert np.array_equal(sd[k], ob.a[k]), f"reset {k}"
    rng = np.random.default_rng(seed)
    for t in range(T):
        env.load_state_dict({k: ob.a[k].copy() for k in ob.a})
        if oc.action_type == 0:
            act = rng.integers(0, 5, size=n).astype(np.int32)
        else:
            act = rng.uniform(-1.2, 1.2, size=(n, 2)).astype(np.float32)  # exercises the clip
        o_obs, o_rew, o_term, o_trunc = ob.step(act)
        obs, rew, term, trunc, _ = env.step(act)
        sd = env.state_dict()
        for k in ("x", "y", "heading", "speed", "timer", "target_speed"):
            assert np.max(np.abs(sd[k] - ob.a[k])) <= 1e-7, (t, k)
        for k in ("lane", "target_lane", "crashed", "has_impact"):
            assert np.array_equal(sd[k].astype(np.int32), ob.a[k].astype(np.int32)), (t, k)
        assert np.max(np.abs(rew.cpu().numpy() - o_rew)) <= 1e-9
        assert np.array_equal(term.cpu().numpy(), o_term.astype(bool))
        assert np.array_equal(trunc.cpu().numpy(), o_trunc.astype(bool))
        assert np.max(np.abs(obs.cpu().numpy() - o_obs)) <= 1e-6


@pytest.mark.parametrize("vehicles", [1, 8, 31, 32, 63, 64, 100, 127])
def test_vehicle_counts_fast(vehicles):
    run_case("highway-fast-v0", {"vehicles_count": vehicles}, 24, 6, 100 + vehicles)


@pytest.mark.parametrize("vehicles", [15, 40, 70])
def test_vehicle_counts_all_pairs(vehicles):
    run_case("highway-v0", {"vehicles_count": vehicles}, 12, 4, 300 + vehicles)


@pytest.mark.parametrize("lanes", [1, 2, 5, 8])
def test_lane_counts(lanes):
    run_case("highway-fast-v0", {"vehicles_count": 30, "lanes_count": lanes}, 16, 6, 500 + lanes)


@pytest.mark.parametrize("obs", [
    {"type": "Kinematics", "vehicles_count": 3},
    {"type": "Kinematics", "vehicles_count": 12, "see_behind": True},
    {"type": "Kinematics", "absolute": True, "normalize": False},
    {"type": "Kinematics", "clip": False},
])
def test_observation_options(obs):
    run_case("highway-fast-v0", {"vehicles_count": 25, "observation": obs}, 16, 5, 700)


@pytest.mark.parametrize("over", [
    {"simulation_frequency": 10, "policy_frequency": 2, "duration": 4},
    {"simulation_frequency": 15, "policy_frequency": 5},
    {"normalize_reward": False, "collision_reward": -2.0, "right_lane_reward": 0.3, "high_speed_reward": 0.7,
     "reward_speed_range": [15, 28]},
    {"offroad_terminal": True, "initial_lane_id": 0, "ego_spacing": 1.0, "vehicles_density": 1.7},
    {"action": {"type": "DiscreteMetaAction", "target_speeds": [10, 15, 20, 25, 30]}},
    {"action": {"type": "ContinuousAction", "acceleration_range": [-3.0, 2.0], "steering_range": [-0.3, 0.3]}},
])
def test_misc_options(over):
    run_case("highway-fast-v0", dict({"vehicles_count": 20}, **over), 16, 6, 900)


def test_get_available_actions_on_device_state():
    """env.get_available_actions(): bool [N, 5] from the device state (action.py:262-299)"""
    import highwayenv_b200 as hb

    env = hb.make("highway-fast-v0", num_envs=32, config={"lanes_count": 3})
    env.reset(seed=0)
    m = env.get_available_actions().cpu().numpy()
    sd = env.state_dict()
    lane, si = sd["lane"][:, 0], sd["speed_index"]
    assert m.shape == (32, 5) and m[:, 1].all()
    assert np.array_equal(m[:, 0], lane > 0) and np.array_equal(m[:, 2], lane < 2)
    assert np.array_equal(m[:, 3], si < 2) and np.array_equal(m[:, 4], si > 0)
    for _ in range(3):
        env.step(np.full(32, 3, dtype=np.int32))  # FASTER (envs that crash meanwhile are reset: SameStep)
    m = env.get_available_actions().cpu().numpy()
    sd = env.state_dict()
    lane, si = sd["lane"][:, 0], sd["speed_index"]
    assert np.array_equal(m[:, 0], lane > 0) and np.array_equal(m[:, 2], lane < 2)
    assert np.array_equal(m[:, 3], si < 2) and np.array_equal(m[:, 4], si > 0) and (si == 2).any()


def test_single_env_facade_matches_the_reference_episode():
    """make_single: gymnasium.Env-shaped returns (numpy obs, python float / bool, info of python scalars incl.
    info["rewards"]); reset(seed=s) + the golden actions reproduce the reference's episode for seed s."""
    import highwayenv_b200 as hb
    from parity_utils import load_golden

    g = load_golden("highway_fast_v20")
    env = hb.make_single("highway-fast-v0")
    seed = int(g["seeds"][1])
    obs, info = env.reset(seed=seed)
    assert isinstance(obs, np.ndarray) and obs.shape == (5, 5) and obs.dtype == np.float32
    assert np.max(np.abs(obs - g["obs"][1, 0])) <= 1e-6
    assert set(info) >= {"speed", "crashed"}
    for t in range(8):
        obs, reward, terminated, truncated, info = env.step(int(g["actions"][1, t]))
        assert isinstance(reward, float) and isinstance(terminated, bool) and isinstance(truncated, bool)
        assert abs(reward - g["reward"][1, t]) <= 1e-6 and terminated == bool(g["terminated"][1, t])
        assert np.max(np.abs(obs - g["obs"][1, t + 1])) <= 1e-5
        assert set(info["rewards"]) == {"collision_reward", "right_lane_reward", "high_speed_reward", "on_road_reward"}
        assert isinstance(info["speed"], float) and isinstance(info["crashed"], bool)
        if terminated or truncated:
            break
    assert env.action_space.n == 5 and 1 in env.get_available_actions()
    env2 = hb.make_single("intersection-v0")
    o, _ = env2.reset(seed=3)
    o, r, te, tr, info = env2.step(1)
    assert o.shape == (15, 7) and isinstance(r, float) and "rewards" in info
