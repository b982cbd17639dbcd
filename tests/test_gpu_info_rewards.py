"""info["rewards"] (AbstractEnv._info, envs/common/abstract.py:213-216): the un-weighted terms of `_rewards` of every
step, per env family.  Recombining them with the reference's `_reward` formula must give the step's reward."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def lmap(v, x, y):
    return y[0] + (v - x[0]) * (y[1] - y[0]) / (x[1] - x[0])


def rollout(env_id, n, steps, n_act, config=None, seed=5):
    import highwayenv_b200 as hb

    env = hb.make(env_id, num_envs=n, config=config, autoreset_mode="Disabled")
    env.reset(seed=seed)
    rng = np.random.default_rng(0)
    out = []
    for _ in range(steps):
        act = rng.integers(0, n_act, size=n).astype(np.int32)
        _, rew, term, trunc, info = env.step(act)
        terms = {k: v.cpu().numpy().copy() for k, v in info["rewards"].items()}
        out.append((act, rew.cpu().numpy().copy(), terms))
    return env, out


def test_highway_reward_terms():
    env, out = rollout("highway-fast-v0", 256, 12, 5, {"vehicles_count": 30})
    cfg = env.config
    assert list(out[0][2]) == ["collision_reward", "right_lane_reward", "high_speed_reward", "on_road_reward"]
    seen_crash = False
    for act, rew, t in out:  # highway_env.py:100-137
        r = sum(cfg.get(k, 0) * t[k] for k in t)
        r = lmap(r, [cfg["collision_reward"], cfg["high_speed_reward"] + cfg["right_lane_reward"]], [0, 1])
        r = r * t["on_road_reward"]
        assert np.max(np.abs(r - rew)) <= 1e-12
        assert set(np.unique(t["collision_reward"])) <= {0.0, 1.0}
        seen_crash |= bool(t["collision_reward"].any())
    assert seen_crash


def test_roundabout_and_merge_reward_terms():
    env, out = rollout("roundabout-v0", 128, 8, 5)
    cfg = env.config
    for act, rew, t in out:  # roundabout_env.py:44-65
        r = sum(cfg.get(k, 0) * t[k] for k in t)
        r = lmap(r, [cfg["collision_reward"], cfg["high_speed_reward"]], [0, 1]) * t["on_road_reward"]
        assert np.max(np.abs(r - rew)) <= 1e-12
        assert np.array_equal(t["lane_change_reward"], np.isin(act, (0, 2)).astype(float))
    env, out = rollout("merge-v0", 128, 8, 5)
    cfg = env.config
    for act, rew, t in out:  # merge_env.py:39-77
        r = sum(cfg.get(k, 0) * t[k] for k in t)
        r = lmap(r, [cfg["collision_reward"] + cfg["merging_speed_reward"],
                     cfg["high_speed_reward"] + cfg["right_lane_reward"]], [0, 1])
        assert np.max(np.abs(r - rew)) <= 1e-12


def test_intersection_reward_terms():
    for env_id, n_act, A in (("intersection-v0", 3, 1), ("intersection-multi-agent-v0", 3, 2)):
        import highwayenv_b200 as hb

        env = hb.make(env_id, num_envs=96, autoreset_mode="Disabled")
        env.reset(seed=9)
        rng = np.random.default_rng(1)
        cfg = env.config
        for _ in range(10):
            act = rng.integers(0, n_act, size=(96, A) if A > 1 else 96).astype(np.int32)
            _, rew, _, _, info = env.step(act)
            t = {k: v.cpu().numpy() for k, v in info["rewards"].items()}
            assert list(t) == ["collision_reward", "high_speed_reward", "arrived_reward", "on_road_reward"]
            if A == 1:  # _agent_reward, intersection_env.py:79-93
                r = sum(cfg.get(k, 0) * t[k] for k in t)
                r = np.where(t["arrived_reward"] > 0, cfg["arrived_reward"], r) * t["on_road_reward"]
                if cfg["normalize_reward"]:
                    r = lmap(r, [cfg["collision_reward"], cfg["arrived_reward"]], [0, 1])
                assert np.max(np.abs(r - rew.cpu().numpy())) <= 1e-12
            else:  # every term is the mean over the agents (:68-77)
                assert np.all((t["on_road_reward"] >= 0) & (t["on_road_reward"] <= 1))
                assert np.all(np.isin(t["collision_reward"] * A, np.arange(A + 1)))
