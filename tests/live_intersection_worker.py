"""Worker of tests/test_intersection_oracle_live.py: runs in ITS OWN process because IntersectionEnv._make_vehicles
rewrites IDMVehicle class constants for the whole interpreter (envs/intersection_env.py:262-265).

    python tests/live_intersection_worker.py <env_id> <obs type> <first seed> <n seeds>

Free-running: the oracle is reset with the same seed as the live reference and then both run on their own under the
same random actions; after every step the state, the reward / flags, the observation and the words of the numpy
generator are compared.  Prints `OK <env-steps compared> <worst float diff>`."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import net_oracle as no  # noqa: E402
import ref_harness as rh  # noqa: E402
from test_net_oracle_golden import compare_inter  # noqa: E402

env_id, obs_type, seed0, n_seeds = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
over = None if obs_type == "default" else {"observation": {"type": obs_type}}
worst, compared = 0.0, 0
for seed in range(seed0, seed0 + n_seeds):
    env = rh.make_reference_env(env_id, over)
    obs_ref, _ = env.reset(seed=seed)
    cfg = dict(env.config)
    net = rh.dump_network(env)
    A = int(cfg.get("controlled_vehicles", 1))
    ob = no.IntersectionOracle(no.graph_from_arrays(net), no.cfg_from_dict(cfg), 1, net, cfg)
    ob.reset_env(0, seed=seed)
    st = rh.dump_state(env, 32)
    compare_inter(st, ob.a, 0, f"{env_id} seed {seed} reset", tol=1e-8)
    assert np.max(np.abs(ob.observe().reshape(np.asarray(obs_ref).shape) - obs_ref)) <= 1e-6
    rng = np.random.default_rng(seed)
    m64 = (1 << 64) - 1
    for t in range(int(cfg["duration"] * cfg["policy_frequency"]) + 1):
        a = rng.integers(0, 3, size=A)
        o, r, te, tr, _ = env.step(tuple(int(x) for x in a) if A > 1 else int(a[0]))
        oo, ro, teo, tro = ob.step(a.reshape(1, A).astype(np.int32) if A > 1 else a.astype(np.int32))
        ctx = f"{env_id} seed {seed} t={t}"
        st = rh.dump_state(env, 32)
        n = int(st["count"])
        # utils.not_zero(speed) in the steering law (controller.py:166,178) amplifies 1-ulp libm differences by > 1e6
        # per policy step once a non-crashed vehicle crawls below ~1 m/s (tests/parity_utils.py well_conditioned):
        # free-running parity is asserted up to that point, teacher-forced parity (the fixtures) on every state
        if np.any(~st["crashed"][:n].astype(bool) & (np.abs(st["speed"][:n]) < 1.0)):
            break
        compare_inter(st, ob.a, 0, ctx, tol=1e-5)
        worst = max(worst, float(np.max(np.abs(st["x"][:n] - ob.a["x"][0][:n]))), float(np.max(np.abs(st["speed"][:n] - ob.a["speed"][0][:n]))))
        assert abs(r - ro[0]) <= 1e-6 and te == bool(teo[0]) and tr == bool(tro[0]), ctx
        assert np.max(np.abs(np.asarray(o, dtype=np.float64).reshape(-1) - oo[0].reshape(-1))) <= 1e-4, ctx
        w = env.np_random.bit_generator.state
        words = [w["state"]["state"] >> 64, w["state"]["state"] & m64, w["state"]["inc"] >> 64, w["state"]["inc"] & m64,
                 (int(w["has_uint32"]) << 32) | int(w["uinteger"])]
        assert [int(x) for x in ob.rng_words(0)] == words, ctx + " numpy stream"
        compared += 1
        if te or tr:
            break
print("OK", compared, worst)
