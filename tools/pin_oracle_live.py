"""Pin the C oracle to the LIVE reference on seeds that are in no fixture (build container only: needs /root/reference).

    python tools/pin_oracle_live.py [n_highway_seeds=80] [n_network_seeds=8] [n_intersection_seeds=16]

Runs the checks of tests/test_oracle_live.py, tests/test_net_oracle_live.py and tests/live_intersection_worker.py over
many fresh seeds (teacher-forced on every step; highway and intersection ids also free-running from the common seed,
intersection with the numpy generator words compared after every step) and writes the tally to
profiles/r2_oracle_pin_live.json.  Any mismatch beyond the tests' tolerances raises."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import hwy_oracle as ho  # noqa: E402
import ref_harness as rh  # noqa: E402
from parity_utils import compare_state, well_conditioned  # noqa: E402

assert rh.reference_available(), "the reference is not mounted here"
import test_net_oracle_live as tnl  # noqa: E402
import test_oracle_live as tol_  # noqa: E402

n_hw = int(sys.argv[1]) if len(sys.argv) > 1 else 80
n_net = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_int = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = {"tool": "tools/pin_oracle_live.py", "reference": "/root/reference (unmodified, live)", "cases": []}
t0 = time.time()


def free_running_highway(env_id, over, T, seed):
    """Common seed, then both run on their own under the same actions; compared while the golden state is well
    conditioned (parity_utils.well_conditioned).  Returns (steps compared, steps run, worst float difference)."""
    env = rh.make_reference_env(env_id, over)
    cfg = dict(env.config)
    cfg["_others_check_collisions"] = 0 if env_id == "highway-fast-v0" else 1
    oc = ho.cfg_from_dict(cfg)
    ob = ho.OracleBatch(oc, 1, seeds=[seed])
    obs_ref, _ = env.reset(seed=seed)
    assert np.array_equal(ob.reset()[0], obs_ref)
    rng = np.random.default_rng(seed + 1)
    compared, worst = 0, 0.0
    for t in range(T):
        if oc.action_type == 0:
            a = int(rng.integers(5))
            act = [a]
        else:
            a = rng.uniform(-1, 1, size=2).astype(np.float32)
            act = a[None]
        o, r, te, tr, _ = env.step(a)
        oo, ro, teo, tro = ob.step(act)
        st = rh.dump_state(env)
        if not well_conditioned(st):
            return compared, t + 1, worst
        got = {k: ob.a[k][0] for k in ob.a if k not in ("speed_index", "time")}
        got["speed_index"] = ob.a["speed_index"][0]
        worst = max(worst, compare_state(st, got, ctx=f"{env_id} free-running seed {seed} t={t}"))
        assert abs(r - ro[0]) < 1e-9 and te == bool(teo[0]) and tr == bool(tro[0])
        assert np.max(np.abs(o - oo[0])) <= 1e-6
        compared += 1
        if te or tr:
            break
    return compared, t + 1, worst


for env_id, over, T, seed0, n in (
        ("highway-fast-v0", {"vehicles_count": 50}, 12, 910000, n_hw),
        ("highway-fast-v0", None, 15, 920000, max(8, n_hw // 4)),
        ("highway-v0", {"vehicles_count": 30, "lanes_count": 5, "action": {"type": "ContinuousAction"}}, 6, 930000,
         max(8, n_hw // 4))):
    steps = cmp_free = run_free = 0
    worst = 0.0
    for seed in range(seed0, seed0 + n):
        tol_.test_oracle_matches_live_reference(env_id, over, T, seed)  # teacher-forced, asserts inside
        steps += T
        c, r, w = free_running_highway(env_id, over, T, seed)
        cmp_free, run_free, worst = cmp_free + c, run_free + r, max(worst, w)
    out["cases"].append({"env_id": env_id, "config": over, "seeds": [seed0, seed0 + n - 1], "teacher_forced_steps": steps,
                         "teacher_forced_failures": 0, "free_running_steps_compared": cmp_free,
                         "free_running_steps_run": run_free, "free_running_worst_abs_diff": worst})
    print(out["cases"][-1], flush=True)

for env_id, over, T in (("roundabout-v0", {"observation": {"type": "TimeToCollision", "horizon": 10}}, 11),
                        ("roundabout-v1", None, 11), ("merge-v0", None, 14), ("merge-v1", None, 14),
                        ("two-way-v0", None, 10), ("u-turn-v0", None, 10), ("u-turn-v1", None, 10)):
    seeds = tuple(range(940000, 940000 + n_net))
    tnl.test_net_oracle_matches_live_reference(env_id, over, T, seeds)  # teacher-forced, asserts inside
    out["cases"].append({"env_id": env_id, "config": over, "seeds": [seeds[0], seeds[-1]],
                         "teacher_forced_steps": T * len(seeds), "teacher_forced_failures": 0})
    print(out["cases"][-1], flush=True)

for env_id, obs in (("intersection-v0", "default"), ("intersection-v0", "OccupancyGrid"), ("intersection-v2", "default"),
                    ("intersection-multi-agent-v0", "default")):
    # own subprocess: IntersectionEnv._make_vehicles rewrites IDMVehicle class constants for the whole interpreter
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "live_intersection_worker.py"), env_id, obs, "950000",
                          str(n_int)], capture_output=True, text=True, timeout=3600)
    assert res.returncode == 0, res.stderr[-1500:]
    last = res.stdout.strip().splitlines()[-1].split()
    assert last[0] == "OK"
    out["cases"].append({"env_id": env_id, "observation": obs, "seeds": [950000, 950000 + n_int - 1],
                         "free_running_steps_compared": int(last[1]), "free_running_worst_abs_diff": float(last[2]),
                         "numpy_generator_words": "equal after every compared step"})
    print(out["cases"][-1], flush=True)

out["wall_s"] = time.time() - t0
with open(os.path.join(ROOT, "profiles", "r2_oracle_pin_live.json"), "w") as f:
    json.dump(out, f, indent=1)
print("written profiles/r2_oracle_pin_live.json")
