"""Per-core speed of the UNMODIFIED Python reference on every BASELINE.json config (build container only).

    python tools/time_reference.py [--seconds 20] [--out profiles/r2_python_reference.json]

BASELINE.md §4 asks for two timings of the reference itself, never of a restatement:
  (i)  `Road.act() + Road.step(dt)` alone (highway_env/road/road.py:464-481) — the loop north_star names;
  (ii) the full `env.step` (highway_env/envs/common/abstract.py:259-285), resets included when an episode ends.
Both are reported per core (one process, one env, random actions) in env-steps/s; (i) is converted with the
config's substeps per step so that the two are comparable.  The reference is imported from /root/reference under the
stub gymnasium of oracle/shim (oracle/ref_harness.py); it cannot travel to the GPU box, so bench.py quotes this
file as "measured elsewhere".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

import ref_harness as rh  # noqa: E402
from bench import CONFIGS, cpu_model  # noqa: E402


def time_config(key: str, seconds: float) -> dict:
    c = CONFIGS[key]
    env = rh.make_reference_env(c["env_id"], c["config"])
    rng = np.random.default_rng(1234)
    env.reset(seed=0)
    sample = env.action_space.sample

    def act():
        if c["actions"] == "box2":
            return rng.uniform(-1, 1, size=2).astype(np.float32)
        return int(rng.integers(0, 3 if c["actions"] == "discrete3" else 5))

    # (ii) full env.step with resets
    for _ in range(5):
        _, _, te, tr, _ = env.step(act())
        if te or tr:
            env.reset()
    steps, resets, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, te, tr, _ = env.step(act())
        steps += 1
        if te or tr:
            env.reset()
            resets += 1
    full = steps / (time.perf_counter() - t0)
    # (i) Road.act + Road.step only, on live episodes (the ego keeps its last action)
    env.reset(seed=1)
    dt = 1 / env.config["simulation_frequency"]
    sub = int(env.config["simulation_frequency"] // env.config["policy_frequency"])
    n_sub, t_road, t_end = 0, 0.0, time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        env.action_type.act(act())
        a = time.perf_counter()
        for _ in range(sub):
            env.road.act()
            env.road.step(dt)
        t_road += time.perf_counter() - a
        n_sub += sub
        env.time += 1 / env.config["policy_frequency"]
        if env._is_terminated() or env._is_truncated():
            env.reset()
    road_sub = n_sub / t_road
    del sample
    return {"env_step_per_core": full, "road_substeps_per_core": road_sub,
            "road_only_env_steps_per_core": road_sub / sub, "substeps_per_step": sub,
            "steps_timed": steps, "resets": resets, "unit": "env-steps/s per core",
            "vehicles": len(env.road.vehicles)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_python_reference.json"))
    args = ap.parse_args()
    if not rh.reference_available():
        raise SystemExit("the reference is not mounted here (/root/reference)")
    out = {"host": f"build container, {cpu_model()}, 1 process = 1 core, python {sys.version.split()[0]}, numpy {np.__version__}",
           "seconds_per_timing": args.seconds, "configs": {}}
    for key in CONFIGS:
        out["configs"][key] = time_config(key, args.seconds)
        print(key, json.dumps(out["configs"][key]), flush=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
