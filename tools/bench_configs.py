"""Device-timed throughput of the OTHER BASELINE.json configs (the headline config lives in bench.py).

    python tools/bench_configs.py [--steps K]

One JSON line per config: env-steps/s through the C ABI with actions resident in HBM, CUDA events
around each launch, L2 flushed between launches.  highway configs use the fused SameStep
autoreset; roundabout-v0 the step kernel followed by the device reset + observe kernels."""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import highwayenv_b200 as hb  # noqa: E402

CONFIGS = [
    ("cfg0 highway-fast-v0 V=21", "highway-fast-v0", None, 4096, "SameStep", "discrete"),
    ("cfg1 highway-fast-v0 V=51 (headline)", "highway-fast-v0", {"vehicles_count": 50}, 4096, "SameStep", "discrete"),
    ("cfg1 at 8192 envs", "highway-fast-v0", {"vehicles_count": 50}, 8192, "SameStep", "discrete"),
    ("cfg4 highway-v0 V=101 ContinuousAction", "highway-v0",
     {"vehicles_count": 100, "action": {"type": "ContinuousAction"}}, 8192, "SameStep", "box"),
    ("highway-v0 defaults V=51", "highway-v0", None, 4096, "SameStep", "discrete"),
    ("cfg3 roundabout-v0 TimeToCollision", "roundabout-v0",
     {"observation": {"type": "TimeToCollision", "horizon": 10}}, 8192, "SameStep", "discrete"),
    ("roundabout-v0 defaults (Kinematics)", "roundabout-v0", None, 8192, "SameStep", "discrete"),
    ("merge-v0 defaults (Kinematics, 5 vehicles + obstacle)", "merge-v0", None, 8192, "SameStep", "discrete"),
    ("two-way-v0 defaults (TimeToCollision, 6 vehicles)", "two-way-v0", None, 8192, "SameStep", "discrete"),
    ("u-turn-v0 defaults (TimeToCollision 16 s, 7 vehicles)", "u-turn-v0", None, 8192, "SameStep", "discrete"),
    ("cfg2 intersection-v0 OccupancyGrid", "intersection-v0", {"observation": {"type": "OccupancyGrid"}}, 8192,
     "SameStep", "discrete3"),
    ("intersection-v0 defaults (Kinematics 15x7)", "intersection-v0", None, 8192, "SameStep", "discrete3"),
    ("cfg2 intersection-v0 OccupancyGrid, step kernel only", "intersection-v0",
     {"observation": {"type": "OccupancyGrid"}}, 8192, "Disabled", "discrete3"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--only", default=None, help="substring filter on the config name")
    args = ap.parse_args()
    dev = torch.device("cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for name, env_id, cfg, n, mode, akind in CONFIGS:
        if args.only and args.only not in name:
            continue
        env = hb.make(env_id, num_envs=n, config=cfg, autoreset_mode=mode)
        env.reset(seed=0)
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        K, W = args.steps, 5
        if akind.startswith("discrete"):
            acts = torch.randint(0, 3 if akind == "discrete3" else 5, (K + W, n), generator=g, device=dev, dtype=torch.int32)
        else:
            acts = (torch.rand((K + W, n, 2), generator=g, device=dev, dtype=torch.float32) * 2 - 1)
        for t in range(W):
            env.step(acts[t])
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(K):
            flush.fill_(k & 0xFF)
            ev[k][0].record()
            env.step(acts[W + k])
            ev[k][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / K
        print(json.dumps({"config": name, "env_id": env_id, "envs": n, "autoreset": mode,
                          "ms_per_step": ms, "env_steps_per_s": n / (ms * 1e-3)}), flush=True)
        del env


if __name__ == "__main__":
    main()
