"""Small driver for `ncu` captures: steps one bench config (bench.py CONFIGS, SameStep autoreset, so the reset and
observe kernels of the step are launched too) a few times.

    ncu --set full --clock-control none --import-source on -k regex:network_step -s 3 -c 1 \
        -o gpurun_out/net python tools/ncu_target.py cfg3 [n_envs] [steps]
    (legacy form: python tools/ncu_target.py intersection-v0 8192 OccupancyGrid)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import highwayenv_b200 as hb  # noqa: E402
from bench import CONFIGS, _make_actions  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
if key in CONFIGS:
    c = CONFIGS[key]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else c["envs_per_gpu"]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    env = hb.make(c["env_id"], num_envs=n, config=c["config"])
    kind = c["actions"]
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    steps = 8
    cfg = {"observation": {"type": sys.argv[3]}} if len(sys.argv) > 3 else None
    if cfg and sys.argv[3] == "TimeToCollision":
        cfg["observation"]["horizon"] = 10
    env = hb.make(key, num_envs=n, config=cfg, autoreset_mode="Disabled")
    kind = "discrete3" if env.single_action_space.n == 3 else "discrete5"
env.reset(seed=0)
g = torch.Generator(device="cuda")
g.manual_seed(1234)
acts = _make_actions(kind, n, steps, g, torch.device("cuda"), torch)
for t in range(steps):
    env.step(acts[t])
torch.cuda.synchronize()
