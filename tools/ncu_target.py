"""Small driver for `ncu` captures of one env family's step kernel:
    ncu --set full --clock-control none --import-source on -k regex:network_step -s 3 -c 1 \
        -o gpurun_out/net python tools/ncu_target.py intersection-v0 8192 OccupancyGrid"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import highwayenv_b200 as hb  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "intersection-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cfg = {"observation": {"type": sys.argv[3]}} if len(sys.argv) > 3 else None
if cfg and sys.argv[3] == "TimeToCollision":
    cfg["observation"]["horizon"] = 10
env = hb.make(env_id, num_envs=n, config=cfg, autoreset_mode="Disabled")
env.reset(seed=0)
hi = env.single_action_space.n
g = torch.Generator(device="cuda")
g.manual_seed(1)
acts = torch.randint(0, hi, (8, n), generator=g, device="cuda", dtype=torch.int32)
for t in range(8):
    env.step(acts[t])
torch.cuda.synchronize()
