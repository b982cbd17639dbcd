"""Driver for `ncu` captures of the headline step kernel (highway-fast-v0, 50 vehicles, 4096 envs, SameStep)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import highwayenv_b200 as hb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = hb.make("highway-fast-v0", num_envs=n, config={"vehicles_count": 50})
env.reset(seed=0)
g = torch.Generator(device="cuda")
g.manual_seed(1234)
acts = torch.randint(0, 5, (12, n), generator=g, device="cuda", dtype=torch.int32)
for t in range(12):
    env.step(acts[t])
torch.cuda.synchronize()
