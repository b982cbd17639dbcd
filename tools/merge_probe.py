"""Small merge-v0 run for `compute-sanitizer` / ncu: reset, teacher-forced steps from the golden fixture, autoreset."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import highwayenv_b200 as hb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
env = hb.make("merge-v0", num_envs=n)
env.reset(seed=0)
torch.cuda.synchronize()
print("reset ok", flush=True)
rng = np.random.default_rng(0)
for t in range(25):
    env.step(rng.integers(0, 5, size=n).astype(np.int32))
    torch.cuda.synchronize()
print("steps ok", flush=True)
