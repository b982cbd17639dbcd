"""Every registered scenario for a few steps at a small batch size (for compute-sanitizer memcheck / synccheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import highwayenv_b200 as hb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(0)
for env_id in sorted(hb.REGISTRY):
    for mode in ("SameStep", "NextStep"):
        env = hb.make(env_id, num_envs=n, autoreset_mode=mode)
        env.reset(seed=1)
        sp = env.single_action_space
        hi = sp.n if hasattr(sp, "n") else int(sp.high.max()) + 1
        for t in range(14):
            env.step(rng.integers(0, hi, size=(n,) + tuple(getattr(sp, "shape", ()) or ())).astype(np.int32))
        torch.cuda.synchronize()
    print(env_id, "ok", flush=True)
env = hb.make("highway-v0", num_envs=n, config={"vehicles_count": 100, "action": {"type": "ContinuousAction"}})
env.reset(seed=2)
for t in range(3):
    env.step(rng.uniform(-1, 1, size=(n, 2)).astype(np.float32))
torch.cuda.synchronize()
print("highway-v0 V=101 continuous ok", flush=True)
