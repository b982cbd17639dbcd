"""BASELINE.json's configurations at their full per-GPU sizes, checked through size-independent properties:

* batch-partition invariance — env i of the full batch (seed0 + i) must evolve bit-identically to the same
  env stepped inside a small batch placed at `env_index_offset=i0` with the same actions, through resets:
  any cross-env leak, indexing or scheduling dependence at scale breaks it; the small batches are the sizes
  the oracle / reference parity tests cover, so parity carries over to every env of the full batch;
* bookkeeping invariants of `AbstractEnv.step` (time, flags, observation ranges, generator increments)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = [
    # (name, env id, config, envs, steps, action kind)
    ("cfg1 highway-fast-v0 V=51", "highway-fast-v0", {"vehicles_count": 50}, 4096, 40, 5),
    ("cfg2 intersection-v0 OccupancyGrid", "intersection-v0", {"observation": {"type": "OccupancyGrid"}}, 8192, 18, 3),
    ("cfg3 roundabout-v0 TimeToCollision", "roundabout-v0",
     {"observation": {"type": "TimeToCollision", "horizon": 10}}, 16384, 15, 5),
    ("cfg4 highway-v0 V=101 ContinuousAction", "highway-v0",
     {"vehicles_count": 100, "action": {"type": "ContinuousAction"}}, 8192, 12, 0),
]
SUB = 48          # envs per probe batch
OFFSETS = (0, 0.37, 1.0)   # where the probe batches sit in the full index range


def _actions(kind, steps, n, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    if kind == 0:
        return torch.rand((steps, n, 2), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    return torch.randint(0, kind, (steps, n), generator=g, device="cuda", dtype=torch.int32)


def _same(a, b, ctx):
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.dtype.kind == "f":
            assert np.array_equal(x, y, equal_nan=True), f"{ctx}: {k} differs by {np.nanmax(np.abs(x - y))}"
        else:
            assert np.array_equal(x, y), f"{ctx}: {k}"


def _live_view(sd, lo, hi):
    """state rows [lo, hi); slots beyond a dynamic population's `count` are don't-care"""
    out = {}
    count = sd.get("count")
    for k, v in sd.items():
        if k == "rng":
            out[k] = v[:, lo:hi]
            continue
        v = v[lo:hi]
        if count is not None and v.ndim >= 2 and v.shape[1] == 32:
            live = np.arange(32)[None, :] < count[lo:hi, None]
            v = np.where(live.reshape(live.shape + (1,) * (v.ndim - 2)), v, 0)
        out[k] = v
    return out


@pytest.mark.parametrize("name,env_id,cfg,n,steps,kind", CONFIGS, ids=[c[0].split()[0] for c in CONFIGS])
def test_full_size_partition_invariance_and_bookkeeping(name, env_id, cfg, n, steps, kind):
    import highwayenv_b200 as hb

    seed0 = 1000
    full = hb.make(env_id, num_envs=n, config=cfg)
    obs, _ = full.reset(seed=seed0)
    acts = _actions(kind, steps, n, 77)
    starts = sorted({min(int(f * (n - SUB)), n - SUB) for f in OFFSETS})
    probes = []
    for i0 in starts:
        p = hb.make(env_id, num_envs=SUB, config=cfg, env_index_offset=i0)
        p.reset(seed=seed0)
        _same(_live_view(full.state_dict(), i0, i0 + SUB), _live_view(p.state_dict(), 0, SUB), f"{name} reset @{i0}")
        probes.append((i0, p))
    duration = float(full.config["duration"])
    policy_dt = 1.0 / full.config["policy_frequency"]
    inc0 = full.state_dict().get("rng")
    resets = 0
    for t in range(steps):
        t_before = full.state_dict()["time"]
        obs, rew, term, trunc, info = full.step(acts[t])
        sd = full.state_dict()
        o, r = obs.cpu().numpy(), rew.cpu().numpy()
        te, tr = term.cpu().numpy(), trunc.cpu().numpy()
        done = te | tr
        resets += int(done.sum())
        # ---- bookkeeping (abstract.py:259-285, SameStep autoreset)
        assert np.isfinite(o).all() and np.isfinite(r).all()
        assert o.min() >= -1.0 - 1e-6 and o.max() <= 1.0 + 1e-6
        assert np.all(sd["time"][done] == 0.0)
        assert np.allclose(sd["time"][~done], t_before[~done] + policy_dt, rtol=0, atol=1e-12)
        assert np.array_equal(tr, t_before + policy_dt >= duration - 1e-12)
        crashed = info["crashed"].cpu().numpy().astype(bool)
        assert np.all(te[crashed])  # a crashed controlled vehicle always terminates the episode
        if inc0 is not None:
            assert np.array_equal(sd["rng"][2:4], inc0[2:4])  # PCG64 increments are per-env constants
        if "count" in sd:
            assert sd["count"].min() >= 1 and sd["count"].max() <= 32
        # ---- partition invariance
        for i0, p in probes:
            po, pr, pte, ptr, _ = p.step(acts[t, i0:i0 + SUB].contiguous())
            ctx = f"{name} t={t} @{i0}"
            assert np.array_equal(po.cpu().numpy(), o[i0:i0 + SUB]), ctx
            assert np.array_equal(pr.cpu().numpy(), r[i0:i0 + SUB]), ctx
            assert np.array_equal(pte.cpu().numpy(), te[i0:i0 + SUB]) and np.array_equal(ptr.cpu().numpy(), tr[i0:i0 + SUB]), ctx
            _same(_live_view(sd, i0, i0 + SUB), _live_view(p.state_dict(), 0, SUB), ctx)
    assert resets > 0, "the run must cross episode ends so that the device autoreset is exercised at scale"
