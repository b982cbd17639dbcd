"""torch.ops.hwyb200.* (highwayenv_b200/torch_ops.py): the registered operators, their loud CPU failure, and on the GPU
that a run through the operators is the run through env.step."""
import pytest
import torch

from highwayenv_b200 import torch_ops  # noqa: F401  (registers the library)


def test_ops_registered_and_cuda_only():
    assert hasattr(torch.ops.hwyb200, "step") and hasattr(torch.ops.hwyb200, "reset") and hasattr(torch.ops.hwyb200, "observe")
    with pytest.raises(NotImplementedError):  # no CPU kernel behind the op: the dispatcher refuses, nothing falls back
        torch.ops.hwyb200.step(1, torch.zeros(4, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no live env"):
        torch.ops.hwyb200.reset(10 ** 9, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["highway-fast-v0", "roundabout-v0", "intersection-v0"])
def test_ops_match_env_step(env_id):
    import highwayenv_b200 as hb

    n = 64
    a, b = hb.make(env_id, num_envs=n), hb.make(env_id, num_envs=n)
    h = torch_ops.register(b)
    obs_a, _ = a.reset(seed=11)
    obs_b = torch.ops.hwyb200.reset(h, 11)
    assert torch.equal(obs_a, obs_b)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    hi = a.single_action_space.n
    for _ in range(12):
        act = torch.randint(0, hi, (n,), generator=gen, device="cuda", dtype=torch.int32)
        oa, ra, ta, ua, _ = a.step(act)
        ob, rb, tb, ub = torch.ops.hwyb200.step(h, act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(ua, ub)
    assert torch.equal(torch.ops.hwyb200.observe(h), b.observe())
    del b  # the table holds a weak reference only
    import gc

    gc.collect()
    with pytest.raises(RuntimeError, match="no live env"):
        torch.ops.hwyb200.observe(h)
