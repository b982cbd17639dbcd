"""The batched envs as PyTorch custom operators (``torch.ops.hwyb200.*``).

``torch.library`` registration over the same C ABI the env classes call (include/hwyb200.h via ``_native``): a
training loop written against operators — e.g. one that wants the step to show up in a profiler trace or behind a
dispatcher key — can hold an integer handle instead of the Python env object:

    import highwayenv_b200 as hb
    from highwayenv_b200 import torch_ops
    h = torch_ops.register(hb.make("highway-fast-v0", num_envs=4096))
    obs = torch.ops.hwyb200.reset(h, 0)
    obs, reward, terminated, truncated = torch.ops.hwyb200.step(h, actions)

The operators run on the env's device and stream; results are fresh tensors (the env's own output buffers are
reused from step to step, which an operator's outputs must not be).  CUDA only — like the env classes there is no CPU
implementation, and calling the op with CPU tensors fails in the dispatcher.
"""
from __future__ import annotations

import itertools
import weakref

import torch

_ENVS: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_NEXT = itertools.count(1)

_LIB = torch.library.Library("hwyb200", "DEF")
_LIB.define("reset(int handle, int seed) -> Tensor")
_LIB.define("step(int handle, Tensor action) -> (Tensor, Tensor, Tensor, Tensor)")
_LIB.define("observe(int handle) -> Tensor")


def register(env) -> int:
    """Give `env` (any env of ``highwayenv_b200.make``) a handle for the operators; the table holds a weak reference."""
    handle = next(_NEXT)
    _ENVS[handle] = env
    return handle


def _env(handle: int):
    env = _ENVS.get(int(handle))
    if env is None:
        raise RuntimeError(f"hwyb200: no live env behind handle {handle} (register() it, and keep the env alive)")
    return env


def _reset(handle: int, seed: int) -> torch.Tensor:
    obs, _ = _env(handle).reset(seed=int(seed))
    return obs.clone()


def _step(handle: int, action: torch.Tensor):
    obs, reward, terminated, truncated, _ = _env(handle).step(action)
    return obs.clone(), reward.clone(), terminated.clone(), truncated.clone()


def _observe(handle: int) -> torch.Tensor:
    env = _env(handle)
    return env.observe().clone() if hasattr(env, "observe") else env._obs.clone()


# `reset` and `observe` take no tensor argument, so the dispatcher has no device to key on: CompositeExplicitAutograd
# routes them to the implementation whatever the (absent) inputs; `step` is keyed by the action tensor's device.
_LIB.impl("reset", _reset, "CompositeExplicitAutograd")
_LIB.impl("observe", _observe, "CompositeExplicitAutograd")
_LIB.impl("step", _step, "CUDA")
