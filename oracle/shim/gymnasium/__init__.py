"""Minimal stand-in for `gymnasium` (pinned 1.3.0 by the reference, not installed in
this image). TEST INFRASTRUCTURE ONLY: lets `oracle/ref_harness.py` import the
unmodified reference from /root/reference. Only the subset the reference touches at
import time and on reset/step is provided; RNG seeding follows gymnasium's
`seeding.np_random` (Generator(PCG64(SeedSequence(seed)))).
"""
from __future__ import annotations

import numpy as np

from . import spaces  # noqa: F401
from . import logger  # noqa: F401


class Env:
    metadata: dict = {"render_modes": []}
    render_mode = None
    spec = None
    _np_random = None
    _np_random_seed = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
            self._np_random_seed = seed

    @property
    def np_random(self):
        if self._np_random is None:
            ss = np.random.SeedSequence()
            self._np_random = np.random.Generator(np.random.PCG64(ss))
            self._np_random_seed = ss.entropy
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    @classmethod
    def wrapper_spec(cls, **kwargs):
        return {"name": cls.__name__, "kwargs": kwargs}


def make(id, **kwargs):
    from .envs.registration import registry
    import importlib

    if ":" in id:
        mod, id = id.split(":")
        importlib.import_module(mod)
    spec = registry[id]
    mod, cls = spec["entry_point"].split(":")
    env_cls = getattr(importlib.import_module(mod), cls)
    return env_cls(**{**spec.get("kwargs", {}), **kwargs})
