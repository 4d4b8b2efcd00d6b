"""A minimal stand-in for the `gym` package (not installed in this image): Env, spaces.Box/Discrete, a registry and
make() resolving 'package:id' entry points the way gym does.  TEST INFRASTRUCTURE: tests install it in sys.modules
BEFORE importing gym_go, exactly where a user's real gym would be."""
import importlib
import sys
import types

import numpy as np


def install():
    gym = types.ModuleType("gym")
    registry = {}

    class Env(object):
        metadata = {}

    class Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

    class Discrete(object):
        def __init__(self, n):
            self.n = int(n)

        def contains(self, a):
            return 0 <= int(a) < self.n

    def register(id, entry_point, **kwargs):
        if id in registry:
            raise ValueError("Cannot re-register id: %s" % id)
        registry[id] = entry_point

    def make(id, **kwargs):
        package, _, name = id.rpartition(":")
        if package:
            importlib.import_module(package)            # 'gym_go:go-v0' imports gym_go, which registers the ids
        module, cls = registry[name].split(":")
        return getattr(importlib.import_module(module), cls)(**kwargs)

    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Discrete = Box, Discrete
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = register
    envs.registration = registration
    gym.Env, gym.spaces, gym.envs, gym.make, gym.register, gym.registry = Env, spaces, envs, make, register, registry
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs, "gym.envs.registration": registration})
    return gym
