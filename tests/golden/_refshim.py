"""Import shim for the UNMODIFIED reference (only usable where /root/reference exists).

TEST/GOLDEN-GENERATION INFRASTRUCTURE ONLY.  Nothing in the product path, the
`-m gpu` tests, `smoke()` or `bench.py` imports this module: /root/reference does
not exist on the GPU box.  It is used (a) by `make_golden.py` to generate the
committed fixtures and (b) by the optional `tests/test_oracle_vs_reference.py`
which is skipped when the reference is absent.

The reference cannot be imported as-is in this image:
  * `gym_go/__init__.py:1` imports `gym` (not installed),
  * `gym_go/envs/go_env.py:6` -> `gym_go/rendering.py:2` imports `pyglet` (not installed),
  * `gym_go/gogame.py:250` uses `np.int`, removed from numpy >= 1.24.
We insert stub modules for `gym`/`pyglet` and alias `np.int`; the reference's
own source files are loaded untouched from /root/reference.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("GYMGO_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gym_go", "gogame.py"))


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        registry = {}

        class Env(object):
            pass

        class _Box(object):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        class _Discrete(object):
            def __init__(self, n):
                self.n = n

        def register(id, entry_point, **kw):
            registry[id] = entry_point

        def make(id, **kwargs):
            import importlib
            name = id.split(":")[-1]
            if ":" in id:
                importlib.import_module(id.split(":")[0])
            mod_name, cls_name = registry[name].split(":")
            mod = importlib.import_module(mod_name)
            return getattr(mod, cls_name)(**kwargs)

        spaces = types.ModuleType("gym.spaces")
        spaces.Box, spaces.Discrete = _Box, _Discrete
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")
        registration.register = register
        envs.registration = registration
        gym.Env, gym.spaces, gym.envs, gym.make, gym.register = Env, spaces, envs, make, register
        gym._registry = registry
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                            "gym.envs.registration": registration})
    if "pyglet" not in sys.modules:
        sys.modules["pyglet"] = types.ModuleType("pyglet")
    if not hasattr(np, "int"):
        np.int = int  # noqa: reference gogame.py:250
    if not hasattr(np, "bool"):
        np.bool = bool  # noqa: reference gogame.py:261


def load_reference():
    """Returns (gym_stub, gogame_module, govars_module) of the real reference."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    # Our repo ships a drop-in package that is ALSO called gym_go; make sure the
    # reference's one wins for this process by purging + path priority.
    for k in [k for k in sys.modules if k == "gym_go" or k.startswith("gym_go.")]:
        del sys.modules[k]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gym_go  # noqa
        from gym_go import gogame, govars
    assert os.path.realpath(gogame.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), gogame.__file__
    return sys.modules["gym"], gogame, govars
