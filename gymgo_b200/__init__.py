"""gymgo_b200 - B200-native batched Go environment, drop-in for huangeddie/GymGo's hot path.

    from gymgo_b200 import make, BatchedGoEnv, gogame, govars
    env = make('gym_go:go-v0', size=7)                       # reference-compatible single board
    venv = BatchedGoEnv(batch_size=65536, size=9)            # packed device batch, one launch per ply

Importing the package never needs a GPU (the driver's build check runs on a CPU box); creating an
environment or calling gogame functions does - there is no CPU fallback."""
from . import _cabi, govars  # noqa: F401

__version__ = "0.1.0"

_IDS = {"go-v0": "GoEnv", "go-extrahard-v0": "GoExtraHardEnv"}


def make(id="gym_go:go-v0", **kwargs):
    """Local stand-in for gym.make('gym_go:go-v0', size=..., komi=..., reward_method=...)
    (gym_go/__init__.py:3-10); `gym` itself is optional."""
    from . import envs
    name = id.split(":")[-1]
    if name not in _IDS:
        raise KeyError("unknown environment id %r" % (id,))
    return getattr(envs, _IDS[name])(**kwargs)


def register_with_gym():
    """Register go-v0 / go-extrahard-v0 with gym or gymnasium when one of them is installed."""
    done = []
    for modname in ("gym", "gymnasium"):
        try:
            mod = __import__(modname + ".envs.registration", fromlist=["register"])
            for env_id, cls in _IDS.items():
                try:
                    mod.register(id=env_id, entry_point="gymgo_b200.envs:%s" % cls)
                except Exception:       # noqa: BLE001 - already registered
                    pass
            done.append(modname)
        except Exception:               # noqa: BLE001
            continue
    return done


def __getattr__(name):
    # lazy: these import torch
    if name in ("gogame", "engine", "envs"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name in ("BatchedGoEnv", "GoEnv", "GoExtraHardEnv", "GoVectorEnv"):
        from . import envs
        return getattr(envs, name)
    if name == "GoEngine":
        from .engine import GoEngine
        return GoEngine
    raise AttributeError(name)
