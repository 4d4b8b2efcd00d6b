"""`gym_go` compatibility package: the reference's import paths (`import gym_go`, `from gym_go import gogame,
govars`, `gym.make('gym_go:go-v0')`) resolved to the B200-native implementation in gymgo_b200."""
import gymgo_b200
from gymgo_b200 import govars  # noqa: F401

gymgo_b200.register_with_gym()


def __getattr__(name):
    if name in ("gogame", "envs"):
        import importlib
        return importlib.import_module("gym_go." + name)
    raise AttributeError(name)
