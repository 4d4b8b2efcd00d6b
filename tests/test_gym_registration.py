"""`gym.make('gym_go:go-v0', ...)` - the first item of the north-star API (reference gym_go/__init__.py:1-10,
envs/go_env.py:35-37).  `gym` is not installed here, so each test runs in a fresh interpreter that installs the stub
from tests/gymstub.py before anything imports gym_go (a real gym would sit in the same place)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_isolated(body):
    code = "import sys\nsys.path[:0] = [%r, %r]\nimport gymstub\ngym = gymstub.install()\n" % (ROOT, os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-c", code + textwrap.dedent(body)], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, cwd=ROOT)
    assert p.returncode == 0, p.stdout
    return p.stdout


def test_importing_gym_go_registers_the_reference_ids():
    out = run_isolated("""
        import gym_go
        assert gym.registry["go-v0"] == "gymgo_b200.envs:GoEnv", gym.registry
        assert gym.registry["go-extrahard-v0"] == "gymgo_b200.envs:GoExtraHardEnv"
        import importlib
        importlib.reload(gym_go)                      # a second import must not trip over the existing ids
        from gym_go import gogame, govars
        assert govars.NUM_CHNLS == 6 and gogame.action_size(board_size=7) == 50
        from gym_go.envs import GoEnv
        assert issubclass(GoEnv, gym.Env)
        print("registered")
    """)
    assert "registered" in out


def test_make_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    out = run_isolated("""
        from gymgo_b200._cabi import GymGoB200Error
        try:
            gym.make('gym_go:go-v0', size=7)
        except GymGoB200Error as exc:
            print("refused:", exc)
        else:
            raise SystemExit("gym.make built an environment without CUDA")
    """)
    assert "refused" in out and "no CPU fallback" in out


@pytest.mark.gpu
def test_gym_make_returns_the_drop_in_env_and_plays_a_reference_sequence():
    out = run_isolated("""
        import numpy as np
        sys.path.insert(0, 'tests')
        import golden_io
        env = gym.make('gym_go:go-v0', size=7, komi=0, reward_method='real')
        from gym_go.envs import GoEnv
        assert type(env) is GoEnv and isinstance(env, gym.Env)
        assert env.observation_space.shape == (6, 7, 7) and env.observation_space.dtype == np.float32
        assert env.observation_space.low == 0 and env.observation_space.high == 6
        assert env.action_space.n == 50
        state = env.reset()
        assert env.observation_space.contains(state) and state.dtype == np.float64
        # the reference's ko test sequence (gym_go/tests/test_invalid_moves.py:43-83), from the golden fixture
        ko = [c for c in golden_io.kat_cases() if c["name"] == "ko"][0]
        for i, a in enumerate(ko["actions"]):
            state, reward, done, info = env.step(int(a))
            assert env.action_space.contains(int(a))
            assert np.array_equal(state, ko["states"][i + 1]) and float(reward) == ko["rewards"][i] and done == ko["dones"][i]
        try:
            env.step(int(ko["raises"]))
        except AssertionError:
            print("ko refused")
        hard = gym.make('gym_go:go-extrahard-v0', size=5)
        assert hard.action_space.n == 26
        # the local stand-in resolves the same ids without gym
        import gymgo_b200
        assert type(gymgo_b200.make('gym_go:go-v0', size=5)) is GoEnv
        print("played")
    """)
    assert "ko refused" in out and "played" in out
