"""GoEnv - the reference's stateful single-board Gym environment (gym_go/envs/go_env.py:19-158) re-hosted on
the batched CUDA backend as a batch of one.  Same constructor kwargs, methods, return types and errors:
`done` is an int, `turn` an int, rewards/areas numpy float64, an illegal or post-game step raises
AssertionError, `state()` returns a fresh float64 [6,N,N] array.  GUI rendering (pyglet) is out of scope."""
from enum import Enum

import numpy as np
import torch

from .. import _cabi, gogame, govars
from ..engine import GoEngine

try:                                    # gym / gymnasium are optional (not installed in this image)
    import gym as _gym
    _EnvBase = _gym.Env
except Exception:                       # noqa: BLE001
    try:
        import gymnasium as _gym
        _EnvBase = _gym.Env
    except Exception:                   # noqa: BLE001
        _gym, _EnvBase = None, object


class RewardMethod(Enum):
    """go_env.py:9-16"""
    REAL = 'real'
    HEURISTIC = 'heuristic'


class GoEnv(_EnvBase):
    metadata = {'render.modes': ['terminal']}
    govars = govars
    gogame = gogame

    def __init__(self, size, komi=0, reward_method='real', device=None):
        self.size = size
        self.komi = komi
        self.reward_method = RewardMethod(reward_method)
        self._engine = GoEngine(size, device)
        if _gym is not None:
            self.observation_space = _gym.spaces.Box(np.float32(0), np.float32(govars.NUM_CHNLS),
                                                     shape=(govars.NUM_CHNLS, size, size))
            self.action_space = _gym.spaces.Discrete(size * size + 1)
        self.reset()

    # -- internal: device record <-> host copy of the dense state
    def _sync_host(self):
        self.state_ = self._engine.unpack(self._rec, dtype=torch.float64)[0].cpu().numpy()

    def reset(self):
        """go_env.py:40-47"""
        self._rec = self._engine.new_records(1)
        self.state_ = gogame.init_state(self.size)
        self.done = False
        return np.copy(self.state_)

    def step(self, action):
        """go_env.py:49-64"""
        assert not self.done
        if isinstance(action, tuple) or isinstance(action, list) or isinstance(action, np.ndarray):
            assert 0 <= action[0] < self.size
            assert 0 <= action[1] < self.size
            action = self.size * action[0] + action[1]
        elif action is None:
            action = self.size ** 2
        res = self._engine.step(self._rec, [int(action)], out=self._rec, refuse_done=True)
        status = int(res["status"][0])
        assert status == _cabi.GG_ST_OK, ("Invalid move", action, status)
        self._sync_host()
        self.done = gogame.game_ended(self.state_)
        return np.copy(self.state_), self.reward(), self.done, self.info()

    def game_ended(self):
        return self.done

    def turn(self):
        return gogame.turn(self.state_)

    def prev_player_passed(self):
        return gogame.prev_player_passed(self.state_)

    def valid_moves(self):
        """go_env.py:75-76 (all ones once the game is over, gogame.py:155-156)"""
        return self._engine.valid_moves(self._rec, ended_quirk=True, dtype=torch.float64)[0].cpu().numpy()

    def uniform_random_action(self):
        """go_env.py:78-81"""
        valid_move_idcs = np.argwhere(self.valid_moves()).flatten()
        return np.random.choice(valid_move_idcs)

    def info(self):
        """go_env.py:83-91"""
        return {
            'turn': gogame.turn(self.state_),
            'invalid_moves': 1 - self.valid_moves(),
            'prev_player_passed': gogame.prev_player_passed(self.state_),
        }

    def state(self):
        return np.copy(self.state_)

    def canonical_state(self):
        return self._engine.unpack(self._engine.canonical(self._rec), dtype=torch.float64)[0].cpu().numpy()

    def children(self, canonical=False, padded=True):
        """go_env.py:105-109"""
        res = self._engine.children(self._rec, canonical=canonical, obs_dtype=torch.uint8, want_rec=False)
        assert int(res["status"][0]) == 0, "Invalid move in children()"
        kids = res["obs"][0]
        if not padded:
            kids = kids[res["valid"][0].bool()]
        return kids.cpu().numpy().astype(np.float64)

    def _areas(self):
        ar = self._engine.areas(self._rec)[0].cpu().numpy().astype(np.float64)
        return ar[0], ar[1]

    def winning(self):
        """go_env.py:111-115"""
        black_area, white_area = self._areas()
        return np.sign(black_area - white_area - self.komi)

    def winner(self):
        """go_env.py:117-126"""
        if self.game_ended():
            return self.winning()
        return 0

    def reward(self):
        """go_env.py:128-149"""
        if self.reward_method == RewardMethod.REAL:
            return self.winner()
        elif self.reward_method == RewardMethod.HEURISTIC:
            black_area, white_area = self._areas()
            komi_correction = black_area - white_area - self.komi
            if self.game_ended():
                return (1 if komi_correction > 0 else -1) * self.size ** 2
            return komi_correction
        raise Exception("Unknown Reward Method")

    def __str__(self):
        return gogame.str(self.state_)

    def close(self):
        pass

    def render(self, mode='terminal'):
        if mode != 'terminal':
            raise NotImplementedError("only the terminal renderer is provided (pyglet GUI is out of scope)")
        print(self.__str__())


class GoExtraHardEnv(GoEnv):
    """gym_go/envs/go_extrahard_env.py:1-5 - an empty subclass in the reference too."""
