"""GoEnv - the reference's stateful single-board Gym environment (gym_go/envs/go_env.py:19-158) re-hosted on
the batched CUDA backend as a batch of one.  Same constructor kwargs, methods, return types and errors:
`done` is an int, `turn` an int, rewards/areas numpy float64, an illegal or post-game step raises
AssertionError, `state()` returns a fresh float64 [6,N,N] array.  GUI rendering (pyglet) is out of scope.

The board lives on the device as one packed record (`self._rec`); `self.state_` is the host mirror of the dense
[6,N,N] array the reference keeps, refreshed after every step, so the cheap whole-plane readers (turn, pass, done)
work on it exactly like the reference's."""
from enum import Enum

import numpy as np
import torch

from .. import _cabi, gogame, govars
from ..engine import GoEngine


def _find_gym():
    """gym or gymnasium when one of them is importable (neither is installed in the build image)"""
    for name in ("gym", "gymnasium"):
        try:
            return __import__(name)
        except Exception:               # noqa: BLE001 - optional dependency
            continue
    return None


_gym = _find_gym()
_EnvBase = object if _gym is None else _gym.Env


class RewardMethod(Enum):
    """go_env.py:9-16.  REAL: 0 while the game runs, then +1 / 0 / -1 for a black win / tie / white win.
    HEURISTIC: black area - white area - komi while the game runs, then +-N*N (a tie counts as -N*N)."""
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
            # go_env.py:35-37: a float32 box over the six planes, one discrete action per point plus the pass
            self.observation_space = _gym.spaces.Box(np.float32(0), np.float32(govars.NUM_CHNLS),
                                                     shape=(govars.NUM_CHNLS, size, size))
            self.action_space = _gym.spaces.Discrete(gogame.action_size(board_size=size))
        self.reset()

    # ------------------------------------------------------------------ episode control
    def reset(self):
        """go_env.py:40-47: empty board, black to move; returns a copy of the state"""
        self._rec = self._engine.new_records(1)
        self.state_ = gogame.init_state(self.size)
        self.done = False
        return self.state()

    def _flat_action(self, action):
        """the three action spellings of go_env.py:55-60 -> index in [0, N*N]: a (row, col) pair is checked against
        the board, None is the pass, anything else is already an index"""
        if action is None:
            return self.size * self.size
        if isinstance(action, (tuple, list, np.ndarray)):
            row, col = action[0], action[1]
            assert 0 <= row < self.size, ("row off the board", action)
            assert 0 <= col < self.size, ("column off the board", action)
            return int(row) * self.size + int(col)
        return int(action)

    def step(self, action):
        """go_env.py:49-64: one ply by the player to move -> (state, reward, done, info); AssertionError on a finished
        game or a refused move (occupied point, ko, suicide, index out of range) - the board is unchanged then"""
        assert not self.done, "step() on a finished game"
        index = self._flat_action(action)
        out = self._engine.step(self._rec, [index], out=self._rec, refuse_done=True)
        code = int(out["status"][0])
        assert code == _cabi.GG_ST_OK, ("Invalid move", action, code)
        self.state_ = self._engine.unpack(self._rec, dtype=torch.float64)[0].cpu().numpy()
        self.done = gogame.game_ended(self.state_)
        return self.state(), self.reward(), self.done, self.info()

    # ------------------------------------------------------------------ readers
    def state(self):
        return np.copy(self.state_)

    def canonical_state(self):
        flipped = self._engine.canonical(self._rec)
        return self._engine.unpack(flipped, dtype=torch.float64)[0].cpu().numpy()

    def game_ended(self):
        return self.done

    def turn(self):
        return gogame.turn(self.state_)

    def prev_player_passed(self):
        return gogame.prev_player_passed(self.state_)

    def valid_moves(self):
        """go_env.py:75-76: float64 [N*N+1], 1 = playable; all ones once the game is over (gogame.py:155-156)"""
        mask = self._engine.valid_moves(self._rec, ended_quirk=True, dtype=torch.float64)
        return mask[0].cpu().numpy()

    def uniform_random_action(self):
        """go_env.py:78-81: one numpy draw over the playable indices, pass included"""
        return np.random.choice(np.flatnonzero(self.valid_moves()))

    def info(self):
        """go_env.py:83-91"""
        return dict(turn=self.turn(), invalid_moves=1 - self.valid_moves(),
                    prev_player_passed=self.prev_player_passed())

    def children(self, canonical=False, padded=True):
        """go_env.py:105-109: the state after every playable action (zeros at the others when padded)"""
        out = self._engine.children(self._rec, canonical=canonical, obs_dtype=torch.uint8, want_rec=False)
        assert int(out["status"][0]) == 0, "Invalid move in children()"
        states = out["obs"][0]
        if not padded:
            states = states[out["valid"][0].bool()]
        return states.cpu().numpy().astype(np.float64)

    # ------------------------------------------------------------------ scoring
    def _margin(self):
        """black area - white area - komi of the current position (Tromp-Taylor areas from gg_areas)"""
        black, white = self._engine.areas(self._rec)[0].cpu().numpy().astype(np.float64)
        return black - white - self.komi

    def winning(self):
        """go_env.py:111-115: who is ahead right now from black's side, +1 / 0 / -1"""
        return np.sign(self._margin())

    def winner(self):
        """go_env.py:117-126: winning() once the game is over, 0 before"""
        return self.winning() if self.game_ended() else 0

    def reward(self):
        """go_env.py:128-149"""
        if self.reward_method is RewardMethod.REAL:
            return self.winner()
        if self.reward_method is RewardMethod.HEURISTIC:
            margin = self._margin()
            if not self.game_ended():
                return margin
            points = self.size ** 2
            return points if margin > 0 else -points        # a tie is scored like a loss (go_env.py:145-146)
        raise Exception("Unknown Reward Method")

    # ------------------------------------------------------------------ presentation
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
