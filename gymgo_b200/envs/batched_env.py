"""BatchedGoEnv - the object the headline metric measures: tens of thousands of independent boards in one
bit-packed device tensor, stepped by one kernel launch.  Semantics per board are GoEnv's
(gym_go/envs/go_env.py:24-149): same action encoding, rewards, done flags; observations are the 6xNxN
state tensors, float32 like GoEnv.observation_space (go_env.py:35-36) unless another dtype is asked."""
import numpy as np
import torch

from .. import _cabi
from ..engine import _TORCH2GG, GoEngine

_REWARD = {"real": _cabi.GG_REWARD_REAL, "heuristic": _cabi.GG_REWARD_HEURISTIC}


class BatchedGoEnv(object):
    def __init__(self, batch_size, size, komi=0, reward_method="real", device=None, obs_dtype=torch.float32,
                 strict=False, seed=0, board_offset=0, use_cuda_graph=False):
        """board_offset: global index of board 0 (so that rollouts are identical however the global batch is
        sharded over GPUs); strict: raise AssertionError when any board refuses its action;
        use_cuda_graph: step() replays a captured CUDA graph (reset + ply kernels) instead of enqueuing the
        launches from Python - the C ABI only enqueues work on the caller's stream, so it is capturable."""
        if reward_method not in _REWARD:
            raise ValueError("reward_method must be 'real' or 'heuristic'")
        self.engine = GoEngine(size, device)
        self.batch_size, self.size, self.komi = int(batch_size), int(size), komi
        self.reward_method, self.reward_mode = reward_method, _REWARD[reward_method]
        self.obs_dtype, self.strict, self.seed, self.board_offset = obs_dtype, strict, int(seed), int(board_offset)
        e = self.engine
        self.rec = e.new_records(self.batch_size)
        self.obs = e.empty((self.batch_size, 6, size, size), dtype=obs_dtype)
        self.reward = e.empty((self.batch_size,), dtype=torch.float32)
        self.done = e.empty((self.batch_size,))
        self.status = e.empty((self.batch_size,))
        self.actions = e.empty((self.batch_size,), dtype=torch.int32)
        self._step_actions = e.empty((self.batch_size,), dtype=torch.int32)    # static input of step() (graph-safe)
        self.use_cuda_graph, self._graphs, self._c_args = bool(use_cuda_graph), {}, None
        self.t = 0
        self.reset()

    # gym-style API ----------------------------------------------------------------------------
    def reset(self, mask=None):
        """all boards (or those with mask != 0) back to the empty position; returns observations"""
        if mask is not None:
            mask = torch.as_tensor(np.asarray(mask) if not isinstance(mask, torch.Tensor) else mask)
            mask = mask.to(self.done.device).ne(0)
        self.engine.reset(self.rec, mask)
        if mask is None:
            self.done.zero_()
        else:
            self.done.masked_fill_(mask, 0)
        return self.engine.unpack(self.rec, out=self.obs)

    @property
    def action_buffer(self):
        """the env's static int32 [B] action tensor: fill it (e.g. copy_ from pinned host memory) and pass it to
        step() to avoid an extra device copy"""
        return self._step_actions

    def _enqueue_step(self, auto_reset):
        """enqueue (optional reset of finished boards) + one ply on torch's current stream; inputs/outputs are the
        env's static tensors, so the argument lists are built once and the same sequence can be captured into a
        CUDA graph"""
        e = self.engine
        s = e._enter()
        if self._c_args is None:
            self._c_args = (
                (self.rec.data_ptr(), self.batch_size, self.size, self.done.data_ptr()),
                (self.rec.data_ptr(), self._step_actions.data_ptr(), self.rec.data_ptr(), self.status.data_ptr(),
                 self.batch_size, self.size, _cabi.GG_STEP_REFUSE_DONE, self.obs.data_ptr(),
                 _TORCH2GG[self.obs_dtype], self.done.data_ptr(), None,
                 self.reward.data_ptr(), self.reward_mode, float(self.komi)))
        reset_args, step_args = self._c_args
        if auto_reset:
            _cabi.check(e.lib.gg_reset(*reset_args, s))
        _cabi.check(e.lib.gg_step(*step_args, s))

    def _graph(self, auto_reset):
        g = self._graphs.get(auto_reset)
        if g is None:
            # make sure both kernels are loaded before capturing (lazy module loading): run them on a scratch board
            scratch = BatchedGoEnv(1, self.size, device=self.engine.device, obs_dtype=self.obs_dtype)
            scratch._step_actions.fill_(self.size * self.size)
            scratch._enqueue_step(True)
            torch.cuda.synchronize(self.engine.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue_step(auto_reset)
            self._graphs[auto_reset] = g
        return g

    def step(self, actions, auto_reset=False):
        """actions: int [B] (N*N = pass) -> (obs [B,6,N,N], reward [B] f32, done [B] u8, info).
        Finished boards refuse to step (status 3, GoEnv's `assert not self.done`) until reset;
        auto_reset=True first resets the boards that finished on the previous step (vector-env style)."""
        if actions is not self._step_actions:                       # (passing env.action_buffer skips all of this)
            a = self.engine._actions(actions, self.batch_size)
            if a.data_ptr() != self._step_actions.data_ptr():
                self._step_actions.copy_(a, non_blocking=True)
        if self.use_cuda_graph:
            self.engine._enter()
            self._graph(bool(auto_reset)).replay()
        else:
            self._enqueue_step(auto_reset)
        if self.strict and bool(self.status.any()):
            i = int(torch.nonzero(self.status)[0])
            raise AssertionError(("refused action", int(self._step_actions[i]), "board %d" % i,
                                  "status %d" % int(self.status[i])))
        return self.obs, self.reward, self.done, {"status": self.status}

    def random_step(self):
        """fused: auto-reset finished boards, draw a uniformly random legal action (incl. pass), play it.
        -> (obs, reward, done, actions)"""
        self.engine.rollout_step(self.rec, self.seed, self.board_offset, self.t, actions=self.actions, obs=self.obs,
                                 done=self.done, reward=self.reward, reward_mode=self.reward_mode, komi=self.komi)
        self.t += 1
        return self.obs, self.reward, self.done, self.actions

    # accessors ----------------------------------------------------------------------------------
    def state(self, dtype=None):
        return self.engine.unpack(self.rec, dtype=dtype or self.obs_dtype)

    def canonical_state(self, dtype=None):
        return self.engine.unpack(self.engine.canonical(self.rec), dtype=dtype or self.obs_dtype)

    def valid_moves(self, dtype=torch.float32):
        return self.engine.valid_moves(self.rec, ended_quirk=True, dtype=dtype)

    def invalid_moves(self, dtype=torch.float32):
        return 1 - self.valid_moves(dtype)

    def uniform_random_action(self):
        a = self.engine.sample_legal(self.rec, self.seed, self.board_offset, self.t)
        self.t += 1
        return a

    def children(self, canonical=False, obs_dtype=None, want_rec=False):
        return self.engine.children(self.rec, canonical=canonical, obs_dtype=obs_dtype or self.obs_dtype,
                                    want_rec=want_rec)

    def areas(self):
        return self.engine.areas(self.rec)

    def winning(self):
        ar = self.areas()
        return torch.sign(ar[:, 0].float() - ar[:, 1].float() - self.komi)

    def turn(self):
        return self.engine.flags(self.rec) & 1

    def prev_player_passed(self):
        return (self.engine.flags(self.rec) >> 1) & 1

    def game_ended(self):
        return (self.engine.flags(self.rec) >> 2) & 1

    def __len__(self):
        return self.batch_size
