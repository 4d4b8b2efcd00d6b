"""BatchedGoEnv - the object the headline metric measures: tens of thousands of independent boards in one
bit-packed device tensor, stepped by one kernel launch.  Semantics per board are GoEnv's
(gym_go/envs/go_env.py:24-149): same action encoding, rewards, done flags; observations are the 6xNxN
state tensors, float32 like GoEnv.observation_space (go_env.py:35-36) unless another dtype is asked."""
import numpy as np
import torch

from .. import _cabi, hostmem
from ..engine import _TORCH2GG, GoEngine

_REWARD = {"real": _cabi.GG_REWARD_REAL, "heuristic": _cabi.GG_REWARD_HEURISTIC}


class BatchedGoEnv(object):
    def __init__(self, batch_size, size, komi=0, reward_method="real", device=None, obs_dtype=torch.float32,
                 strict=False, seed=0, board_offset=0, use_cuda_graph=False):
        """board_offset: global index of board 0 (so that rollouts are identical however the global batch is
        sharded over GPUs); strict: raise AssertionError when any board refuses its action;
        use_cuda_graph: step() replays a captured CUDA graph instead of enqueuing the launch from Python - the C ABI
        only enqueues work on the caller's stream, so it is capturable."""
        if reward_method not in _REWARD:
            raise ValueError("reward_method must be 'real' or 'heuristic'")
        self.engine = GoEngine(size, device)
        self.batch_size, self.size, self.komi = int(batch_size), int(size), komi
        self.reward_method, self.reward_mode = reward_method, _REWARD[reward_method]
        self.obs_dtype, self.strict, self.seed, self.board_offset = obs_dtype, strict, int(seed), int(board_offset)
        e = self.engine
        self.rec = e.new_records(self.batch_size)
        self.obs = e.empty((self.batch_size, 6, size, size), dtype=obs_dtype)
        # reward (f32) and done (u8) live back to back in one buffer so that a host-facing loop fetches both with a
        # single device->host copy (HostStepper below)
        self._tail = e.empty((5 * self.batch_size,))
        self.reward = self._tail[:4 * self.batch_size].view(torch.float32)
        self.done = self._tail[4 * self.batch_size:]
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

    def _enqueue_step(self, auto_reset, want_obs=True):
        """enqueue one ply (finished boards restart first when auto_reset) on torch's current stream: ONE kernel launch
        (GG_STEP_AUTO_RESET does the reset inside gg_step).  Inputs/outputs are the env's static tensors, so the
        argument tuples are built once and the launch can be captured into a CUDA graph."""
        e = self.engine
        s = e._enter()
        if self._c_args is None:
            def args(flags, obs):
                return (self.rec.data_ptr(), self._step_actions.data_ptr(), self.rec.data_ptr(), self.status.data_ptr(),
                        self.batch_size, self.size, flags, self.obs.data_ptr() if obs else None, _TORCH2GG[self.obs_dtype],
                        self.done.data_ptr(), None, self.reward.data_ptr(), self.reward_mode, float(self.komi))
            modes = {False: _cabi.GG_STEP_REFUSE_DONE, True: _cabi.GG_STEP_REFUSE_DONE | _cabi.GG_STEP_AUTO_RESET,
                     "skip": _cabi.GG_STEP_REFUSE_DONE | _cabi.GG_STEP_AUTO_RESET | _cabi.GG_STEP_RESET_SKIPS_ACTION}
            self._c_args = {(k, obs): args(f, obs) for k, f in modes.items() for obs in (True, False)}
            self._gg_step = e.lib.gg_step
        rc = self._gg_step(*self._c_args[(auto_reset if auto_reset == "skip" else bool(auto_reset), bool(want_obs))], s)
        if rc:
            _cabi.check(rc)

    def _graph(self, auto_reset):
        g = self._graphs.get(auto_reset)
        if g is None:
            # make sure the kernel is loaded before capturing (lazy module loading): run it on a scratch board
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
        auto_reset=True first resets the boards that finished on the previous step and plays their action on the fresh
        board; auto_reset="skip" resets them and ignores their action for this step (gymnasium next-step autoreset).
        The returned tensors are the env's static buffers: the next step() overwrites them - clone() what you keep."""
        if actions is not self._step_actions:                       # (passing env.action_buffer skips all of this)
            a = self.engine._actions(actions, self.batch_size)
            if a.data_ptr() != self._step_actions.data_ptr():
                self._step_actions.copy_(a, non_blocking=True)
        if self.use_cuda_graph:
            self.engine._enter()
            self._graph(auto_reset if auto_reset == "skip" else bool(auto_reset)).replay()
        else:
            self._enqueue_step(auto_reset)
        if self.strict and bool(self.status.any()):
            i = int(torch.nonzero(self.status)[0])
            raise AssertionError(("refused action", int(self._step_actions[i]), "board %d" % i,
                                  "status %d" % int(self.status[i])))
        return self.obs, self.reward, self.done, {"status": self.status}

    def host_stepper(self, returns="obs", auto_reset=True, use_cuda_graph=True, follow_current_stream=True,
                     transport="auto", threads=0):
        """-> HostStepper: the step with HOST buffers (pinned actions in, pinned results out); see the class"""
        return HostStepper(self, returns=returns, auto_reset=auto_reset, use_cuda_graph=use_cuda_graph,
                           follow_current_stream=follow_current_stream, transport=transport, threads=threads)

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


class HostStepper(object):
    """BatchedGoEnv.step for a consumer that lives in host memory: every call moves this ply's actions host->device,
    plays the ply (finished boards restart first when auto_reset), moves the results device->host and waits for
    them.  All host buffers are pinned and allocated once, preferably on the GPU's NUMA node (gymgo_b200.hostmem);
    the whole sequence - copy in, ONE kernel, copies out - is a CUDA graph, so a step costs one graph launch.

        hs = env.host_stepper(returns="obs")
        hs.actions[:] = ...            # int32 [B] pinned; N*N = pass
        obs, reward, done = hs.step()  # pinned host tensors, valid until the next step()

    returns: "obs"    the dense [B,6,N,N] observation in env.obs_dtype (what GoEnv.step returns, go_env.py:64)
             "packed" the packed records [B, rec_bytes] (40x fewer bytes than f32 on 9x9); hs.expand() unpacks them on
                      the host (gg_host_unpack, multi-threaded C++)
             "none"   reward and done only (the observation is consumed on the device)
    reward (f32 [B]) and done (u8 [B]) always come back, in one copy.
    transport (returns="obs" only): how the observation reaches host memory -
             "dense"  the kernel writes the dense tensor on the device and it crosses PCIe (bus-bound: 127.7 MB per
                      9x9 x 65,536 f32 step);
             "packed" the kernel writes NO dense tensor: the packed records cross PCIe (40x fewer bytes) and the same
                      [B,6,N,N] tensor is expanded next to the CPU by gg_host_unpack (AVX-512 mask moves, streaming
                      stores, `threads` workers - 0: this process's share of the usable cores) inside step().
             "auto"   (default) "packed" when the host codec has its AVX-512 path and at least 4 worker threads, else
                      "dense".  Measured on the B200 boxes (profiles/r02_host_codec_probe.json): 16 threads expand at
                      180 GB/s of dense output, 4 threads at 53 GB/s, the dense tensor crosses PCIe at 55 GB/s on one
                      GPU and 12 GB/s per GPU when eight GPUs share the host.
             All return bit-identical tensors (with "packed" the device-side `env.obs` is not refreshed: the consumer is
             on the host; `env.state()` unpacks on the device when needed).  The codec's workers want the host cores to themselves: a caller whose
             own CPU work leaves spinning OpenMP threads behind (torch's intra-op pool after a small CPU op) should
             set OMP_WAIT_POLICY=passive or torch.set_num_threads(1).
    The step runs on the stepper's own stream; with follow_current_stream (default) it first waits for the work already
    enqueued on torch's current stream (e.g. an env.reset()), so mixing it with the device API is safe."""

    def __init__(self, env, returns="obs", auto_reset=True, use_cuda_graph=True, follow_current_stream=True,
                 transport="auto", threads=0):
        if returns not in ("obs", "packed", "none"):
            raise ValueError("returns must be 'obs', 'packed' or 'none'")
        if transport not in ("auto", "dense", "packed"):
            raise ValueError("transport must be 'auto', 'dense' or 'packed'")
        self.env, self.returns, self.auto_reset = env, returns, bool(auto_reset)
        self.threads = int(threads) if threads and int(threads) > 0 else hostmem.codec_threads()
        if transport == "auto":
            fast_codec = env.engine.lib.gg_host_unpack_path().startswith(b"avx512")
            transport = "packed" if (fast_codec and self.threads >= 4) else "dense"
        self.transport = transport if returns == "obs" else ("packed" if returns == "packed" else "dense")
        dev = env.engine.device.index
        b = env.batch_size
        self.actions = hostmem.pinned_empty((b,), torch.int32, dev)
        self.actions.fill_(env.size * env.size)
        self._tail = hostmem.pinned_empty((5 * b,), torch.uint8, dev)
        self.reward = self._tail[:4 * b].view(torch.float32)
        self.done = self._tail[4 * b:]
        self.obs = hostmem.pinned_empty(tuple(env.obs.shape), env.obs_dtype, dev) if returns == "obs" else None
        self.rec = hostmem.pinned_empty(tuple(env.rec.shape), torch.uint8, dev) if self.transport == "packed" else None
        self._obs_over_pcie = self.obs is not None and self.transport == "dense"
        self.h2d_bytes = self.actions.numel() * 4
        self.d2h_bytes = self._tail.numel() + (self.obs.numel() * self.obs.element_size() if self._obs_over_pcie else 0) \
            + (self.rec.numel() if self.rec is not None else 0)
        self.host_expanded_bytes = self.obs.numel() * self.obs.element_size() if (self.obs is not None
                                                                                   and not self._obs_over_pcie) else 0
        self._gg_dtype = _TORCH2GG[env.obs_dtype]
        self._stream = torch.cuda.Stream(device=env.engine.device)
        self.follow_current_stream = bool(follow_current_stream)
        self._graph = None
        self.use_cuda_graph = bool(use_cuda_graph)
        self.placement = hostmem.describe(dev)

    def _enqueue(self):
        env = self.env
        env._step_actions.copy_(self.actions, non_blocking=True)
        env._enqueue_step(self.auto_reset, want_obs=not self.host_expanded_bytes)
        if self._obs_over_pcie:
            self.obs.copy_(env.obs, non_blocking=True)
        if self.rec is not None:
            self.rec.copy_(env.rec, non_blocking=True)
        self._tail.copy_(env._tail, non_blocking=True)

    def step(self):
        env = self.env
        env.engine._enter()
        if self.follow_current_stream:
            self._stream.wait_stream(torch.cuda.current_stream(env.engine.device))
        with torch.cuda.stream(self._stream):
            if self.use_cuda_graph:
                if self._graph is None:
                    scratch = BatchedGoEnv(1, env.size, device=env.engine.device, obs_dtype=env.obs_dtype)
                    scratch._step_actions.fill_(env.size * env.size)
                    scratch._enqueue_step(True)                      # load the kernel before capturing
                    torch.cuda.synchronize(env.engine.device)
                    self._graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self._graph, stream=self._stream):
                        self._enqueue()
                    # capture does not execute: fall through to the first replay
                self._graph.replay()
            else:
                self._enqueue()
        self._stream.synchronize()
        if self.host_expanded_bytes:                                 # records arrived: expand them next to the CPU
            rc = env.engine.lib.gg_host_unpack(self.rec.data_ptr(), env.batch_size, env.size, self._gg_dtype,
                                               self.obs.data_ptr(), self.threads)
            if rc:
                _cabi.check(rc)
        first = self.obs if self.returns == "obs" else (self.rec if self.returns == "packed" else None)
        return first, self.reward, self.done

    def expand(self, out=None, dtype=torch.float32, threads=0):
        """packed records of the last step -> dense [B,6,N,N] on the HOST (gg_host_unpack; threads=0: all usable cores)"""
        if self.rec is None:
            raise ValueError("expand() needs returns='packed' or transport='packed'")
        env = self.env
        if out is None:
            out = torch.empty((env.batch_size, 6, env.size, env.size), dtype=dtype)
        if out.device.type != "cpu" or not out.is_contiguous() or tuple(out.shape) != (env.batch_size, 6, env.size, env.size):
            raise ValueError("out must be a contiguous host tensor of shape [B,6,N,N]")
        _cabi.check(env.engine.lib.gg_host_unpack(self.rec.data_ptr(), env.batch_size, env.size, _TORCH2GG[out.dtype],
                                                  out.data_ptr(), int(threads) if threads and threads > 0 else self.threads))
        return out
