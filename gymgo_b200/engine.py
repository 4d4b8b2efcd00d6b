"""Tensor-level host API over the C ABI: PyTorch owns the device memory and the stream, the kernels in
libgymgo_b200.so do the work.  `GoEngine` is stateless (records in, records out); `BatchedGoEnv`
(gymgo_b200/envs/batched_env.py) keeps one packed tensor of boards.

Everything here requires a CUDA device - there is no CPU execution path."""
import threading

import numpy as np
import torch

from . import _cabi

_TORCH2GG = {torch.uint8: _cabi.GG_U8, torch.float32: _cabi.GG_F32, torch.float64: _cabi.GG_F64,
             torch.bfloat16: _cabi.GG_BF16, torch.float16: _cabi.GG_F16}
_OBS_DTYPES = (torch.float32, torch.uint8, torch.bfloat16, torch.float16)
_tls = threading.local()


def _require_cuda():
    if not torch.cuda.is_available():
        raise _cabi.GymGoB200Error("gymgo_b200 needs a CUDA device (sm_100a kernels); no CPU fallback exists")


def _ptr(t):
    return None if t is None else t.data_ptr()


class GoEngine(object):
    """Kernels for boards of side `size` on one device."""

    def __init__(self, size, device=None):
        _require_cuda()
        self.lib = _cabi.lib()
        if not self.lib.gg_supported(int(size)):
            raise _cabi.GymGoB200Error("board size %r not supported (2..19)" % (size,))
        self.size = int(size)
        self.points = self.size * self.size
        self.actions = self.points + 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.GymGoB200Error("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.layout = _cabi.layout(self.size)
        self.rec_bytes = self.layout["rec_bytes"]

    # ------------------------------------------------------------------ plumbing
    def _enter(self):
        """Make this engine's device current and return torch's current stream on it.
        The library links its own CUDA runtime, but the "current device" is a property of the thread's driver
        context that both runtimes follow (measured: gg_set_device(1) moves torch.cuda.current_device() to 1), so,
        like torch.cuda.set_device, this switches the calling thread to the engine's device and leaves it there.
        One process per GPU (the intended deployment) never notices."""
        idx = self.device.index
        if torch.cuda.current_device() != idx:
            torch.cuda.set_device(idx)
            _tls.device = None
        if getattr(_tls, "device", None) != idx:
            _cabi.check(self.lib.gg_set_device(idx))
            _tls.device = idx
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check_rec(self, rec):
        if rec.dtype != torch.uint8 or rec.dim() != 2 or rec.shape[1] != self.rec_bytes or not rec.is_contiguous() \
                or rec.device != self.device:
            raise ValueError("records must be a contiguous uint8 [B, %d] tensor on %s" % (self.rec_bytes, self.device))

    def _actions(self, actions, batch):
        a = torch.as_tensor(actions)
        if a.dim() == 0:
            a = a.reshape(1)
        a = a.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
        if a.shape != (batch,):
            raise ValueError("actions must have shape [%d]" % batch)
        return a

    def _check_out(self, t, name, shape, dtypes):
        """optional output tensor: right device, contiguous, exact shape, allowed dtype (the C ABI trusts pointers)"""
        if t is None:
            return
        if not isinstance(dtypes, (tuple, list)):
            dtypes = (dtypes,)
        if t.device != self.device or not t.is_contiguous() or tuple(t.shape) != tuple(shape) or t.dtype not in dtypes:
            raise ValueError("%s must be a contiguous %s tensor of shape %s on %s (got %s %s on %s)"
                             % (name, "/".join(str(d) for d in dtypes), tuple(shape), self.device, t.dtype,
                                tuple(t.shape), t.device))

    def empty(self, *shape, dtype=torch.uint8):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ state type
    def new_records(self, batch):
        """`batch` empty boards, black to move (gogame.batch_init_state, gogame.py:28-31)."""
        return torch.zeros((int(batch), self.rec_bytes), dtype=torch.uint8, device=self.device)

    def pack(self, dense):
        """dense [B,6,N,N] (uint8/float32/float64, values 0/1) -> packed records."""
        d = torch.as_tensor(dense)
        if d.dtype not in _TORCH2GG:
            d = d.to(torch.float32)
        d = d.to(self.device).contiguous()
        if d.dim() != 4 or tuple(d.shape[1:]) != (6, self.size, self.size):
            raise ValueError("dense states must be [B,6,%d,%d]" % (self.size, self.size))
        rec = self.empty((d.shape[0], self.rec_bytes))
        s = self._enter()
        _cabi.check(self.lib.gg_pack(_ptr(d), _TORCH2GG[d.dtype], d.shape[0], self.size, _ptr(rec), s))
        return rec

    def unpack(self, rec, dtype=torch.float32, out=None):
        self._check_rec(rec)
        if out is None:
            out = self.empty((rec.shape[0], 6, self.size, self.size), dtype=dtype)
        self._check_out(out, "out", (rec.shape[0], 6, self.size, self.size), tuple(_TORCH2GG))
        s = self._enter()
        _cabi.check(self.lib.gg_unpack(_ptr(rec), rec.shape[0], self.size, _TORCH2GG[out.dtype], _ptr(out), s))
        return out

    def reset(self, rec, mask=None):
        self._check_rec(rec)
        if mask is not None:
            mask = torch.as_tensor(mask)
            if tuple(mask.shape) != (rec.shape[0],):
                raise ValueError("mask must have shape [%d] (one entry per board), got %s" % (rec.shape[0], tuple(mask.shape)))
            mask = mask.ne(0).to(device=self.device, dtype=torch.uint8).contiguous()
        s = self._enter()
        _cabi.check(self.lib.gg_reset(_ptr(rec), rec.shape[0], self.size, _ptr(mask), s))
        return rec

    # ------------------------------------------------------------------ the hot path
    def step(self, rec, actions, out=None, canonical=False, refuse_done=False, obs=None, obs_dtype=None,
             want_status=True, want_done=False, want_areas=False, reward_mode=0, komi=0.0, auto_reset=False,
             kernel=None):
        """One ply per board.  Returns dict(rec, status, obs, done, areas, reward) of device tensors
        (entries not asked for are None).  `out=rec` steps in place.  auto_reset: boards whose record is finished
        restart from the empty position before playing their action (GG_STEP_AUTO_RESET); auto_reset="skip" also
        ignores their action for this ply (GG_STEP_RESET_SKIPS_ACTION).  kernel: None (chosen by the library),
        "lanes" or "thread" (GG_STEP_KERNEL_*; identical results)."""
        self._check_rec(rec)
        b = rec.shape[0]
        a = self._actions(actions, b)
        if out is None:
            out = torch.empty_like(rec)
        else:
            self._check_rec(out)
            if out.shape[0] != b:
                raise ValueError("out must hold %d records" % b)
        if obs is None and obs_dtype is not None:
            obs = self.empty((b, 6, self.size, self.size), dtype=obs_dtype)
        self._check_out(obs, "obs", (b, 6, self.size, self.size), _OBS_DTYPES)
        status = self.empty((b,)) if want_status else None
        done = self.empty((b,)) if want_done else None
        areas = self.empty((b, 2), dtype=torch.int32) if want_areas else None
        reward = self.empty((b,), dtype=torch.float32) if reward_mode else None
        flags = (_cabi.GG_STEP_CANONICAL if canonical else 0) | (_cabi.GG_STEP_REFUSE_DONE if refuse_done else 0) \
            | (_cabi.GG_STEP_AUTO_RESET if auto_reset else 0) \
            | (_cabi.GG_STEP_RESET_SKIPS_ACTION if auto_reset == "skip" else 0) \
            | {None: 0, "lanes": _cabi.GG_STEP_KERNEL_LANES, "thread": _cabi.GG_STEP_KERNEL_THREAD}[kernel]
        s = self._enter()
        _cabi.check(self.lib.gg_step(_ptr(rec), _ptr(a), _ptr(out), _ptr(status), b, self.size, flags, _ptr(obs),
                                     _TORCH2GG[obs.dtype] if obs is not None else 0, _ptr(done), _ptr(areas),
                                     _ptr(reward), int(reward_mode), float(komi), s))
        return dict(rec=out, status=status, obs=obs, done=done, areas=areas, reward=reward)

    def rollout_step(self, rec, seed, board0, t, actions=None, obs=None, done=None, areas=None, reward=None,
                     reward_mode=0, komi=0.0):
        """Fused auto-reset + uniform-random-legal action + ply, in place on `rec`.  All outputs are optional
        preallocated tensors (nothing is allocated here: this is the benchmark loop)."""
        self._check_rec(rec)
        b = rec.shape[0]
        self._check_out(actions, "actions", (b,), torch.int32)
        self._check_out(obs, "obs", (b, 6, self.size, self.size), _OBS_DTYPES)
        self._check_out(done, "done", (b,), torch.uint8)
        self._check_out(areas, "areas", (b, 2), torch.int32)
        self._check_out(reward, "reward", (b,), torch.float32)
        s = self._enter()
        _cabi.check(self.lib.gg_rollout_step(_ptr(rec), rec.shape[0], self.size, int(seed), int(board0), int(t),
                                             _ptr(actions), _ptr(obs), _TORCH2GG[obs.dtype] if obs is not None else 0,
                                             _ptr(done), _ptr(areas), _ptr(reward), int(reward_mode), float(komi), s))

    def rollout(self, rec, seed, board0, t0, steps, plies_per_launch=32, actions_log=None, obs_ring=None,
                done_log=None, reward_log=None, reward_mode=0, komi=0.0, kernel=_cabi.GG_KERNEL_AUTO, dynamic=True, block_plies=0):
        """`steps` fused rollout plies by the persistent kernel (gg_rollout), `plies_per_launch` plies per launch.
        obs_ring: [R,B,6,N,N] ring of observation slots (ply t writes slot t % R); actions_log int32 [steps,B],
        done_log uint8 [steps,B], reward_log float32 [steps,B] - all optional, preallocated (logs may be longer
        than `steps`: only the first `steps` rows are written).  kernel: GG_KERNEL_* (AUTO = the measured choice; the
        kernels are bit-identical, the explicit values exist for A/B measurements and parity tests).  dynamic: give the
        library a scheduling workspace so that long launches are load-balanced over the SMs (same results);
        block_plies: plies per scheduling block (0 = the library's measured default)."""
        self._check_rec(rec)
        b, steps = rec.shape[0], int(steps)
        for t, name, dt in ((actions_log, "actions_log", torch.int32), (done_log, "done_log", torch.uint8),
                            (reward_log, "reward_log", torch.float32)):
            if t is not None and (t.dim() != 2 or t.shape[0] < steps):
                raise ValueError("%s needs at least %d rows" % (name, steps))
            self._check_out(t, name, (t.shape[0], b) if t is not None else None, dt)
        if obs_ring is not None:
            self._check_out(obs_ring, "obs_ring", (obs_ring.shape[0], b, 6, self.size, self.size), _OBS_DTYPES)
        s = self._enter()
        ring = 0 if obs_ring is None else int(obs_ring.shape[0])
        ws = self._workspace(b) if dynamic else None
        _cabi.check(self.lib.gg_rollout_with(int(kernel), _ptr(rec), rec.shape[0], self.size, int(seed), int(board0),
                                             int(t0), int(steps), int(plies_per_launch), _ptr(actions_log), _ptr(obs_ring),
                                             _TORCH2GG[obs_ring.dtype] if obs_ring is not None else 0, ring,
                                             _ptr(done_log), _ptr(reward_log), int(reward_mode), float(komi),
                                             _ptr(ws), 0 if ws is None else ws.numel(), int(block_plies), s))

    def _workspace(self, batch):
        """scheduling workspace of gg_rollout_with for `batch` boards.  Allocated per call from torch's stream-aware
        caching allocator (microseconds): two rollouts of one engine on different streams must not share tickets, and
        a block freed while its kernel is still queued is only ever reused in stream order."""
        return self.empty((int(self.lib.gg_rollout_workspace_bytes(self.size, batch)),))

    def sample_legal(self, rec, seed, board0, t):
        self._check_rec(rec)
        out = self.empty((rec.shape[0],), dtype=torch.int32)
        s = self._enter()
        _cabi.check(self.lib.gg_sample_legal(_ptr(rec), rec.shape[0], self.size, int(seed), int(board0), int(t),
                                             _ptr(out), s))
        return out

    # ------------------------------------------------------------------ derived quantities
    def valid_moves(self, rec, ended_quirk=False, dtype=torch.float32):
        self._check_rec(rec)
        out = self.empty((rec.shape[0], self.actions), dtype=dtype)
        s = self._enter()
        _cabi.check(self.lib.gg_valid_moves(_ptr(rec), rec.shape[0], self.size, int(bool(ended_quirk)),
                                            _TORCH2GG[dtype], _ptr(out), s))
        return out

    def children(self, rec, canonical=False, obs_dtype=torch.float32, want_rec=True, want_obs=True):
        """-> dict(rec [B,A,rec_bytes], obs [B,A,6,N,N], valid [B,A] uint8, status [B] uint8)."""
        self._check_rec(rec)
        b = rec.shape[0]
        crec = self.empty((b, self.actions, self.rec_bytes)) if want_rec else None
        cobs = self.empty((b, self.actions, 6, self.size, self.size), dtype=obs_dtype) if want_obs else None
        valid = self.empty((b, self.actions))
        status = self.empty((b,))
        s = self._enter()
        _cabi.check(self.lib.gg_children(_ptr(rec), b, self.size, _cabi.GG_STEP_CANONICAL if canonical else 0,
                                         _ptr(crec), _ptr(cobs), _TORCH2GG[obs_dtype] if want_obs else 0,
                                         _ptr(valid), _ptr(status), s))
        return dict(rec=crec, obs=cobs, valid=valid, status=status)

    def update_pieces(self, rec, touch, player):
        """Capture removal alone (state_utils.update_pieces, state_utils.py:159-180), in place on `rec`: groups of colour
        1 - player[b] touching a point of `touch` (records whose BLACK plane carries the adjacent locations) without a
        liberty are removed.  -> records whose BLACK plane holds the removed stones."""
        self._check_rec(rec)
        self._check_rec(touch)
        b = rec.shape[0]
        if touch.shape[0] != b:
            raise ValueError("touch must hold %d records" % b)
        player = torch.as_tensor(player).reshape(-1).to(device=self.device, dtype=torch.int32).contiguous()
        if player.shape != (b,):
            raise ValueError("player must have shape [%d]" % b)
        killed = torch.empty_like(rec)
        s = self._enter()
        _cabi.check(self.lib.gg_update_pieces(_ptr(rec), _ptr(touch), _ptr(player), _ptr(killed), b, self.size, s))
        return killed

    def areas(self, rec):
        self._check_rec(rec)
        out = self.empty((rec.shape[0], 2), dtype=torch.int32)
        s = self._enter()
        _cabi.check(self.lib.gg_areas(_ptr(rec), rec.shape[0], self.size, _ptr(out), s))
        return out

    def canonical(self, rec, out=None):
        self._check_rec(rec)
        if out is None:
            out = torch.empty_like(rec)
        else:
            self._check_rec(out)
            if out.shape[0] != rec.shape[0]:
                raise ValueError("out must hold %d records" % rec.shape[0])
        s = self._enter()
        _cabi.check(self.lib.gg_canonical(_ptr(rec), _ptr(out), rec.shape[0], self.size, s))
        return out

    def symmetry(self, rec, sym):
        """packed records transformed by dihedral symmetry `sym` (0..7, the order of gogame.all_symmetries)"""
        self._check_rec(rec)
        out = torch.empty_like(rec)
        s = self._enter()
        _cabi.check(self.lib.gg_symmetry(_ptr(rec), _ptr(out), rec.shape[0], self.size, int(sym), s))
        return out

    # flags word of every record as int32 [B]: bit0 turn, bit1 previous pass, bit2 game over
    def flags(self, rec):
        self._check_rec(rec)
        off = 3 * self.layout["lanes_per_board"] * (self.layout["word_bits"] // 8)
        return rec[:, off:off + 4].contiguous().view(torch.int32).reshape(-1)


_ENGINES = {}


def engine(size, device=None):
    """cached GoEngine per (size, device)"""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (int(size), dev.index)
    if key not in _ENGINES:
        _ENGINES[key] = GoEngine(size, dev)
    return _ENGINES[key]


def to_numpy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
