"""ctypes binding of the plain-C oracle (oracle/go_oracle.c).  TEST INFRASTRUCTURE ONLY - see the
header of go_oracle.c for who may use it.  Builds itself with gcc on first use."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgo_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "go_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libgo_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        u8p, i32p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32)
        _lib.go_next_state.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p]
        _lib.go_next_state.restype = ctypes.c_int
        _lib.go_batch_next_states.argtypes = [u8p, i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, u8p]
        _lib.go_batch_next_states.restype = None
        _lib.go_valid_moves.argtypes = [u8p, ctypes.c_int, ctypes.c_int, u8p]
        _lib.go_areas.argtypes = [u8p, ctypes.c_int, i32p, i32p]
        _lib.go_batch_areas.argtypes = [u8p, ctypes.c_int, ctypes.c_int, i32p]
        _lib.go_children.argtypes = [u8p, ctypes.c_int, ctypes.c_int, u8p, u8p]
        _lib.go_children.restype = ctypes.c_int
        _lib.go_invalid_mask.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p]
        _lib.go_rollout.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_uint64]
        _lib.go_rollout.restype = ctypes.c_uint64
    return _lib


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a, t=ctypes.c_uint8):
    return a.ctypes.data_as(ctypes.POINTER(t))


def batch_next_states(states, actions, canonical=False):
    """states [B,6,N,N] (any 0/1 dtype), actions [B] -> (next uint8 [B,6,N,N], status uint8 [B])."""
    s = _u8(states)
    a = np.ascontiguousarray(actions, dtype=np.int32)
    out = np.empty_like(s)
    status = np.zeros(len(s), dtype=np.uint8)
    if len(s):
        lib().go_batch_next_states(_p(s), _p(a, ctypes.c_int32), len(s), s.shape[2], int(canonical), _p(out), _p(status))
    return out, status


def next_state(state, action, canonical=False):
    out, status = batch_next_states(np.asarray(state)[None], [action], canonical)
    if status[0]:
        raise AssertionError(("Invalid move", int(action), int(status[0])))
    return out[0]


def valid_moves(state, ended_quirk=True):
    s = _u8(state)
    n = s.shape[1]
    out = np.zeros(n * n + 1, dtype=np.uint8)
    lib().go_valid_moves(_p(s), n, int(ended_quirk), _p(out))
    return out


def batch_areas(states):
    s = _u8(states)
    out = np.zeros((len(s), 2), dtype=np.int32)
    if len(s):
        lib().go_batch_areas(_p(s), len(s), s.shape[2], _p(out, ctypes.c_int32))
    return out


def areas(state):
    return tuple(int(x) for x in batch_areas(np.asarray(state)[None])[0])


def children(state, canonical=False):
    s = _u8(state)
    n = s.shape[1]
    out = np.zeros((n * n + 1,) + s.shape, dtype=np.uint8)
    valid = np.zeros(n * n + 1, dtype=np.uint8)
    bad = lib().go_children(_p(s), n, int(canonical), _p(out), _p(valid))
    return out, valid, bad


def invalid_mask(black, white, next_player, ko=-1):
    b, w = _u8(black), _u8(white)
    out = np.zeros_like(b)
    lib().go_invalid_mask(_p(b), _p(w), b.shape[0], int(next_player), int(ko), _p(out))
    return out


def rollout(n, steps, seed=0):
    return int(lib().go_rollout(n, int(steps), int(seed)))
