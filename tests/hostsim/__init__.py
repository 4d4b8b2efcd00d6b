"""ctypes binding of the host simulator of the device algorithm (tests/hostsim/hostsim.cpp).
TEST INFRASTRUCTURE ONLY: it lets the CPU suite exercise gymgo_b200/csrc/gg_algo.cuh - the code the
CUDA kernels are built from - against the oracle without a GPU.  Not a fallback for anything."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgg_hostsim.so")
_ALGO = os.path.join(_HERE, "..", "..", "gymgo_b200", "csrc", "gg_algo.cuh")
_lib = None


def build():
    src = os.path.join(_HERE, "hostsim.cpp")
    newest = max(os.path.getmtime(src), os.path.getmtime(_ALGO))
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-Wall", "-x", "c++", src, "-o", _SO])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hs_philox.restype = ctypes.c_uint32
        _lib.hs_philox.argtypes = [ctypes.c_uint64] * 3
    return _lib


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def layout(n):
    out = np.zeros(5, dtype=np.int32)
    assert lib().hs_layout(n, _ptr(out)) == 0, "unsupported size"
    return dict(rec_bytes=int(out[0]), lpb=int(out[1]), rpl=int(out[2]), wordbits=int(out[3]), bpw=int(out[4]))


def pack(dense):
    d = np.ascontiguousarray(dense, dtype=np.uint8)
    n = d.shape[2]
    recs = np.zeros((len(d), layout(n)["rec_bytes"] // 4), dtype=np.uint32)
    assert lib().hs_pack(n, _ptr(d), len(d), _ptr(recs)) == 0
    return recs


def unpack(recs, n):
    r = np.ascontiguousarray(recs, dtype=np.uint32)
    out = np.zeros((len(r), 6, n, n), dtype=np.uint8)
    assert lib().hs_unpack(n, _ptr(r), len(r), _ptr(out)) == 0
    return out


def step(recs, actions, n, opts=0):
    r = np.ascontiguousarray(recs, dtype=np.uint32)
    a = np.ascontiguousarray(actions, dtype=np.int32)
    out = np.zeros_like(r)
    status = np.zeros(len(r), dtype=np.uint8)
    assert lib().hs_step(n, _ptr(r), _ptr(a), len(r), opts, _ptr(out), _ptr(status)) == 0
    return out, status


def areas(recs, n):
    r = np.ascontiguousarray(recs, dtype=np.uint32)
    out = np.zeros((len(r), 2), dtype=np.int32)
    assert lib().hs_areas(n, _ptr(r), len(r), _ptr(out)) == 0
    return out


def rollout_step(recs, n, seed, board0, t):
    """in-place on `recs` (uint32 [B, rec_words]); returns the sampled actions"""
    assert recs.dtype == np.uint32 and recs.flags.c_contiguous
    actions = np.zeros(len(recs), dtype=np.int32)
    assert lib().hs_rollout_step(n, _ptr(recs), len(recs), ctypes.c_uint64(seed), ctypes.c_uint64(board0),
                                 ctypes.c_uint64(t), _ptr(actions)) == 0
    return actions


def philox(board, t, seed):
    return int(lib().hs_philox(board, t, seed))
