"""ctypes binding of libgymgo_b200.so (include/gymgo_b200.h) - the only way Python reaches the kernels.

There is NO CPU fallback: if the library is missing it is built with nvcc (gymgo_b200.build); if that
fails, or a call is made without a CUDA device, an exception is raised."""
import ctypes
import os

from . import build as _build

_LIB = None

# return codes / enums of include/gymgo_b200.h
GG_OK, GG_EINVAL, GG_ESIZE, GG_EALIGN, GG_ECUDA = 0, -1, -2, -3, -4
GG_ST_OK, GG_ST_INVALID_MOVE, GG_ST_OUT_OF_RANGE, GG_ST_GAME_OVER = 0, 1, 2, 3
GG_U8, GG_F32, GG_F64, GG_BF16, GG_F16 = 0, 1, 2, 3, 4
GG_STEP_CANONICAL, GG_STEP_REFUSE_DONE, GG_STEP_AUTO_RESET, GG_STEP_RESET_SKIPS_ACTION = 1, 2, 4, 8
GG_STEP_KERNEL_LANES, GG_STEP_KERNEL_THREAD = 16, 32
GG_KERNEL_AUTO, GG_KERNEL_LANES, GG_KERNEL_THREAD = -1, 0, 1
GG_VERSION = 200

GG_REWARD_NONE, GG_REWARD_REAL, GG_REWARD_HEURISTIC = 0, 1, 2

EXPORTS = ("gg_version", "gg_last_cuda_error", "gg_supported", "gg_set_device", "gg_layout", "gg_pack", "gg_unpack", "gg_reset",
           "gg_step", "gg_rollout_step", "gg_rollout", "gg_rollout_with", "gg_rollout_workspace_bytes", "gg_rollout_kernel", "gg_kernel_name", "gg_update_pieces", "gg_sample_legal", "gg_valid_moves", "gg_children", "gg_areas",
           "gg_canonical", "gg_symmetry", "gg_host_unpack", "gg_host_unpack_path", "gg_probe_write")

_ERR = {GG_EINVAL: "GG_EINVAL (bad argument)", GG_ESIZE: "GG_ESIZE (board size not supported, build has 2..19)",
        GG_EALIGN: "GG_EALIGN (buffer not 16-byte aligned)", GG_ECUDA: "GG_ECUDA"}


class GymGoB200Error(RuntimeError):
    pass


def library_path():
    return _build.LIB


def lib():
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build()           # raises if nvcc is missing / compilation fails
    elif not _build.up_to_date():
        # the sources under csrc/ changed after the library was built: rebuild where a compiler exists (a no-op
        # otherwise would silently run stale kernels); on a box without nvcc fall through to the version check
        try:
            _build.nvcc()
        except RuntimeError:
            pass
        else:
            path = _build.build()
    L = ctypes.CDLL(path)
    L.gg_version.restype = ctypes.c_int
    if L.gg_version() != GG_VERSION:
        raise GymGoB200Error("%s is ABI version %d, this binding needs %d: run `python -m gymgo_b200.build --force`"
                             % (path, L.gg_version(), GG_VERSION))
    vp, i64, u64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32
    ip = ctypes.POINTER(ctypes.c_int)
    L.gg_version.restype = i32
    L.gg_last_cuda_error.restype = ctypes.c_char_p
    L.gg_supported.argtypes = [i32]
    L.gg_layout.argtypes = [i32, ip, ip, ip, ip]
    L.gg_pack.argtypes = [vp, i32, i64, i32, vp, vp]
    L.gg_unpack.argtypes = [vp, i64, i32, i32, vp, vp]
    L.gg_reset.argtypes = [vp, i64, i32, vp, vp]
    f32 = ctypes.c_float
    L.gg_set_device.argtypes = [i32]
    L.gg_step.argtypes = [vp, vp, vp, vp, i64, i32, u32, vp, i32, vp, vp, vp, i32, f32, vp]
    L.gg_rollout_step.argtypes = [vp, i64, i32, u64, u64, u64, vp, vp, i32, vp, vp, vp, i32, f32, vp]
    L.gg_rollout.argtypes = [vp, i64, i32, u64, u64, u64, i32, i32, vp, vp, i32, i32, vp, vp, i32, f32, vp]
    L.gg_rollout_with.argtypes = [i32] + L.gg_rollout.argtypes[:-1] + [vp, i64, i32, vp]
    L.gg_rollout_workspace_bytes.argtypes = [i32, i64]
    L.gg_rollout_workspace_bytes.restype = i64
    L.gg_rollout_kernel.argtypes = [i32, i64]
    L.gg_rollout_kernel.restype = ctypes.c_char_p
    L.gg_kernel_name.argtypes = [i32]
    L.gg_kernel_name.restype = ctypes.c_char_p
    L.gg_update_pieces.argtypes = [vp, vp, vp, vp, i64, i32, vp]
    L.gg_sample_legal.argtypes = [vp, i64, i32, u64, u64, u64, vp, vp]
    L.gg_valid_moves.argtypes = [vp, i64, i32, i32, i32, vp, vp]
    L.gg_children.argtypes = [vp, i64, i32, u32, vp, vp, i32, vp, vp, vp]
    L.gg_areas.argtypes = [vp, i64, i32, vp, vp]
    L.gg_canonical.argtypes = [vp, vp, i64, i32, vp]
    L.gg_symmetry.argtypes = [vp, vp, i64, i32, i32, vp]
    L.gg_host_unpack.argtypes = [vp, i64, i32, i32, vp, i32]
    L.gg_host_unpack_path.restype = ctypes.c_char_p
    L.gg_probe_write.argtypes = [vp, i64, i64, vp]
    for name in EXPORTS:
        getattr(L, name)                # fail early if a symbol is missing
    _LIB = L
    return L


def check(rc):
    if rc != GG_OK:
        msg = _ERR.get(rc, "error %d" % rc)
        if rc == GG_ECUDA:
            msg += ": " + lib().gg_last_cuda_error().decode()
        raise GymGoB200Error(msg)


def layout(n):
    """-> dict(rec_bytes, lanes_per_board, rows_per_lane, word_bits) of the packed record for side n."""
    a, b, c, d = (ctypes.c_int() for _ in range(4))
    check(lib().gg_layout(int(n), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
    return dict(rec_bytes=a.value, lanes_per_board=b.value, rows_per_lane=c.value, word_bits=d.value)
