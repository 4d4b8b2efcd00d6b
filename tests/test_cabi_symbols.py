"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/gymgo_b200.h declares;
argument checks that need no device work; the package refuses to run the hot path without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gymgo_b200.h")).read()
    return sorted(set(re.findall(r"GG_API\s+[\w\s\*]+?\b(gg_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    from gymgo_b200 import _cabi
    lib = _cabi.lib()
    names = declared_symbols()
    assert len(names) >= 18 and set(names) == set(_cabi.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name
    assert lib.gg_version() == _cabi.GG_VERSION == 200


def test_layout_and_argument_errors_without_gpu():
    from gymgo_b200 import _cabi
    lib = _cabi.lib()
    assert [lib.gg_supported(n) for n in (1, 2, 9, 19, 20)] == [0, 1, 1, 1, 0]
    lay = _cabi.layout(9)
    assert lay == dict(rec_bytes=48, lanes_per_board=3, rows_per_lane=3, word_bits=32)
    assert _cabi.layout(19)["rec_bytes"] == 176 and _cabi.layout(7)["rec_bytes"] == 32
    with pytest.raises(_cabi.GymGoB200Error):
        _cabi.layout(25)
    buf = ctypes.create_string_buffer(64)
    addr = ctypes.addressof(buf)
    # argument validation happens before any CUDA call, so it is testable on a CPU box
    assert lib.gg_step(None, None, None, None, 4, 9, 0, None, 0, None, None, None, 0, 0.0, None) == _cabi.GG_EINVAL
    assert lib.gg_step(addr, addr, addr, None, 4, 21, 0, None, 0, None, None, None, 0, 0.0, None) == _cabi.GG_ESIZE
    assert lib.gg_step(addr, addr, addr, None, -1, 9, 0, None, 0, None, None, None, 0, 0.0, None) == _cabi.GG_EINVAL
    assert lib.gg_areas((addr | 15) + 1 + 4, 4, 9, addr, None) == _cabi.GG_EALIGN
    assert lib.gg_rollout(addr, 4, 9, 0, 0, 0, 3, 0, None, None, 0, 0, None, None, 0, 0.0, None) == _cabi.GG_EINVAL


def test_rollout_kernel_choice_and_new_argument_checks():
    from gymgo_b200 import _cabi
    lib = _cabi.lib()
    assert lib.gg_kernel_name(_cabi.GG_KERNEL_THREAD).decode().startswith("k_rollout_tpb")
    assert lib.gg_kernel_name(_cabi.GG_KERNEL_LANES).decode().startswith("k_rollout (")
    assert lib.gg_kernel_name(7) == b""
    assert lib.gg_rollout_kernel(9, 65536).decode().startswith("k_rollout_tpb")      # configs[1]
    assert "tpb" not in lib.gg_rollout_kernel(19, 16384).decode()                      # configs[2]
    buf = ctypes.create_string_buffer(64)
    addr = ctypes.addressof(buf)
    bad_kernel = lib.gg_rollout_with(5, addr, 4, 9, 0, 0, 0, 3, 1, None, None, 0, 0, None, None, 0, 0.0, None, 0, 0, None)
    assert bad_kernel == _cabi.GG_EINVAL
    assert lib.gg_rollout_workspace_bytes(9, 65536) >= 4 * (65536 // 40 + 2) and lib.gg_rollout_workspace_bytes(9, 65536) % 16 == 0
    small_ws = lib.gg_rollout_with(0, addr, 4096, 9, 0, 0, 0, 32, 32, None, None, 0, 0, None, None, 0, 0.0, addr, 16, 0, None)
    assert small_ws == _cabi.GG_EINVAL                                   # workspace too small for the batch
    assert lib.gg_rollout_with(0, addr, 4, 9, 0, 0, 0, 3, 1, None, None, 0, 0, None, None, 0, 0.0, None, 0, -1, None) == _cabi.GG_EINVAL
    assert lib.gg_step(addr, addr, addr, None, 4, 9, 64, None, 0, None, None, None, 0, 0.0, None) == _cabi.GG_EINVAL
    both = _cabi.GG_STEP_KERNEL_LANES | _cabi.GG_STEP_KERNEL_THREAD                      # contradictory kernel choice
    assert lib.gg_step(addr, addr, addr, None, 4, 9, both, None, 0, None, None, None, 0, 0.0, None) == _cabi.GG_EINVAL
    assert lib.gg_update_pieces(addr, addr, addr, addr, 4, 9, None) == _cabi.GG_EINVAL       # killed aliases rec
    assert lib.gg_update_pieces(None, addr, addr, addr, 4, 9, None) == _cabi.GG_EINVAL
    assert lib.gg_host_unpack(None, 4, 9, 1, addr, 1) == _cabi.GG_EINVAL
    assert lib.gg_host_unpack(addr, 4, 30, 1, addr, 1) == _cabi.GG_ESIZE


def test_host_codec_matches_the_record_layout():
    """gg_host_unpack (the only host-side entry point: packed records in host memory -> dense) against the host
    simulator's independent packer, every size and dtype"""
    import numpy as np
    import hostsim
    from gymgo_b200 import _cabi
    lib = _cabi.lib()
    for n in range(2, 20):
        rng = np.random.RandomState(n)
        st = (rng.uniform(size=(4100, 6, n, n)) < 0.5).astype(np.uint8)
        st[:, [2, 4, 5]] = rng.randint(2, size=(4100, 3))[:, :, None, None]
        rec = np.ascontiguousarray(hostsim.pack(st)).view(np.uint8)
        for dt, code in ((np.uint8, _cabi.GG_U8), (np.float32, _cabi.GG_F32), (np.float64, _cabi.GG_F64)):
            for threads in (1, 3):
                out = np.empty((4100, 6, n, n), dtype=dt)
                assert lib.gg_host_unpack(rec.ctypes.data, 4100, n, code, out.ctypes.data, threads) == 0
                assert np.array_equal(out, st.astype(dt)), (n, dt)
        half = np.empty((4100, 6, n, n), dtype=np.uint16)
        assert lib.gg_host_unpack(rec.ctypes.data, 4100, n, _cabi.GG_BF16, half.ctypes.data, 2) == 0
        assert np.array_equal(half, st.astype(np.uint16) * 0x3F80)


def test_host_codec_pool_is_safe_under_concurrent_callers_and_odd_shapes():
    """the codec's persistent worker pool: concurrent callers are serialised inside the library, thread counts above the
    work-item count and batches below one chunk work, and a misaligned destination takes the unaligned-store path"""
    import threading
    import numpy as np
    import hostsim
    from gymgo_b200 import _cabi
    lib = _cabi.lib()
    assert lib.gg_host_unpack_path().decode().split()[0] in ("avx512", "scalar")
    n = 9
    rng = np.random.RandomState(5)
    st = (rng.uniform(size=(9000, 6, n, n)) < 0.3).astype(np.uint8)
    st[:, [2, 4, 5]] = rng.randint(2, size=(9000, 3))[:, :, None, None]
    rec = np.ascontiguousarray(hostsim.pack(st)).view(np.uint8)
    outs = [np.empty((9000, 6, n, n), dtype=np.float32) for _ in range(4)]
    errs = []

    def work(k):
        for _ in range(5):
            if lib.gg_host_unpack(rec.ctypes.data, 9000, n, _cabi.GG_F32, outs[k].ctypes.data, 2 + k) != 0:
                errs.append(k)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs and all(np.array_equal(o, st.astype(np.float32)) for o in outs)
    for batch in (0, 1, 31, 33):                                         # below / around one 32-board chunk, 64 threads asked
        out = np.full((max(batch, 1), 6, n, n), 7, dtype=np.float16)
        assert lib.gg_host_unpack(rec.ctypes.data, batch, n, _cabi.GG_F16, out.ctypes.data, 64) == 0
        assert np.array_equal(out[:batch].view(np.uint16), st[:batch].astype(np.uint16) * 0x3C00)
    raw = np.empty(9000 * 6 * n * n * 8 + 64, dtype=np.uint8)            # float64, destination off by 8 bytes from 64
    off = (-raw.ctypes.data) % 64 + 8
    dst = raw[off:off + 9000 * 6 * n * n * 8].view(np.float64).reshape(9000, 6, n, n)
    assert lib.gg_host_unpack(rec.ctypes.data, 9000, n, _cabi.GG_F64, dst.ctypes.data, 3) == 0
    assert np.array_equal(dst, st.astype(np.float64))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import gymgo_b200
    from gymgo_b200 import _cabi
    with pytest.raises(_cabi.GymGoB200Error):
        gymgo_b200.make("gym_go:go-v0", size=7)
    with pytest.raises(_cabi.GymGoB200Error):
        from gymgo_b200 import gogame
        gogame.next_state(gogame.init_state(5), 3)
