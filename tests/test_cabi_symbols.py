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
    assert lib.gg_version() == 100


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
