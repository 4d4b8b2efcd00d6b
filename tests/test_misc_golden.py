"""Helpers around the step against outputs of the real reference (tests/golden/misc.npz):
the 8 symmetries element by element, random_symmetry under seeded numpy streams, the text rendering, liberties,
and update_pieces / adj_data.  The oracle functions and the pure-array host helpers run on the CPU; everything
that touches packed records or a kernel is in the `gpu` tests."""
import numpy as np
import pytest

import golden_io
from oracle import gogame_np as og


@pytest.mark.parametrize("n", golden_io.SYM_SIZES)
def test_oracle_symmetries_follow_the_reference_order(n):
    image, want, picks = golden_io.symmetries(n)
    got = og.all_symmetries(image)
    assert len(got) == 8
    for i in range(8):
        assert np.array_equal(got[i], want[i]), i
    for seed in range(len(picks)):
        np.random.seed(seed)
        assert np.array_equal(og.random_symmetry(image), picks[seed]), seed


@pytest.mark.parametrize("n", golden_io.SYM_SIZES)
def test_host_symmetries_follow_the_reference_order(n):
    """gymgo_b200.gogame.all_symmetries / random_symmetry are plain array helpers (no kernel): element i must be the
    reference's element i, for numpy and for torch inputs, and random_symmetry must consume the numpy stream alike"""
    import torch
    from gymgo_b200 import gogame
    image, want, picks = golden_io.symmetries(n)
    got = gogame.all_symmetries(image)
    got_t = gogame.all_symmetries(torch.from_numpy(image))
    for i in range(8):
        assert np.array_equal(got[i], want[i]), i
        assert np.array_equal(got_t[i].numpy(), want[i]), i
        assert np.array_equal(gogame.symmetry(image, i), want[i])
    for seed in range(len(picks)):
        np.random.seed(seed)
        assert np.array_equal(gogame.random_symmetry(image), picks[seed]), seed
    # leading batch axes: the board axes are the last two
    batch = np.stack([image, image[::-1]])
    for i in range(8):
        assert np.array_equal(gogame.symmetry(batch, i)[0], want[i])
    with pytest.raises(ValueError):
        gogame.symmetry(image, 8)


def test_oracle_text_and_liberties():
    for state, text, libs in golden_io.texts():
        assert og.to_text(state) == text
        b, w = og.liberties(state)
        assert np.array_equal(b, libs[0]) and np.array_equal(w, libs[1])
        assert og.num_liberties(state) == (int(libs[0].sum()), int(libs[1].sum()))


def test_oracle_update_pieces():
    cases = golden_io.update_pieces_cases()
    assert sum(c["groups"] > 0 for c in cases) > 20
    for c in cases:
        st = c["before"].copy()
        pts, surrounded = og.neighbours_and_surrounded(st, c["point"], c["player"])
        assert sorted(pts) == sorted(map(tuple, c["adj"])) and surrounded == c["surrounded"]
        killed = og.remove_captured(st, pts, c["player"])
        assert np.array_equal(st, c["after"])
        assert len(killed) == c["groups"]
        for k, grp in enumerate(killed):
            assert np.array_equal(grp, np.argwhere(c["killed"] == k + 1))


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n", golden_io.SYM_SIZES)
def test_gg_symmetry_matches_reference_elements(n):
    """gg_symmetry(sym) on packed records == element `sym` of the reference's all_symmetries (stone and INVD planes
    move, the constant planes are symmetric by nature)"""
    import torch
    from gymgo_b200.engine import GoEngine
    image, want, _ = golden_io.symmetries(n)
    e = GoEngine(n, "cuda:0")
    rec = e.pack(torch.from_numpy(image[None]).cuda())
    for sym in range(8):
        got = e.unpack(e.symmetry(rec, sym), dtype=torch.float64)[0].cpu().numpy()
        assert np.array_equal(got, want[sym]), sym
    # torch tensors on the device go through the same host helper
    from gymgo_b200 import gogame
    dev = gogame.all_symmetries(torch.from_numpy(image).cuda())
    for sym in range(8):
        assert np.array_equal(dev[sym].cpu().numpy(), want[sym])


@pytest.mark.gpu
def test_text_and_liberties_drop_in():
    from gymgo_b200 import gogame
    for state, text, libs in golden_io.texts():
        assert gogame.str(state) == text
        b, w = gogame.liberties(state)
        assert np.array_equal(b, libs[0]) and np.array_equal(w, libs[1])
        assert gogame.num_liberties(state) == (int(libs[0].sum()), int(libs[1].sum()))


@pytest.mark.gpu
def test_update_pieces_drop_in():
    """state_utils.update_pieces / batch_update_pieces (state_utils.py:159-211) through the capture kernel: the state is
    modified in place and the killed groups come back in the reference's order"""
    from gymgo_b200 import state_utils
    cases = golden_io.update_pieces_cases()
    for c in cases:
        st = c["before"].copy()
        adj, surrounded = state_utils.adj_data(st, np.array(c["point"]), c["player"])
        assert np.array_equal(np.asarray(adj), c["adj"]) and bool(surrounded) == c["surrounded"]
        killed = state_utils.update_pieces(st, adj, c["player"])
        assert np.array_equal(st, c["after"])
        assert len(killed) == c["groups"]
        for k, grp in enumerate(killed):
            assert np.array_equal(grp, np.argwhere(c["killed"] == k + 1))
    # batch variant: boards of one size together, "non-pass" = all of them
    for n in (5, 9):
        sub = [c for c in cases if c["n"] == n]
        batch = np.stack([c["before"] for c in sub])
        idx = np.arange(len(sub))
        killed = state_utils.batch_update_pieces(idx, batch, [c["adj"] for c in sub], np.array([c["player"] for c in sub]))
        assert np.array_equal(batch, np.stack([c["after"] for c in sub]))
        for c, groups in zip(sub, killed):
            assert len(groups) == c["groups"]
            for k, grp in enumerate(groups):
                assert np.array_equal(grp, np.argwhere(c["killed"] == k + 1))
