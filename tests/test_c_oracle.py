"""Pins the plain-C oracle (oracle/go_oracle.c) against the reference-generated fixtures and against
the numpy/scipy port on fresh random positions.  CPU only."""
import numpy as np
import pytest

import golden_io
from oracle import c_oracle as co
from oracle import gogame_np as og


@pytest.mark.parametrize("n", golden_io.TRAJ_SIZES)
def test_trajectories(n):
    S, A, AR, VM = golden_io.trajectory(n)
    idx = np.flatnonzero(A >= 0)
    out, status = co.batch_next_states(S[idx], A[idx])
    assert not status.any()
    assert np.array_equal(out, S[idx + 1].astype(np.uint8))
    assert np.array_equal(co.batch_areas(S), AR.astype(np.int32))
    for i in range(0, len(S), 7):
        assert np.array_equal(co.valid_moves(S[i]), VM[i].astype(np.uint8))


@pytest.mark.parametrize("n", golden_io.SOUP_SIZES)
def test_soup(n):
    S0, A, S1, AR = golden_io.soup(n)
    out, status = co.batch_next_states(S0, A)
    assert not status.any()
    assert np.array_equal(out, S1.astype(np.uint8))
    assert np.array_equal(co.batch_areas(S0), AR.astype(np.int32))


@pytest.mark.parametrize("n", golden_io.CHILDREN_SIZES)
def test_children(n):
    P, C0, C1 = golden_io.children(n)
    for i in range(len(P)):
        for canon, ref in ((False, C0[i]), (True, C1[i])):
            kids, valid, bad = co.children(P[i], canonical=canon)
            assert not bad
            assert np.array_equal(kids, ref.astype(np.uint8))
            assert np.array_equal(valid, og.valid_moves(P[i]).astype(np.uint8))


@pytest.mark.parametrize("case", golden_io.kat_cases(), ids=lambda c: c["name"])
def test_kat(case):
    st = case["states"][0]
    for i, a in enumerate(case["actions"]):
        st = co.next_state(st, int(a))
        assert np.array_equal(st, case["states"][i + 1].astype(np.uint8))
    if case["raises"] >= 0 and not og.game_ended(st):
        with pytest.raises(AssertionError):
            co.next_state(st, case["raises"])


def test_status_codes():
    st = np.zeros((6, 5, 5), dtype=np.uint8)
    st[3, 2, 2] = 1
    out, status = co.batch_next_states(np.stack([st, st, st]), [12, 26, -1])
    assert list(status) == [1, 2, 2]
    assert np.array_equal(out, np.stack([st, st, st]))      # refused boards are left unchanged


@pytest.mark.parametrize("n", (2, 3, 6, 9))
def test_against_numpy_port_on_random_soup(n):
    rng = np.random.RandomState(77 + n)
    for _ in range(150):
        dens = rng.uniform(0.1, 0.98)
        r = rng.uniform(size=(n, n))
        st = np.zeros((6, n, n))
        st[0] = r < dens / 2
        st[1] = (r >= dens / 2) & (r < dens)
        t = int(rng.randint(2))
        st[2] = t
        ko = int(rng.randint(-1, n * n))
        mask = og.invalid_mask(st, 1 - t, None if ko < 0 else (ko // n, ko % n))
        assert np.array_equal(co.invalid_mask(st[0], st[1], t, ko), mask.astype(np.uint8))
        st[3] = mask
        st[4] = int(rng.randint(2))
        acts = np.flatnonzero(np.append(1 - st[3].flatten(), 1))
        a = int(rng.choice(acts))
        for canon in (False, True):
            assert np.array_equal(co.next_state(st, a, canon), og.next_state(st, a, canon).astype(np.uint8))
        assert co.areas(st) == tuple(int(x) for x in og.areas(st))


def test_rollout_runs():
    assert co.rollout(9, 2000, 1) != co.rollout(9, 2000, 2)
