"""CPU check of the DEVICE algorithm (gymgo_b200/csrc/gg_algo.cuh) through the host simulator:
record layout round trip, step / areas / sampler against the reference-generated fixtures and the
C oracle on random play, soups and long rollouts.  No GPU needed; the same template is what the
sm_100a kernels instantiate."""
import numpy as np
import pytest

import golden_io
import hostsim
from oracle import c_oracle as co

SIZES = tuple(range(2, 20))


def random_soup(n, count, rng):
    dens = rng.uniform(0.05, 1.0, size=(count, 1, 1))
    cut = rng.uniform(0.2, 0.8, size=(count, 1, 1))
    r = rng.uniform(size=(count, n, n))
    st = np.zeros((count, 6, n, n), dtype=np.uint8)
    st[:, 0] = r < dens * cut
    st[:, 1] = (r >= dens * cut) & (r < dens)
    turn = rng.randint(2, size=count)
    st[:, 2] = turn[:, None, None]
    for i in range(count):
        ko = int(rng.randint(-1, n * n)) if rng.uniform() < 0.3 else -1
        st[i, 3] = co.invalid_mask(st[i, 0], st[i, 1], int(turn[i]), ko)
    st[:, 4] = rng.randint(2, size=count)[:, None, None]
    st[:, 5] = (rng.uniform(size=count) < 0.1)[:, None, None]
    return st


@pytest.mark.parametrize("n", SIZES)
def test_pack_unpack_roundtrip(n):
    rng = np.random.RandomState(n)
    st = (rng.uniform(size=(64, 6, n, n)) < 0.5).astype(np.uint8)
    st[:, [2, 4, 5]] = rng.randint(2, size=(64, 3))[:, :, None, None]
    recs = hostsim.pack(st)
    assert recs.shape[1] * 4 == hostsim.layout(n)["rec_bytes"] and recs.shape[1] % 4 == 0
    assert np.array_equal(hostsim.unpack(recs, n), st)


@pytest.mark.parametrize("n", golden_io.TRAJ_SIZES)
def test_golden_trajectories(n):
    S, A, AR, _ = golden_io.trajectory(n)
    idx = np.flatnonzero(A >= 0)
    out, status = hostsim.step(hostsim.pack(S[idx]), A[idx], n)
    assert not status.any()
    assert np.array_equal(hostsim.unpack(out, n), S[idx + 1].astype(np.uint8))
    assert np.array_equal(hostsim.areas(hostsim.pack(S), n), AR.astype(np.int32))


@pytest.mark.parametrize("n", golden_io.SOUP_SIZES)
def test_golden_soup(n):
    S0, A, S1, AR = golden_io.soup(n)
    out, status = hostsim.step(hostsim.pack(S0), A, n)
    assert not status.any()
    assert np.array_equal(hostsim.unpack(out, n), S1.astype(np.uint8))
    assert np.array_equal(hostsim.areas(hostsim.pack(S0), n), AR.astype(np.int32))


@pytest.mark.parametrize("case", golden_io.kat_cases(), ids=lambda c: c["name"])
def test_golden_kat(case):
    n = case["states"].shape[2]
    rec = hostsim.pack(case["states"][:1])
    for i, a in enumerate(case["actions"]):
        rec, status = hostsim.step(rec, [int(a)], n, opts=2)
        assert status[0] == 0
        assert np.array_equal(hostsim.unpack(rec, n)[0], case["states"][i + 1].astype(np.uint8)), (case["name"], i)
    if case["raises"] >= 0:
        _, status = hostsim.step(rec, [case["raises"]], n, opts=2)
        assert status[0] in (1, 3)


@pytest.mark.parametrize("n", SIZES)
def test_random_soup_vs_c_oracle(n):
    rng = np.random.RandomState(100 + n)
    count = 3000 if n <= 9 else 800
    st = random_soup(n, count, rng)
    valid = 1 - st[:, 3].reshape(count, -1)
    acts = np.empty(count, dtype=np.int32)
    for i in range(count):
        choices = np.append(np.flatnonzero(valid[i]), n * n)
        acts[i] = rng.choice(choices)
    # sprinkle refused actions: occupied/invalid points and out-of-range indices
    bad = rng.uniform(size=count) < 0.1
    for i in np.flatnonzero(bad):
        inv = np.flatnonzero(st[i, 3].reshape(-1))
        acts[i] = rng.choice(inv) if len(inv) and rng.uniform() < 0.7 else rng.choice([-1, n * n + 1, 10 ** 6])
    for opts, canon in ((0, False), (1, True)):
        want, wstatus = co.batch_next_states(st, acts, canon)
        got, gstatus = hostsim.step(hostsim.pack(st), acts, n, opts=opts)
        assert np.array_equal(gstatus, wstatus)
        assert np.array_equal(hostsim.unpack(got, n), want)
    assert np.array_equal(hostsim.areas(hostsim.pack(st), n), co.batch_areas(st))


def test_refuse_done_option():
    st = np.zeros((2, 6, 5, 5), dtype=np.uint8)
    st[1, 5] = 1
    rec = hostsim.pack(st)
    _, status = hostsim.step(rec, [3, 3], 5, opts=2)
    assert list(status) == [0, 3]
    out, status = hostsim.step(rec, [3, 3], 5, opts=0)      # gogame.next_state itself does not care (A.3)
    assert list(status) == [0, 0]
    assert hostsim.unpack(out, 5)[1, 5].all() and hostsim.unpack(out, 5)[1, 0, 0, 3] == 1


def philox_numpy(board, t, seed):
    """independent vectorised Philox4x32-10, word 0"""
    board = np.asarray(board, dtype=np.uint64)
    c = [board & 0xFFFFFFFF, board >> 32, np.full_like(board, t & 0xFFFFFFFF), np.full_like(board, t >> 32)]
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> 32) ^ c[1] ^ np.uint64(k0), p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ np.uint64(k1), p0 & 0xFFFFFFFF]
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c[0].astype(np.uint64)


def test_philox_known_answer_and_numpy_twin():
    # Random123 known-answer vector: counter = key = 0 -> first word 0x6627e8d5
    assert hostsim.philox(0, 0, 0) == 0x6627E8D5
    b = np.array([0, 1, 2, 65535, 2 ** 33 + 5], dtype=np.uint64)
    want = philox_numpy(b, 7, 0x1234567890)
    got = [hostsim.philox(int(x), 7, 0x1234567890) for x in b]
    assert list(want) == got


@pytest.mark.parametrize("n,boards,steps", ((5, 64, 150), (9, 96, 300), (13, 24, 200), (19, 16, 260)))
def test_rollout_matches_oracle_replay(n, boards, steps):
    """The fused reset+sample+step of the rollout, replayed ply by ply through the C oracle, with the
    sampler re-derived in numpy (k-th valid action, pass last)."""
    recs = hostsim.pack(np.zeros((boards, 6, n, n), dtype=np.uint8))
    dense = np.zeros((boards, 6, n, n), dtype=np.uint8)
    seed, board0 = 0xABCDEF12345, 1000
    finished = 0
    for t in range(steps):
        acts = hostsim.rollout_step(recs, n, seed, board0, t)
        done = dense[:, 5, 0, 0] == 1
        finished += int(done.sum())
        dense[done] = 0
        rnd = philox_numpy(np.arange(boards, dtype=np.uint64) + np.uint64(board0), t, seed)
        valid = np.concatenate([1 - dense[:, 3].reshape(boards, -1), np.ones((boards, 1), dtype=np.uint8)], axis=1)
        count = valid.sum(axis=1).astype(np.uint64)
        k = ((rnd * count) >> np.uint64(32)).astype(np.int64)
        want_act = np.array([np.flatnonzero(valid[i])[k[i]] for i in range(boards)], dtype=np.int32)
        assert np.array_equal(acts, want_act), t
        dense, status = co.batch_next_states(dense, acts)
        assert not status.any()
        assert np.array_equal(hostsim.unpack(recs, n), dense), t
    if n <= 9:
        assert finished > 0      # games really end and restart inside the window
