"""Pins the numpy/scipy oracle port (oracle/gogame_np.py) against fixtures produced by the real
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import golden_io
from oracle import gogame_np as og


@pytest.mark.parametrize("case", golden_io.kat_cases(), ids=lambda c: c["name"])
def test_kat_sequences(case):
    n = case["states"].shape[2]
    env = og.EnvOracle(n, komi=case["komi"], reward_method=case["method"])
    st = env.reset()
    assert np.array_equal(st, case["states"][0])
    for i, a in enumerate(case["actions"]):
        st, rew, done, info = env.step(int(a))
        assert np.array_equal(st, case["states"][i + 1]), (case["name"], i)
        assert float(rew) == case["rewards"][i]
        assert int(done) == case["dones"][i]
        assert info["turn"] == case["turns"][i]
        assert int(info["prev_player_passed"]) == case["prev_pass"][i]
    if case["raises"] >= 0:
        with pytest.raises(Exception):
            env.step(case["raises"])


@pytest.mark.parametrize("n", golden_io.TRAJ_SIZES)
def test_trajectories(n):
    S, A, AR, VM = golden_io.trajectory(n)
    step = max(1, len(S) // 400) if n >= 13 else 1   # keep the CPU suite to minutes
    for i in range(0, len(S), step):
        assert np.array_equal(og.valid_moves(S[i]), VM[i])
        assert tuple(og.areas(S[i])) == tuple(AR[i])
        if A[i] >= 0:
            assert np.array_equal(og.next_state(S[i], int(A[i])), S[i + 1]), (n, i)


@pytest.mark.parametrize("n", golden_io.SOUP_SIZES)
def test_soup(n):
    S0, A, S1, AR = golden_io.soup(n)
    for i in range(len(S0)):
        assert np.array_equal(og.next_state(S0[i], int(A[i])), S1[i]), (n, i)
        assert tuple(og.areas(S0[i])) == tuple(AR[i])


@pytest.mark.parametrize("n", golden_io.CHILDREN_SIZES)
def test_children(n):
    P, C0, C1 = golden_io.children(n)
    for i in range(len(P)):
        assert np.array_equal(og.children(P[i], canonical=False, padded=True), C0[i])
        assert np.array_equal(og.children(P[i], canonical=True, padded=True), C1[i])
        k = og.children(P[i], canonical=False, padded=False)
        assert np.array_equal(k, C0[i][og.valid_moves(P[i]) > 0])


def test_env_games():
    for g in golden_io.env_games():
        env = og.EnvOracle(7, komi=g["komi"], reward_method=g["method"])
        st = env.reset()
        for i, a in enumerate(g["actions"]):
            st, rew, done, _ = env.step(int(a))
            assert np.array_equal(st, g["states"][i + 1])
            assert float(rew) == g["rewards"][i], (g["key"], i)
            assert int(done) == g["dones"][i]
        assert float(env.winning()) == g["winning"]


def test_purity_and_batch():
    S0, A, S1 = golden_io.transitions(5)
    keep = S0[:50].copy()
    out = og.batch_next_states(S0[:50], A[:50])
    assert np.array_equal(S0[:50], keep)          # inputs untouched (test_basics.py:48-52)
    assert np.array_equal(out, S1[:50])
    c = og.batch_canonical_form(out)
    assert np.array_equal(og.batch_canonical_form(c), c)   # idempotent (test_batch_fns.py:14-34)
    assert (og.batch_turn(c) == 0).all()
