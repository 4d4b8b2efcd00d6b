"""Property-based CPU checks (hypothesis) of the device algorithm (host simulator) against the C oracle:
arbitrary board sizes, stone soups, turn/ko/pass/done combinations and actions, plus algebraic properties the
domain offers (colour-swap symmetry of the rules, dihedral symmetry of areas, idempotence of canonical form)."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import hostsim
from oracle import c_oracle as co


@st.composite
def positions(draw):
    n = draw(st.integers(2, 19))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.RandomState(seed)
    dens, cut = rng.uniform(0.0, 1.0), rng.uniform(0.1, 0.9)
    r = rng.uniform(size=(n, n))
    s = np.zeros((6, n, n), dtype=np.uint8)
    s[0] = r < dens * cut
    s[1] = (r >= dens * cut) & (r < dens)
    turn = int(rng.randint(2))
    s[2] = turn
    ko = int(rng.randint(-1, n * n)) if rng.uniform() < 0.4 else -1
    s[3] = co.invalid_mask(s[0], s[1], turn, ko)
    s[4] = int(rng.randint(2))
    s[5] = int(rng.uniform() < 0.1)
    action = int(rng.randint(-2, n * n + 3))
    return s, action


@settings(max_examples=300, deadline=None, suppress_health_check=list(HealthCheck))
@given(positions())
def test_step_matches_oracle(pos):
    s, action = pos
    n = s.shape[1]
    for opts, canon in ((0, False), (1, True)):
        want, wstatus = co.batch_next_states(s[None], [action], canon)
        got, gstatus = hostsim.step(hostsim.pack(s[None]), [action], n, opts=opts)
        assert gstatus[0] == wstatus[0]
        assert np.array_equal(hostsim.unpack(got, n), want)
    assert np.array_equal(hostsim.areas(hostsim.pack(s[None]), n), co.batch_areas(s[None]))


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
@given(positions())
def test_colour_swap_symmetry(pos):
    """swapping the colours (stones and side to move) commutes with a legal ply"""
    s, action = pos
    n = s.shape[1]
    t = s.copy()
    t[0], t[1] = s[1], s[0]
    t[2] = 1 - s[2]
    a, sa = hostsim.step(hostsim.pack(s[None]), [action], n)
    b, sb = hostsim.step(hostsim.pack(t[None]), [action], n)
    assert sa[0] == sb[0]
    da, db = hostsim.unpack(a, n)[0], hostsim.unpack(b, n)[0]
    assert np.array_equal(da[0], db[1]) and np.array_equal(da[1], db[0])
    assert np.array_equal(da[3:], db[3:]) and np.array_equal(da[2], 1 - db[2])


@settings(max_examples=100, deadline=None, suppress_health_check=list(HealthCheck))
@given(positions(), st.integers(0, 7))
def test_areas_are_dihedral_invariant(pos, sym):
    s, _ = pos
    n = s.shape[1]
    t = np.rot90(s, sym % 4, axes=(1, 2))
    if sym >= 4:
        t = np.flip(t, 2)
    t = np.ascontiguousarray(t)
    assert np.array_equal(hostsim.areas(hostsim.pack(s[None]), n), hostsim.areas(hostsim.pack(t[None]), n))
