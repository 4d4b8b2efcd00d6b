"""Where the real reference is present (the build container), check both oracles against it directly on FRESH
random play and soups - in a subprocess, because importing the reference's `gym_go` must not shadow this
repo's `gym_go` alias inside the test process.  Skipped on machines without /root/reference (the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _refshim  # noqa: E402

SCRIPT = r'''
import sys, warnings
import numpy as np
warnings.simplefilter("ignore")
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests/golden")
import _refshim
gym, ref, govars = _refshim.load_reference()
from gym_go import state_utils as ref_su
from oracle import gogame_np as og
from oracle import c_oracle as co
rng = np.random.RandomState(%(seed)d)
checked = 0
for n in (3, 5, 7, 9, 13):
    state = ref.init_state(n)
    for t in range(%(plies)d):
        vm = ref.valid_moves(state)
        assert np.array_equal(vm, og.valid_moves(state)) and np.array_equal(vm, co.valid_moves(state))
        a = int(rng.choice(np.flatnonzero(vm)))
        canon = bool(t %% 2)
        nxt = ref.next_state(state, a, canonical=canon)
        assert np.array_equal(nxt, og.next_state(state, a, canon)), (n, t)
        assert np.array_equal(nxt, co.next_state(state, a, canon)), (n, t)
        assert tuple(ref.areas(state)) == tuple(og.areas(state)) == co.areas(state)
        checked += 1
        state = ref.next_state(state, a)
        if ref.game_ended(state):
            state = ref.init_state(n)
    # soups: arbitrary stones, the reference's own mask, one legal action
    for _ in range(60):
        dens = rng.uniform(0.1, 0.95)
        r = rng.uniform(size=(n, n))
        st = ref.init_state(n)
        st[0] = r < dens / 2
        st[1] = (r >= dens / 2) & (r < dens)
        turn = int(rng.randint(2))
        st[2] = turn
        st[3] = ref_su.compute_invalid_moves(st, 1 - turn, None)
        assert np.array_equal(st[3], og.invalid_mask(st, 1 - turn)) 
        assert np.array_equal(st[3], co.invalid_mask(st[0], st[1], turn))
        a = int(rng.choice(np.flatnonzero(np.append(1 - st[3].flatten(), 1))))
        nxt = ref.next_state(st, a)
        assert np.array_equal(nxt, og.next_state(st, a)) and np.array_equal(nxt, co.next_state(st, a))
        checked += 1
    kids = ref.children(state, canonical=True, padded=True)
    assert np.array_equal(kids, og.children(state, canonical=True, padded=True))
    assert np.array_equal(kids, co.children(state, canonical=True)[0])
print("checked", checked)
'''


@pytest.mark.skipif(not _refshim.reference_available(), reason="reference not present on this machine")
def test_oracles_against_the_real_reference():
    code = SCRIPT % dict(root=ROOT, seed=20260925, plies=160)
    p = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "checked" in p.stdout
