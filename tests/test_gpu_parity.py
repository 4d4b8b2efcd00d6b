"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the oracle on the same
inputs - bit-exact.  Run on the B200 box with `-m gpu`."""
import numpy as np
import pytest
import torch

import golden_io
import hostsim
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
ALL_SIZES = tuple(range(2, 20))


@pytest.fixture(scope="module")
def eng():
    from gymgo_b200.engine import engine
    return engine


def soup(n, count, seed):
    from test_device_algo_hostsim import random_soup
    return random_soup(n, count, np.random.RandomState(seed))


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("n", ALL_SIZES)
def test_layout_and_codec_roundtrip(eng, n):
    e = eng(n)
    assert e.layout["rec_bytes"] == hostsim.layout(n)["rec_bytes"]
    rng = np.random.RandomState(n)
    st = (rng.uniform(size=(257, 6, n, n)) < 0.5).astype(np.uint8)
    st[:, [2, 4, 5]] = rng.randint(2, size=(257, 3))[:, :, None, None]
    want_rec = hostsim.pack(st)
    for dt in (torch.uint8, torch.float32, torch.float64):
        rec = e.pack(dev(st, dt))
        assert np.array_equal(rec.cpu().numpy().view(np.uint32), want_rec)      # same record bytes as the host sim
        back = e.unpack(rec, dtype=dt)
        assert back.dtype == dt and np.array_equal(back.cpu().numpy().astype(np.uint8), st)


@pytest.mark.parametrize("kernel", ("lanes", "thread"))
@pytest.mark.parametrize("n", ALL_SIZES)
def test_step_vs_oracle_on_soup(eng, n, kernel):
    """gg_step against the C oracle on random positions incl. refused moves, for BOTH single-ply kernels
    (k_step: lane-sliced boards; k_step_tpb: one board per thread)"""
    e = eng(n)
    count = 4001 if n <= 9 else 1203
    rng = np.random.RandomState(500 + n)
    st = soup(n, count, 500 + n)
    valid = 1 - st[:, 3].reshape(count, -1)
    acts = np.array([rng.choice(np.append(np.flatnonzero(valid[i]), n * n)) for i in range(count)], dtype=np.int32)
    for i in np.flatnonzero(rng.uniform(size=count) < 0.1):      # refused actions
        inv = np.flatnonzero(st[i, 3].reshape(-1))
        acts[i] = rng.choice(inv) if len(inv) and rng.uniform() < 0.7 else rng.choice([-1, n * n + 1, 10 ** 6])
    rec = e.pack(dev(st))
    for canon in (False, True):
        want, wstatus = co.batch_next_states(st, acts, canon)
        res = e.step(rec, acts, canonical=canon, obs_dtype=torch.uint8, want_done=True, want_areas=True, kernel=kernel)
        assert np.array_equal(res["status"].cpu().numpy(), wstatus)
        assert np.array_equal(e.unpack(res["rec"], dtype=torch.uint8).cpu().numpy(), want)
        assert np.array_equal(res["obs"].cpu().numpy(), want)                     # fused observation
        assert np.array_equal(res["done"].cpu().numpy(), want[:, 5, 0, 0])
        assert np.array_equal(res["areas"].cpu().numpy(), co.batch_areas(want))
    # refuse_done option
    res = e.step(rec, acts, refuse_done=True, kernel=kernel)
    done = st[:, 5, 0, 0] == 1
    got = res["status"].cpu().numpy()
    assert (got[done & (acts >= 0) & (acts <= n * n)] == 3).all()
    assert np.array_equal(e.areas(rec).cpu().numpy(), co.batch_areas(st))


@pytest.mark.parametrize("n", golden_io.TRAJ_SIZES)
def test_golden_trajectories(eng, n):
    e = eng(n)
    S, A, AR, VM = golden_io.trajectory(n)
    idx = np.flatnonzero(A >= 0)
    res = e.step(e.pack(dev(S[idx])), A[idx], obs_dtype=torch.float32)
    assert not res["status"].any()
    assert np.array_equal(res["obs"].cpu().numpy(), S[idx + 1].astype(np.float32))
    rec_all = e.pack(dev(S))
    assert np.array_equal(e.areas(rec_all).cpu().numpy(), AR.astype(np.int32))
    assert np.array_equal(e.valid_moves(rec_all, ended_quirk=True, dtype=torch.float64).cpu().numpy(), VM)


@pytest.mark.parametrize("n", golden_io.SOUP_SIZES)
def test_golden_soup(eng, n):
    e = eng(n)
    S0, A, S1, AR = golden_io.soup(n)
    res = e.step(e.pack(dev(S0)), A, obs_dtype=torch.float32)
    assert not res["status"].any()
    assert np.array_equal(res["obs"].cpu().numpy(), S1.astype(np.float32))
    assert np.array_equal(e.areas(e.pack(dev(S0))).cpu().numpy(), AR.astype(np.int32))


@pytest.mark.parametrize("kernel", ("lanes", "thread"))
@pytest.mark.parametrize("n", (3, 7, 9, 19))
@pytest.mark.parametrize("batch", (1, 7, 39, 41, 130, 1000))
@pytest.mark.parametrize("dtype", (torch.float32, torch.uint8, torch.bfloat16, torch.float16))
def test_obs_emission_tail_tiles(eng, n, batch, dtype, kernel):
    """the fused dense output equals the independent unpack kernel for ragged batch sizes (both single-ply kernels)"""
    e = eng(n)
    st = soup(n, batch, 9000 + n + batch)
    st[:, 5] = 0
    rec = e.pack(dev(st))
    acts = e.sample_legal(rec, 3, 0, 0)
    guard = torch.full((batch + 1, 6, n, n), 7, dtype=dtype, device="cuda")      # one extra board as canary
    res = e.step(rec, acts, obs=guard[:batch], kernel=kernel)
    assert torch.equal(res["obs"], e.unpack(res["rec"], dtype=dtype))
    assert torch.equal(res["obs"].float(), e.unpack(res["rec"], dtype=torch.float32))
    assert bool((guard[batch] == 7).all())                                          # nothing written past the end
    assert torch.equal(e.pack(res["obs"]), res["rec"])                              # and it packs back to the record


@pytest.mark.parametrize("n,boards", ((9, 2500), (8, 100), (6, 777), (5, 64), (4, 65), (3, 31), (2, 200), (13, 97), (19, 75)))
def test_step_kernels_are_bit_identical_for_every_option(eng, n, boards):
    """k_step (lane-sliced) and k_step_tpb (one board per thread) agree on records, status, observation, done, areas and
    rewards for every option of gg_step - canonical, refuse-done, auto-reset, reset-skips-action, both reward modes -
    along games that contain finished boards, refused and out-of-range actions; in-place and out-of-place"""
    e = eng(n)
    rec = e.new_records(boards)
    e.rollout(rec, 3, 0, 0, 3 * n * n, plies_per_launch=16)               # mixed game phases incl. finished boards
    rng = np.random.RandomState(n)
    combos = [dict(), dict(canonical=True), dict(refuse_done=True), dict(auto_reset=True, refuse_done=True),
              dict(auto_reset="skip", refuse_done=True), dict(auto_reset=True, canonical=True)]
    a_rec, b_rec = rec.clone(), rec.clone()
    for t in range(40):
        acts = e.sample_legal(a_rec, 11, 0, t)
        bad = torch.from_numpy(rng.uniform(size=boards) < 0.15).cuda()
        junk = torch.from_numpy(rng.randint(-2, n * n + 3, size=boards).astype(np.int32)).cuda()
        acts = torch.where(bad, junk, acts)                                  # illegal / out-of-range actions in the mix
        opt = combos[t % len(combos)]
        dt = (torch.float32, torch.uint8, torch.bfloat16)[t % 3]
        kw = dict(obs_dtype=dt, want_done=True, want_areas=(t % 2 == 0), reward_mode=1 + t % 2, komi=0.5, **opt)
        ra = e.step(a_rec, acts, out=a_rec if t % 4 else None, kernel="lanes", **kw)
        rb = e.step(b_rec, acts, out=b_rec if t % 4 else None, kernel="thread", **kw)
        for key in ("rec", "status", "obs", "done", "areas", "reward"):
            assert (ra[key] is None) == (rb[key] is None), key
            if ra[key] is not None:
                assert torch.equal(ra[key], rb[key]), (t, key, opt)
        a_rec, b_rec = ra["rec"], rb["rec"]


@pytest.mark.parametrize("n,boards,steps", ((5, 333, 120), (9, 1000, 260), (13, 200, 200), (19, 150, 300)))
def test_rollout_vs_oracle_replay(eng, n, boards, steps):
    from test_device_algo_hostsim import philox_numpy
    e = eng(n)
    rec = e.new_records(boards)
    dense = np.zeros((boards, 6, n, n), dtype=np.uint8)
    acts_t = e.empty((boards,), dtype=torch.int32)
    obs = e.empty((boards, 6, n, n), dtype=torch.uint8)
    done_t = e.empty((boards,))
    seed, board0 = 0xC0FFEE12345, 777
    finished = 0
    for t in range(steps):
        e.rollout_step(rec, seed, board0, t, actions=acts_t, obs=obs, done=done_t)
        acts = acts_t.cpu().numpy()
        was_done = dense[:, 5, 0, 0] == 1
        finished += int(was_done.sum())
        dense[was_done] = 0
        rnd = philox_numpy(np.arange(boards, dtype=np.uint64) + np.uint64(board0), t, seed)
        valid = np.concatenate([1 - dense[:, 3].reshape(boards, -1), np.ones((boards, 1), dtype=np.uint8)], axis=1)
        k = ((rnd * valid.sum(axis=1).astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
        want_act = np.array([np.flatnonzero(valid[i])[k[i]] for i in range(boards)], dtype=np.int32)
        assert np.array_equal(acts, want_act), t
        dense, status = co.batch_next_states(dense, acts)
        assert not status.any()
        assert np.array_equal(obs.cpu().numpy(), dense), t
        assert np.array_equal(done_t.cpu().numpy(), dense[:, 5, 0, 0])
    assert np.array_equal(e.unpack(rec, dtype=torch.uint8).cpu().numpy(), dense)
    if n <= 9:
        assert finished > 0


@pytest.mark.parametrize("n,boards", ((9, 1003), (19, 149), (7, 333), (5, 77), (17, 41), (13, 64)))
@pytest.mark.parametrize("dtype", (torch.float32, torch.uint8, torch.bfloat16))
def test_persistent_rollout_equals_single_plies(eng, n, boards, dtype):
    """gg_rollout (boards resident in registers for several plies, per-warp observation emission incl. ragged
    tiles and unaligned warp slices) reproduces the ply-by-ply kernel bit for bit: records, observations of every
    ply, actions, done flags, rewards."""
    from gymgo_b200 import _cabi
    e = eng(n)
    steps, ppl, seed, board0, t0 = 37, 5, 99, 1234, 17
    warm = e.new_records(boards)
    for t in range(t0):
        e.rollout_step(warm, seed, board0, t)
    a = warm.clone()
    b = warm.clone()
    ring = e.empty((steps + 1, boards, 6, n, n), dtype=dtype)
    ring.fill_(7)
    acts = torch.full((steps, boards), -5, dtype=torch.int32, device="cuda")
    dones = torch.full((steps, boards), 9, dtype=torch.uint8, device="cuda")
    rews = torch.full((steps, boards), 9.0, dtype=torch.float32, device="cuda")
    e.rollout(a, seed, board0, t0, steps, plies_per_launch=ppl, actions_log=acts, obs_ring=ring, done_log=dones,
              reward_log=rews, reward_mode=_cabi.GG_REWARD_HEURISTIC, komi=0.5)
    obs1 = e.empty((boards, 6, n, n), dtype=dtype)
    a1 = e.empty((boards,), dtype=torch.int32)
    d1 = e.empty((boards,))
    r1 = e.empty((boards,), dtype=torch.float32)
    for p in range(steps):
        t = t0 + p
        e.rollout_step(b, seed, board0, t, actions=a1, obs=obs1, done=d1, reward=r1,
                       reward_mode=_cabi.GG_REWARD_HEURISTIC, komi=0.5)
        assert torch.equal(ring[t % (steps + 1)], obs1), (p, "obs")
        assert torch.equal(acts[p], a1) and torch.equal(dones[p], d1) and torch.equal(rews[p], r1), p
    assert torch.equal(a, b)
    untouched = [s for s in range(steps + 1) if s not in {(t0 + p) % (steps + 1) for p in range(steps)}]
    assert all(bool((ring[s] == 7).all()) for s in untouched)


@pytest.mark.parametrize("n,boards", ((9, 1003), (7, 333), (5, 77), (3, 130), (8, 64), (19, 141), (13, 97), (16, 33), (2, 50)))
@pytest.mark.parametrize("dtype", (torch.float32, torch.uint8, torch.float16))
def test_rollout_kernels_are_bit_identical(eng, n, boards, dtype):
    """the two persistent rollout kernels - lane-sliced and thread-per-board - with static (one CTA per tile) and
    dynamic ((tile, 8-ply block) work items with tickets) scheduling produce the same records, observations and logs"""
    from gymgo_b200 import _cabi
    e = eng(n)
    outs = []
    for kernel, dynamic in ((_cabi.GG_KERNEL_LANES, False), (_cabi.GG_KERNEL_THREAD, False),
                            (_cabi.GG_KERNEL_LANES, True), (_cabi.GG_KERNEL_THREAD, True)):
        rec = e.new_records(boards)
        ring = e.empty((23, boards, 6, n, n), dtype=dtype)
        ring.fill_(3)
        acts = torch.empty((22, boards), dtype=torch.int32, device="cuda")
        dones = torch.empty((22, boards), dtype=torch.uint8, device="cuda")
        rews = torch.empty((22, boards), dtype=torch.float32, device="cuda")
        e.rollout(rec, 5, 77, 0, 150, plies_per_launch=50, kernel=kernel, dynamic=dynamic)     # 7 blocks when dynamic
        e.rollout(rec, 5, 77, 150, 22, plies_per_launch=22 if dynamic else 6, actions_log=acts, obs_ring=ring,
                  done_log=dones, reward_log=rews, reward_mode=2, komi=1.5, kernel=kernel, dynamic=dynamic)
        torch.cuda.synchronize()
        outs.append((rec, ring, acts, dones, rews))
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert torch.equal(x, y)


@pytest.mark.parametrize("kernel", (0, 1))
@pytest.mark.parametrize("n,boards,ppl,ring", ((19, 1000, 7, 3), (19, 64, 32, 32), (13, 500, 5, 2), (9, 2000, 9, 4), (6, 77, 3, 1),
                                               (9, 5000, 40, 3), (19, 700, 33, 2), (4, 300, 64, 5)))
def test_rollout_small_rings_keep_the_last_ply(eng, n, boards, ppl, ring, kernel):
    """many plies per launch into rings smaller than a launch: every ring slot ends up holding the observation of the
    LAST ply mapped to it (what the single-ply kernel writes), for both persistent kernels"""
    e = eng(n)
    a, b = e.new_records(boards), e.new_records(boards)
    obs_ring = e.empty((ring, boards, 6, n, n), dtype=torch.float32)
    one = e.empty((boards, 6, n, n), dtype=torch.float32)
    t = 0
    for launch in range(6):
        e.rollout(a, 11, 5, t, ppl, plies_per_launch=ppl, obs_ring=obs_ring, kernel=kernel)
        last = {}
        for p in range(ppl):
            e.rollout_step(b, 11, 5, t + p, obs=one)
            last[(t + p) % ring] = one.clone()
        for slot, want in last.items():
            assert torch.equal(obs_ring[slot], want), (launch, slot)
        t += ppl
    assert torch.equal(a, b)


@pytest.mark.parametrize("n", golden_io.CHILDREN_SIZES)
def test_children_golden(eng, n):
    e = eng(n)
    P, C0, C1 = golden_io.children(n)
    rec = e.pack(dev(P))
    for canon, ref in ((False, C0), (True, C1)):
        res = e.children(rec, canonical=canon, obs_dtype=torch.float32, want_rec=True)
        assert not res["status"].any()
        assert np.array_equal(res["obs"].cpu().numpy(), ref.astype(np.float32))
        a = n * n + 1
        flat = res["rec"].reshape(-1, e.rec_bytes)
        assert np.array_equal(e.unpack(flat, dtype=torch.uint8).cpu().numpy().reshape(len(P), a, 6, n, n),
                              ref.astype(np.uint8))
        want_valid = np.stack([co.valid_moves(p) for p in P])
        assert np.array_equal(res["valid"].cpu().numpy(), want_valid)


@pytest.mark.parametrize("n", (4, 9, 19))
def test_children_vs_oracle_random(eng, n):
    e = eng(n)
    st = soup(n, 61, 4242 + n)
    st[:, 5] = 0
    res = e.children(e.pack(dev(st)), canonical=False, obs_dtype=torch.uint8, want_rec=False)
    for i in range(len(st)):
        kids, valid, bad = co.children(st[i])
        assert not bad and not int(res["status"][i])
        assert np.array_equal(res["valid"][i].cpu().numpy(), valid)
        assert np.array_equal(res["obs"][i].cpu().numpy(), kids)
    # finished parent with stones: the reference asserts (gogame.py:117) -> status 1
    st[:, 5] = 1
    res = e.children(e.pack(dev(st)), obs_dtype=torch.uint8, want_rec=False)
    has_stones = (st[:, 3].reshape(len(st), -1).sum(axis=1) > 0)
    assert np.array_equal(res["status"].cpu().numpy().astype(bool), has_stones)


def test_canonical_and_valid_variants(eng):
    e = eng(7)
    st = soup(7, 300, 11)
    rec = e.pack(dev(st))
    can = e.unpack(e.canonical(rec), dtype=torch.uint8).cpu().numpy()
    want = st.copy()
    w = st[:, 2, 0, 0] == 1
    want[w, 0], want[w, 1] = st[w, 1], st[w, 0]
    want[w, 2] = 0
    assert np.array_equal(can, want)
    rec2 = rec.clone()
    e.canonical(rec2, out=rec2)                                     # in place
    assert np.array_equal(e.unpack(rec2, dtype=torch.uint8).cpu().numpy(), want)
    vq = e.valid_moves(rec, ended_quirk=True, dtype=torch.uint8).cpu().numpy()
    vb = e.valid_moves(rec, ended_quirk=False, dtype=torch.uint8).cpu().numpy()
    done = st[:, 5, 0, 0] == 1
    base = np.concatenate([1 - st[:, 3].reshape(300, -1), np.ones((300, 1), dtype=np.uint8)], axis=1)
    assert np.array_equal(vb, base)
    assert (vq[done] == 1).all() and np.array_equal(vq[~done], base[~done])


def test_rewards_epilogue(eng):
    from gymgo_b200 import _cabi
    e = eng(5)
    st = soup(5, 500, 5)
    st[:, 5] = 0
    st[:, 4] = (np.arange(500) % 2)[:, None, None]          # half the boards: previous move was a pass
    rec = e.pack(dev(st))
    acts = np.full(500, 25, dtype=np.int32)                  # everybody passes -> half the games end
    for mode, komi in ((_cabi.GG_REWARD_REAL, 0.0), (_cabi.GG_REWARD_REAL, 2.5), (_cabi.GG_REWARD_HEURISTIC, 0.0),
                       (_cabi.GG_REWARD_HEURISTIC, 1.5)):
        res = e.step(rec, acts, want_done=True, reward_mode=mode, komi=komi)
        nxt, _ = co.batch_next_states(st, acts)
        ar = co.batch_areas(nxt).astype(np.float64)
        diff = ar[:, 0] - ar[:, 1] - komi
        over = nxt[:, 5, 0, 0] == 1
        if mode == _cabi.GG_REWARD_REAL:
            want = np.where(over, np.sign(diff), 0.0)
        else:
            want = np.where(over, np.where(diff > 0, 25.0, -25.0), diff)
        assert over.any() and (~over).any()
        assert np.array_equal(res["reward"].cpu().numpy().astype(np.float64), want)


def test_error_codes(eng):
    from gymgo_b200 import _cabi
    L = _cabi.lib()
    assert L.gg_supported(9) == 1 and L.gg_supported(1) == 0 and L.gg_supported(20) == 0
    e = eng(9)
    rec = e.new_records(4)
    a = torch.zeros(4, dtype=torch.int32, device="cuda")
    s = e._enter()
    assert L.gg_step(rec.data_ptr(), a.data_ptr(), rec.data_ptr(), None, 4, 20, 0, None, 0, None, None, None, 0, 0.0, s) == _cabi.GG_ESIZE
    assert L.gg_step(rec.data_ptr() + 4, a.data_ptr(), rec.data_ptr(), None, 4, 9, 0, None, 0, None, None, None, 0, 0.0, s) == _cabi.GG_EALIGN
    assert L.gg_step(None, a.data_ptr(), rec.data_ptr(), None, 4, 9, 0, None, 0, None, None, None, 0, 0.0, s) == _cabi.GG_EINVAL
    assert L.gg_step(rec.data_ptr(), a.data_ptr(), rec.data_ptr(), None, 4, 9, 64, None, 0, None, None, None, 0, 0.0, s) == _cabi.GG_EINVAL
    assert L.gg_step(rec.data_ptr(), a.data_ptr(), rec.data_ptr(), None, 0, 9, 0, None, 0, None, None, None, 0, 0.0, s) == _cabi.GG_OK
