"""GPU tests at BASELINE.json's full sizes.  The oracle cannot replay 65,536 boards ply by ply in seconds, so
these use (a) the host simulator of the device algorithm on the whole batch (C++, bit-identical records and
actions expected), (b) size-independent invariants, (c) the C oracle on a strided sample of transitions."""
import numpy as np
import pytest
import torch

import hostsim
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def invariants(dense):
    black, white, invd = dense[:, 0], dense[:, 1], dense[:, 3]
    assert not (black & white).any()                                   # a point holds one stone at most
    assert ((black | white) <= invd).all()                             # occupied points are invalid
    for ch in (2, 4, 5):                                               # whole-plane facts are constant planes
        p = dense[:, ch].reshape(len(dense), -1)
        assert (p.min(axis=1) == p.max(axis=1)).all()


@pytest.mark.parametrize("n,boards,plies,ppl", ((9, 65536, 48, 8), (19, 16384, 40, 5)))
def test_full_batch_rollout_equals_host_simulation(n, boards, plies, ppl):
    from gymgo_b200.engine import GoEngine
    e = GoEngine(n, "cuda:0")
    rec = e.new_records(boards)
    acts = torch.empty((plies, boards), dtype=torch.int32, device="cuda")
    ring = e.empty((2, boards, 6, n, n), dtype=torch.uint8)
    e.rollout(rec, 11, 5_000_000_000, 0, plies, plies_per_launch=ppl, actions_log=acts, obs_ring=ring)
    sim = hostsim.pack(np.zeros((boards, 6, n, n), dtype=np.uint8))
    sim_acts = np.stack([hostsim.rollout_step(sim, n, 11, 5_000_000_000, t) for t in range(plies)])
    assert np.array_equal(acts.cpu().numpy(), sim_acts)
    assert np.array_equal(rec.cpu().numpy().view(np.uint32), sim)
    last = ring[(plies - 1) % 2].cpu().numpy()
    assert np.array_equal(last, hostsim.unpack(sim, n))
    invariants(last)
    # oracle on a strided sample of the final transition
    idx = np.arange(0, boards, 97)
    prev = e.new_records(boards)
    e.rollout(prev, 11, 5_000_000_000, 0, plies - 1, plies_per_launch=ppl)
    before = e.unpack(prev, dtype=torch.uint8).cpu().numpy()[idx]
    before[before[:, 5, 0, 0] == 1] = 0                                # auto-reset precedes the ply
    want, status = co.batch_next_states(before, sim_acts[-1][idx])
    assert not status.any() and np.array_equal(want, last[idx])


@pytest.mark.parametrize("n,boards,kernel_name", ((9, 65536, "k_rollout_tpb"), (19, 16384, "k_rollout (")))
def test_deep_rollout_soak_against_c_oracle(n, boards, kernel_name):
    """The HEADLINE kernels at the headline batch sizes and launch shape (32 plies per launch into a 32-slot float32
    ring, like bench.py): 320 plies from empty boards (9x9 games last ~124 plies, so boards finish, restart and
    de-synchronise; 19x19 boards fill up to the mid-game), with a strided 1/64 sample of the boards replayed through
    the C oracle at EVERY ply - next state bit for bit, legality of the sampled action, done flag, REAL reward."""
    from gymgo_b200.engine import GoEngine
    e = GoEngine(n, "cuda:0")
    assert e.lib.gg_rollout_kernel(n, boards).decode().startswith(kernel_name)
    ppl, launches, seed, board0 = 32, 10, 2024, 123_456_789
    idx = torch.arange(0, boards, 64, device="cuda")
    rec = e.new_records(boards)
    ring = e.empty((ppl, boards, 6, n, n), dtype=torch.float32)
    acts = torch.empty((ppl, boards), dtype=torch.int32, device="cuda")
    dones = torch.empty((ppl, boards), dtype=torch.uint8, device="cuda")
    rews = torch.empty((ppl, boards), dtype=torch.float32, device="cuda")
    prev = np.zeros((len(idx), 6, n, n), dtype=np.uint8)
    checked = ended = 0
    for launch in range(launches):
        e.rollout(rec, seed, board0, launch * ppl, ppl, plies_per_launch=ppl, actions_log=acts, obs_ring=ring,
                  done_log=dones, reward_log=rews, reward_mode=1, komi=0.5)
        obs = ring[:, idx].to(torch.uint8).cpu().numpy()               # ply p of this launch sits in slot p (32 | t0)
        a, d, r = acts[:, idx].cpu().numpy(), dones[:, idx].cpu().numpy(), rews[:, idx].cpu().numpy()
        for p in range(ppl):
            prev[prev[:, 5, 0, 0] == 1] = 0                             # auto-reset precedes the ply
            want, status = co.batch_next_states(prev, a[p])
            assert not status.any(), (launch, p, "sampled an illegal action")
            assert np.array_equal(want, obs[p]), (launch, p)
            over = want[:, 5, 0, 0] == 1
            assert np.array_equal(d[p], over.astype(np.uint8))
            ar = co.batch_areas(want).astype(np.float64)
            want_r = np.where(over, np.sign(ar[:, 0] - ar[:, 1] - 0.5), 0.0)
            assert np.array_equal(r[p].astype(np.float64), want_r)
            prev = want
            checked += len(idx)
            ended += int(over.sum())
    assert checked == launches * ppl * len(idx) and (ended > 0 or n == 19)
    # the records the kernel left behind are the sampled boards' last states
    assert np.array_equal(e.unpack(rec[idx], dtype=torch.uint8).cpu().numpy(), prev)
    invariants(e.unpack(rec, dtype=torch.uint8).cpu().numpy())


def test_headline_batch_step_kernel_choice_against_c_oracle():
    """BatchedGoEnv.step at the headline batch (9x9 x 65,536: gg_step picks k_step_tpb by itself) along 200 plies of a
    game with auto-reset: the same records as the lane-sliced kernel driven with the same actions, and a strided sample
    of every ply replayed through the C oracle (next state, status, done)"""
    from gymgo_b200.engine import GoEngine
    n, boards = 9, 65536
    e = GoEngine(n, "cuda:0")
    auto, lanes = e.new_records(boards), e.new_records(boards)
    obs = e.empty((boards, 6, n, n), dtype=torch.uint8)
    sample = torch.arange(0, boards, 128, device="cuda")
    for t in range(200):
        acts = e.sample_legal(auto, 21, 0, t)
        if t % 7 == 3:
            acts[::50] = n * n + 5                                           # a few out-of-range actions
        before = e.unpack(auto[sample], dtype=torch.uint8).cpu().numpy()
        ra = e.step(auto, acts, out=auto, obs=obs, auto_reset=True, want_done=True)
        rl = e.step(lanes, acts, out=lanes, auto_reset=True, want_done=True, kernel="lanes")
        assert torch.equal(auto, lanes) and torch.equal(ra["status"], rl["status"]) and torch.equal(ra["done"], rl["done"])
        before[before[:, 5, 0, 0] == 1] = 0                                   # auto-reset precedes the ply
        want, wstatus = co.batch_next_states(before, acts[sample].cpu().numpy())
        assert np.array_equal(obs[sample].cpu().numpy(), want), t
        assert np.array_equal(ra["status"][sample].cpu().numpy(), wstatus), t
    assert bool((e.flags(auto) & 4).any())


def test_children_full_config():
    """configs[3]: 9x9 children() of 4,096 parents after 40 random plies; every slot against per-action gg_step,
    a sample against the C oracle."""
    from gymgo_b200.engine import GoEngine
    e = GoEngine(9, "cuda:0")
    parents = e.new_records(4096)
    e.rollout(parents, 0, 0, 0, 40, plies_per_launch=8)
    e.reset(parents, e.flags(parents) & 4 != 0)                        # drop the few finished games
    res = e.children(parents, obs_dtype=torch.uint8, want_rec=True)
    assert not res["status"].any()
    valid = res["valid"].cpu().numpy()
    dense_parents = e.unpack(parents, dtype=torch.uint8).cpu().numpy()
    assert np.array_equal(valid[:, :81], 1 - dense_parents[:, 3].reshape(4096, 81)) and (valid[:, 81] == 1).all()
    for a in (0, 17, 40, 80, 81):                                      # whole columns of the expansion vs gg_step
        acts = torch.full((4096,), a, dtype=torch.int32, device="cuda")
        step = e.step(parents, acts, obs_dtype=torch.uint8)
        ok = step["status"].cpu().numpy() == 0
        assert np.array_equal(ok, valid[:, a] == 1)
        assert np.array_equal(res["obs"][:, a].cpu().numpy()[ok], step["obs"].cpu().numpy()[ok])
        assert not res["obs"][:, a].cpu().numpy()[~ok].any()           # padded zeros
        assert torch.equal(res["rec"][:, a][torch.from_numpy(ok).cuda()], step["rec"][torch.from_numpy(ok).cuda()])
    for i in range(0, 4096, 311):
        kids, v, bad = co.children(dense_parents[i])
        assert not bad and np.array_equal(res["obs"][i].cpu().numpy(), kids)


def test_shard_invariance_full_size():
    """two half batches with board offsets reproduce the whole batch (what the multi-GPU run relies on)"""
    from gymgo_b200.engine import GoEngine
    e = GoEngine(9, "cuda:0")
    whole, lo, hi = e.new_records(8192), e.new_records(4096), e.new_records(4096)
    e.rollout(whole, 3, 0, 0, 64, plies_per_launch=8)
    e.rollout(lo, 3, 0, 0, 64, plies_per_launch=4)
    e.rollout(hi, 3, 4096, 0, 64, plies_per_launch=16)
    assert torch.equal(whole[:4096], lo) and torch.equal(whole[4096:], hi)
