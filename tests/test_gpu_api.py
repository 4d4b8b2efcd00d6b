"""GPU tests of the reference-facing Python surface: GoEnv / gogame drop-ins replaying the reference's own
test sequences (SURVEY.md Appendix B fixtures) and BASELINE.json configs[0] (7x7 GoEnv.step games)."""
import numpy as np
import pytest
import torch

import golden_io
from oracle import gogame_np as og

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_io.kat_cases(), ids=lambda c: c["name"])
def test_goenv_kat(case):
    import gymgo_b200
    n = case["states"].shape[2]
    env = gymgo_b200.make("gym_go:go-v0", size=n, komi=case["komi"], reward_method=case["method"])
    st = env.reset()
    assert st.dtype == np.float64 and np.array_equal(st, case["states"][0])
    for i, a in enumerate(case["actions"]):
        a = int(a)
        move = None if a == n * n else ((a // n, a % n) if i % 2 else a)       # exercise all action formats
        st, rew, done, info = env.step(move)
        assert np.array_equal(st, case["states"][i + 1]), (case["name"], i)
        assert float(rew) == case["rewards"][i]
        assert isinstance(done, int) and done == case["dones"][i]
        assert info["turn"] == case["turns"][i]
        assert int(info["prev_player_passed"]) == case["prev_pass"][i]
        assert np.array_equal(info["invalid_moves"], og.invalid_moves(case["states"][i + 1]))
    if case["raises"] >= 0:
        with pytest.raises(AssertionError):
            env.step(case["raises"])
        assert np.array_equal(env.state(), case["states"][-1])      # a refused move changes nothing


def test_goenv_full_games_7x7():
    """configs[0]: state-for-state equality of whole 7x7 games incl. rewards for both reward methods"""
    from gymgo_b200.envs import GoEnv
    for g in golden_io.env_games():
        env = GoEnv(7, komi=g["komi"], reward_method=g["method"])
        env.reset()
        for i, a in enumerate(g["actions"]):
            st, rew, done, _ = env.step(int(a))
            assert np.array_equal(st, g["states"][i + 1])
            assert float(rew) == g["rewards"][i]
            assert done == g["dones"][i]
        assert float(env.winning()) == g["winning"]
        with pytest.raises(AssertionError):
            env.step(None)


def test_goenv_misc_surface():
    from gymgo_b200.envs import GoEnv
    env = GoEnv(5)
    with pytest.raises(AssertionError):
        env.step((5, 0))                                            # out of bounds (test_invalid_moves.py:19-24)
    env.step((1, 1))
    env.step(None)
    assert env.turn() == 0 and env.prev_player_passed()
    kids = env.children(canonical=True, padded=True)
    assert kids.shape == (26, 6, 5, 5) and kids.dtype == np.float64
    assert np.array_equal(kids, og.children(env.state(), canonical=True, padded=True))
    assert np.array_equal(env.children(padded=False), og.children(env.state(), padded=False))
    assert np.array_equal(env.canonical_state(), og.canonical_form(env.state()))
    a = env.uniform_random_action()
    assert env.valid_moves()[a] == 1
    assert "Turn" in str(env)


def test_gogame_functional_api_numpy_and_torch():
    from gym_go import gogame, govars      # the reference's import path
    S0, A, S1 = golden_io.transitions(9)
    S0, A, S1 = S0[:200], A[:200], S1[:200]
    keep = S0.copy()
    out = gogame.batch_next_states(S0, A)
    assert out.dtype == np.float64 and np.array_equal(out, S1) and np.array_equal(S0, keep)
    assert np.array_equal(gogame.next_state(S0[3], int(A[3]), canonical=True), og.next_state(S0[3], int(A[3]), True))
    with pytest.raises(AssertionError):
        occupied = int(np.flatnonzero(S1[5][govars.INVD_CHNL].reshape(-1))[0])
        gogame.next_state(S1[5], occupied)
    assert np.array_equal(gogame.valid_moves(S1[7]), og.valid_moves(S1[7]))
    assert np.array_equal(gogame.invalid_moves(S1[7]), og.invalid_moves(S1[7]))
    assert np.array_equal(gogame.batch_valid_moves(S1), og.batch_valid_moves(S1))
    assert np.array_equal(gogame.children(S1[9]), og.children(S1[9]))
    assert tuple(gogame.areas(S1[11])) == tuple(og.areas(S1[11]))
    b, w = gogame.batch_areas(S1)
    ob, ow = og.batch_areas(S1)
    assert np.array_equal(b, ob) and np.array_equal(w, ow)
    assert gogame.winning(S1[11], 2.5) == og.winning(S1[11], 2.5)
    assert np.array_equal(gogame.batch_canonical_form(S1), og.batch_canonical_form(S1))
    assert gogame.turn(S1[0]) == og.turn(S1[0]) and gogame.game_ended(S1[0]) == 0
    assert gogame.num_liberties(S1[20]) is not None and gogame.action_size(S1[0]) == 82
    # torch tensors stay on the device
    t = torch.from_numpy(S0).cuda().float()
    tout = gogame.batch_next_states(t, torch.from_numpy(A).cuda())
    assert tout.is_cuda and tout.dtype == torch.float32 and np.array_equal(tout.cpu().numpy(), S1.astype(np.float32))


def test_batched_env_semantics():
    from gymgo_b200.envs import BatchedGoEnv
    env = BatchedGoEnv(64, 5, reward_method="heuristic", komi=0.5, strict=False)
    ref = [og.EnvOracle(5, komi=0.5, reward_method="heuristic") for _ in range(64)]
    rng = np.random.RandomState(0)
    for t in range(60):
        vm = env.valid_moves(dtype=torch.uint8).cpu().numpy()
        acts = np.array([rng.choice(np.flatnonzero(vm[i])) for i in range(64)], dtype=np.int32)
        live = np.array([not r.done for r in ref])
        obs, rew, done, info = env.step(acts)
        st = info["status"].cpu().numpy()
        assert (st[~live] == 3).all() and (st[live] == 0).all()
        for i in np.flatnonzero(live):
            s, r, d, _ = ref[i].step(int(acts[i]))
            assert np.array_equal(obs[i].cpu().numpy(), s.astype(np.float32))
            assert float(rew[i]) == float(r) and int(done[i]) == int(d)
    strict = BatchedGoEnv(4, 5, strict=True)
    strict.step([0, 1, 2, 3])
    with pytest.raises(AssertionError):
        strict.step([0, 6, 7, 8])


def test_batched_env_random_step_is_shard_invariant():
    from gymgo_b200.envs import BatchedGoEnv
    whole = BatchedGoEnv(96, 9, seed=5, board_offset=0, obs_dtype=torch.uint8)
    lo = BatchedGoEnv(48, 9, seed=5, board_offset=0, obs_dtype=torch.uint8)
    hi = BatchedGoEnv(48, 9, seed=5, board_offset=48, obs_dtype=torch.uint8)
    for _ in range(150):
        ow, _, _, aw = whole.random_step()
        ol, _, _, al = lo.random_step()
        oh, _, _, ah = hi.random_step()
        assert torch.equal(ow[:48], ol) and torch.equal(ow[48:], oh)
        assert torch.equal(aw[:48], al) and torch.equal(aw[48:], ah)


def test_state_utils_and_symmetries():
    from gym_go import gogame, state_utils
    from test_device_algo_hostsim import random_soup
    st = random_soup(7, 40, np.random.RandomState(3)).astype(np.float64)
    for i in range(len(st)):
        player = int(i % 2)
        ko = None if i % 3 else (i % 7, (3 * i) % 7)
        assert np.array_equal(state_utils.compute_invalid_moves(st[i], player, ko), og.invalid_mask(st[i], player, ko))
    s = st[0].copy()
    state_utils.set_turn(s)
    assert (s[2] == 1 - st[0][2]).all()
    nb, sur = state_utils.adj_data(st[1], (0, 0), 0)
    assert len(nb) == 2
    img = np.arange(2 * 6 * 5 * 5).reshape(2, 6, 5, 5)
    sym_np = gogame.all_symmetries(img)
    sym_t = gogame.all_symmetries(torch.from_numpy(img).cuda())
    assert len(sym_np) == 8 and all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(sym_np, sym_t))
    assert len({a.tobytes() for a in sym_np}) == 8


def test_cuda_graph_capture_of_step_calls():
    """the C ABI only enqueues kernels on the caller's stream, so a sequence of plies can be captured into a CUDA
    graph and replayed"""
    from gymgo_b200.engine import GoEngine
    e = GoEngine(9, "cuda:0")
    start = e.new_records(2048)
    e.rollout(start, 1, 0, 0, 30)
    acts = [e.empty((2048,), dtype=torch.int32) for _ in range(4)]
    obs = e.empty((2048, 6, 9, 9), dtype=torch.float32)
    rec = start.clone()

    def four_plies():
        for k in range(4):
            e.lib.gg_sample_legal(rec.data_ptr(), 2048, 9, 5, 0, k, acts[k].data_ptr(), torch.cuda.current_stream().cuda_stream)
            e.lib.gg_step(rec.data_ptr(), acts[k].data_ptr(), rec.data_ptr(), None, 2048, 9, 0, obs.data_ptr(), 1, None,
                          None, None, 0, 0.0, torch.cuda.current_stream().cuda_stream)

    e._enter()
    four_plies()
    torch.cuda.synchronize()
    want_rec, want_obs = rec.clone(), obs.clone()
    rec.copy_(start)
    obs.zero_()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            four_plies()
    rec.copy_(start)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(rec, want_rec) and torch.equal(obs, want_obs)


@pytest.mark.parametrize("mode", ("next_step", "play_on_reset"))
def test_vector_env_autoreset_and_masks(mode):
    from gymgo_b200.envs import GoVectorEnv
    venv = GoVectorEnv(128, 5, reward_method="real", obs_dtype=torch.uint8, seed=4, autoreset_mode=mode)
    obs, info = venv.reset()
    assert obs.shape == (128, 6, 5, 5) and not obs.any() and info["action_mask"].all()
    ref = [og.EnvOracle(5) for _ in range(128)]
    finished_before = np.zeros(128, dtype=bool)
    ended = 0
    for t in range(80):
        acts = venv.sample_actions()
        a = acts.cpu().numpy()
        obs, rew, term, trunc, info = venv.step(acts)
        assert not trunc.any() and not info["status"].any()
        for i in range(128):
            if finished_before[i]:
                s, r, d = ref[i].reset(), 0, False
                if mode == "play_on_reset":                     # the action is played on the fresh board
                    s, r, d, _ = ref[i].step(int(a[i]))
            else:
                s, r, d, _ = ref[i].step(int(a[i]))
            assert np.array_equal(obs[i].cpu().numpy(), s.astype(np.uint8))
            assert float(rew[i]) == float(r) and bool(term[i]) == bool(d)
            assert np.array_equal(info["action_mask"][i].cpu().numpy(), og.valid_moves(s).astype(np.uint8))
            finished_before[i] = bool(d)
        ended += int(term.sum())
    assert ended > 0
    # copy=True hands out tensors that survive the next step
    keep = GoVectorEnv(16, 5, seed=1, copy=True)
    keep.reset()
    o1, r1, *_ = keep.step(keep.sample_actions())
    snap = o1.clone()
    keep.step(keep.sample_actions())
    assert torch.equal(o1, snap)


def test_step_auto_reset_flag_equals_reset_then_step():
    """GG_STEP_AUTO_RESET inside gg_step == gg_reset(mask=done) followed by gg_step (the two-launch loop of round 1)"""
    from gymgo_b200.engine import GoEngine
    e = GoEngine(9, "cuda:0")
    a, b = e.new_records(5000), e.new_records(5000)
    done = torch.zeros(5000, dtype=torch.uint8, device="cuda")
    for t in range(200):
        acts = e.sample_legal(a, 3, 0, t)
        acts = torch.where(torch.arange(5000, device="cuda") % 4 == 0, torch.full_like(acts, 81), acts)   # many passes
        e.reset(b, done)
        rb = e.step(b, acts, out=b, refuse_done=True, obs_dtype=torch.uint8, want_done=True, reward_mode=1)
        ra = e.step(a, acts, out=a, refuse_done=True, auto_reset=True, obs_dtype=torch.uint8, want_done=True, reward_mode=1)
        done = rb["done"]
        for k in ("status", "obs", "done", "reward"):
            assert torch.equal(ra[k], rb[k]), (t, k)
        assert torch.equal(a, b)
    assert bool(done.any())
    with pytest.raises(ValueError):
        e.reset(a, torch.zeros(17, dtype=torch.uint8))              # mask must be [B]
    with pytest.raises(ValueError):
        e.unpack(a, out=torch.empty((5, 6, 9, 9), device="cuda"))   # out must be [B,6,N,N]


@pytest.mark.parametrize("returns,graph", (("obs", True), ("obs", False), ("packed", True), ("none", True)))
def test_host_stepper_matches_device_step(returns, graph):
    """the host-buffer step (pinned actions in, pinned results out, one CUDA graph) == BatchedGoEnv.step"""
    from gymgo_b200.envs import BatchedGoEnv
    env = BatchedGoEnv(3000, 9, reward_method="heuristic", komi=0.5, obs_dtype=torch.float32)
    ref = BatchedGoEnv(3000, 9, reward_method="heuristic", komi=0.5, obs_dtype=torch.float32)
    hs = env.host_stepper(returns=returns, use_cuda_graph=graph, transport="dense")
    assert hs.h2d_bytes == 3000 * 4
    assert hs.d2h_bytes == 3000 * 5 + {"obs": 3000 * 6 * 81 * 4, "packed": 3000 * env.engine.rec_bytes, "none": 0}[returns]
    for t in range(140):
        acts = ref.engine.sample_legal(ref.rec, 9, 0, t)
        hs.actions.copy_(acts.cpu())
        first, rew, done = hs.step()
        o, r, d, _ = ref.step(acts, auto_reset=True)
        assert not first.is_cuda if first is not None else True
        assert torch.equal(rew, r.cpu()) and torch.equal(done, d.cpu())
        if returns == "obs":
            assert torch.equal(first, o.cpu())
        elif returns == "packed":
            assert torch.equal(first, ref.rec.cpu())
            assert torch.equal(hs.expand(), o.cpu())
            assert torch.equal(hs.expand(dtype=torch.uint8, threads=3), o.cpu().to(torch.uint8))
    assert torch.equal(env.rec, ref.rec) and bool(ref.done.any())


@pytest.mark.parametrize("n,dtype", ((9, torch.float32), (19, torch.float32), (9, torch.bfloat16), (7, torch.uint8)))
def test_host_stepper_packed_transport_returns_the_same_observation(n, dtype):
    """host_stepper(transport="packed"): records over PCIe + gg_host_unpack on the host cores == the dense observation the
    kernel writes (transport="dense"), bit for bit, with 40x fewer bytes crossing the bus"""
    from gymgo_b200.envs import BatchedGoEnv
    b = 2500                                                        # ragged: not a multiple of the codec's 32-board chunk
    env = BatchedGoEnv(b, n, obs_dtype=dtype)
    ref = BatchedGoEnv(b, n, obs_dtype=dtype)
    hs = env.host_stepper(returns="obs", transport="packed", threads=3)
    assert hs.d2h_bytes == b * 5 + b * env.engine.rec_bytes and hs.host_expanded_bytes == b * 6 * n * n * env.obs.element_size()
    for t in range(60):
        acts = ref.engine.sample_legal(ref.rec, 4, 0, t)
        hs.actions.copy_(acts.cpu())
        obs, rew, done = hs.step()
        o, r, d, _ = ref.step(acts, auto_reset=True)
        assert not obs.is_cuda and obs.dtype == dtype
        assert torch.equal(obs, o.cpu()) and torch.equal(rew, r.cpu()) and torch.equal(done, d.cpu())
    assert torch.equal(env.rec, ref.rec)
    auto = env.host_stepper(returns="obs")                           # the default picks one of the two, never anything else
    assert auto.transport in ("dense", "packed") and env.host_stepper(returns="none").transport == "dense"
    with pytest.raises(ValueError):
        env.host_stepper(returns="obs", transport="carrier pigeon")


@pytest.mark.parametrize("n", (5, 9, 13, 19))
def test_packed_symmetries_match_numpy(n):
    from gymgo_b200 import gogame
    from gymgo_b200.engine import GoEngine
    from test_device_algo_hostsim import random_soup
    e = GoEngine(n, "cuda:0")
    st = random_soup(n, 50, np.random.RandomState(n))
    rec = e.pack(torch.from_numpy(st).cuda())
    want = gogame.all_symmetries(st)
    for sym in range(8):
        got = e.unpack(e.symmetry(rec, sym), dtype=torch.uint8).cpu().numpy()
        assert np.array_equal(got, want[sym]), sym


def test_batched_env_cuda_graph_step_equals_eager():
    from gymgo_b200.envs import BatchedGoEnv
    eager = BatchedGoEnv(3000, 9, reward_method="heuristic", komi=0.5, obs_dtype=torch.uint8)
    graph = BatchedGoEnv(3000, 9, reward_method="heuristic", komi=0.5, obs_dtype=torch.uint8, use_cuda_graph=True)
    for t in range(150):
        acts = eager.engine.sample_legal(eager.rec, 9, 0, t)
        reset = t % 3 != 0
        oe, re_, de, ie = eager.step(acts, auto_reset=reset)
        og_, rg, dg, ig = graph.step(acts.clone(), auto_reset=reset)
        assert torch.equal(oe, og_) and torch.equal(re_, rg) and torch.equal(de, dg)
        assert torch.equal(ie["status"], ig["status"])
    assert torch.equal(eager.rec, graph.rec) and bool(eager.done.any())


@pytest.mark.parametrize("n,method,komi", ((5, "real", 0), (9, "heuristic", 2.5), (6, "real", 0.5), (19, "heuristic", 0)))
def test_goenv_random_games_against_env_oracle(n, method, komi):
    """whole random games through the drop-in GoEnv next to the oracle's restatement of the reference wrapper:
    states, rewards, done flags, info dict, valid moves, winning()"""
    from gymgo_b200.envs import GoEnv
    rng = np.random.RandomState(n)
    env, ref = GoEnv(n, komi=komi, reward_method=method), og.EnvOracle(n, komi=komi, reward_method=method)
    plies = 0
    for game in range(2 if n < 19 else 1):
        assert np.array_equal(env.reset(), ref.reset())
        done = False
        while not done and plies < (400 if n < 19 else 260):
            vm = env.valid_moves()
            assert np.array_equal(vm, ref.valid_moves())
            a = int(rng.choice(np.flatnonzero(vm)))
            s1, r1, d1, i1 = env.step(a)
            s2, r2, d2, i2 = ref.step(a)
            assert np.array_equal(s1, s2) and float(r1) == float(r2) and d1 == d2
            assert i1["turn"] == i2["turn"] and bool(i1["prev_player_passed"]) == bool(i2["prev_player_passed"])
            assert np.array_equal(i1["invalid_moves"], i2["invalid_moves"])
            done = bool(d1)
            plies += 1
        assert float(env.winning()) == float(ref.winning())
