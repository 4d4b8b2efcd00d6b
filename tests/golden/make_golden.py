#!/usr/bin/env python
"""Generate the committed golden fixtures FROM THE REAL, UNMODIFIED REFERENCE.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Outputs (small, bit-packed .npz files next to this script):
  kat.npz            scripted known-answer sequences (the move lists are
                     the ones the reference's unit tests play: gym_go/tests/test_basics.py,
                     test_invalid_moves.py, test_valid_moves.py) replayed through the reference
                     GoEnv.step, with every intermediate state, reward, done flag and the
                     "next move must raise" expectations.
  traj_n{N}.npz      seeded uniform-random-legal trajectories through reference gogame.next_state
                     for N in {3,5,7,9,13,19}: states, actions, areas, valid_moves.
  soup_n{N}.npz      hand-built ("soup") positions that play never reaches (zero-liberty groups,
                     stale INVD planes, moves on finished games) stepped through the reference.
  children_n{N}.npz  reference gogame.children(state, canonical, padded=True) on mid-game parents.
  env_n7.npz         GoEnv-level behaviour: rewards ('real', 'heuristic', komi), info dict values.
  misc.npz           the helpers around the step: gogame.all_symmetries (element by
                     element, N in {5,9,19}), gogame.random_symmetry under seeded numpy streams,
                     gogame.str renderings, gogame.liberties, and state_utils.update_pieces /
                     adj_data on capturing and non-capturing placements (new state + killed groups).

Only OUTPUTS of the reference are stored, never its code.  All states are 0/1 so
they are stored with np.packbits; `tests/golden_io.py` is the reader.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

warnings.simplefilter("ignore")
gym, gogame, govars = _refshim.load_reference()
from gym_go import state_utils  # noqa: E402  (the reference's)

P = None  # pass marker in the move lists


def pack_states(arr):
    arr = np.asarray(arr)
    assert set(np.unique(arr)).issubset({0.0, 1.0}), "non-binary state"
    return np.packbits(arr.astype(np.uint8).reshape(-1)), np.array(arr.shape, dtype=np.int64)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote %-22s %7.1f KB" % (name, os.path.getsize(path) / 1024.0))


# --------------------------------------------------------------------------- KATs
# (name, board size, reward_method, komi, moves, move_that_must_raise or 'none')
KATS = [
    ("first_move", 7, "real", 0, [(0, 0)], "none"),
    ("turns", 7, "real", 0, [(i, 0) for i in range(7)], "none"),
    ("pass_then_move", 7, "real", 0, [P, (0, 0)], "none"),
    ("double_pass", 7, "real", 0, [P, P], (0, 0)),
    ("move_pass_pass", 7, "real", 0, [(0, 0), P, P], P),
    ("pass_move_pass", 7, "real", 0, [P, (0, 0), P], "none"),
    ("diamond_liberties", 7, "real", 0, [(2, 1), P, (1, 2), P, (2, 3), P, (3, 2), P], "none"),
    ("komi_a", 7, "real", 2.5, [P, P], "none"),
    ("komi_b", 7, "real", 2.5, [0, 2, 1, P, P], "none"),
    ("komi_c", 7, "real", 2.5, [0, P, 1, P, 2, P, P], "none"),
    ("real_a", 7, "real", 0, [(0, 0), P, P], "none"),
    ("real_b", 7, "real", 0, [P, (0, 0), P, P], "none"),
    ("heur_a", 7, "heuristic", 0, [(0, 0), (0, 1), P, (1, 0), P, P], "none"),
    ("heur_b", 7, "heuristic", 0, [(0, 0), P, P], "none"),
    ("occupied", 7, "real", 0, [(3, 3)], (3, 3)),
    ("ko", 7, "real", 0, [(0, 1), (0, 2), (1, 0), (1, 3), (2, 1), (2, 2), (1, 2), (1, 1)], (1, 2)),
    ("ko_expires", 7, "real", 0,
     [(0, 1), (0, 2), (1, 0), (1, 3), (2, 1), (2, 2), (1, 2), (1, 1), (6, 6), P], "none"),
    ("ko_wall", 7, "real", 0, [(1, 0), (0, 0), P, (1, 1), P, (0, 2), (0, 1)], (0, 0)),
    ("ko_wall_expires", 7, "real", 0, [(1, 0), (0, 0), P, (1, 1), P, (0, 2), (0, 1), (6, 6), P], "none"),
    ("suicide", 7, "real", 0, [(0, 1), (0, 2), (1, 0), (1, 4), (2, 1), (2, 2), (1, 2)], (1, 1)),
    ("suicide_3x3", 3, "real", 0, [6, 7, 8, 5, 4, 8, 0, 1], 3),
    ("after_capture_3x3", 3, "real", 0, [0, 8, 6, 4, 1, 2, 3, 7], 5),
    ("two_eyes", 7, "real", 0,
     [(1, 1), (0, 1), (1, 2), (0, 2), (1, 3), (0, 3), (1, 4), (0, 4), (1, 5), (0, 5), (2, 5), (1, 6),
      (3, 5), (2, 6), (3, 4), (3, 6), (3, 3), (4, 5), (2, 3), (4, 4), (3, 2), (4, 3), (3, 1), (4, 2),
      (2, 1), (4, 1), P, (3, 0), P, (2, 0), P, (1, 0), P], (2, 2)),
    ("capturing_move_valid", 7, "real", 0, [(0, 0), (0, 2), (0, 3), (1, 1), (1, 2), (1, 0), (0, 1)], "none"),
    ("simple_capture", 7, "real", 0, [(0, 1), (1, 1), (1, 0), P, (1, 2), P, (2, 1)], "none"),
    ("big_capture", 7, "real", 0,
     [(2, 2), (1, 2), (2, 3), (1, 3), (2, 4), (1, 4), (3, 4), (2, 5), (3, 3), (3, 5), (3, 2), (4, 4),
      P, (4, 3), P, (4, 2), P, (3, 1), P, (2, 1)], "none"),
    ("big_suicide", 7, "real", 0, [(4, 0), (6, 0), (4, 1), (5, 0), (5, 2), (5, 1), (6, 2)], (6, 1)),
    ("edge_capture", 7, "real", 0, [(0, 0), (0, 2), (0, 1), (1, 2), (1, 1), (2, 1), (1, 0), (2, 0)], "none"),
    ("group_kill_no_ko", 7, "real", 0,
     [(0, 5), (0, 4), (1, 5), (1, 4), (2, 5), (2, 4), (2, 6), (3, 5), P, (3, 6), P, (1, 6), (0, 6), (1, 6)],
     "none"),
    ("sizes_13", 13, "real", 0, [(0, 0), (12, 12), (6, 6), P, (0, 1)], "none"),
    ("sizes_19", 19, "real", 0, [(0, 0), (18, 18), (9, 9), P, (0, 1)], "none"),
]


def to_action1d(move, n):
    if move is None:
        return n * n
    if isinstance(move, tuple):
        return move[0] * n + move[1]
    return int(move)


def gen_kats():
    out = {}
    names = []
    for name, n, method, komi, moves, must_raise in KATS:
        env = gym.make("gym_go:go-v0", size=n, komi=komi, reward_method=method)
        env.reset()
        states, rewards, dones, turns, pps = [env.state()], [], [], [], []
        for mv in moves:
            st, rew, done, info = env.step(mv)
            states.append(st)
            rewards.append(float(rew))
            dones.append(int(done))
            turns.append(int(info["turn"]))
            pps.append(int(bool(info["prev_player_passed"])))
        raises = -1
        if must_raise != "none":
            try:
                env.step(must_raise)
                raise SystemExit("KAT %s: reference did not raise" % name)
            except SystemExit:
                raise
            except Exception:
                raises = to_action1d(must_raise, n)
        bits, shape = pack_states(np.stack(states))
        names.append(name)
        out[name + "__states"] = bits
        out[name + "__shape"] = shape
        out[name + "__actions"] = np.array([to_action1d(m, n) for m in moves], dtype=np.int64)
        out[name + "__rewards"] = np.array(rewards, dtype=np.float64)
        out[name + "__dones"] = np.array(dones, dtype=np.int64)
        out[name + "__turns"] = np.array(turns, dtype=np.int64)
        out[name + "__prev_pass"] = np.array(pps, dtype=np.int64)
        out[name + "__raises"] = np.array(raises, dtype=np.int64)
        out[name + "__komi"] = np.array(komi, dtype=np.float64)
        out[name + "__method"] = np.array(method)
    out["names"] = np.array(names)
    save("kat.npz", **out)


# ------------------------------------------------------------------ trajectories
def rollout(n, steps, rng):
    """uniform over valid actions incl. pass (reference go_env.py:78-81), reset on game end."""
    state = gogame.init_state(n)
    S, A, AR, VM, RESET = [], [], [], [], []
    for t in range(steps):
        vm = gogame.valid_moves(state)
        a = int(rng.choice(np.argwhere(vm).flatten()))
        nxt = gogame.next_state(state, a)
        S.append(state)
        A.append(a)
        AR.append(gogame.areas(state))
        VM.append(vm)
        if gogame.game_ended(nxt):
            # record the terminal transition, then the ended-state quirk, then restart
            S.append(nxt)
            A.append(-1)  # marker: no transition out of this state
            AR.append(gogame.areas(nxt))
            VM.append(gogame.valid_moves(nxt))
            state = gogame.init_state(n)
        else:
            state = nxt
    S.append(state)
    A.append(-1)
    AR.append(gogame.areas(state))
    VM.append(gogame.valid_moves(state))
    return np.stack(S), np.array(A, dtype=np.int64), np.array(AR, dtype=np.float64), np.stack(VM)


def gen_traj():
    plan = {3: 600, 5: 1200, 7: 1500, 9: 2500, 13: 1500, 19: 1600}
    for n, steps in plan.items():
        rng = np.random.RandomState(1000 + n)
        S, A, AR, VM = rollout(n, steps, rng)
        # next state of record i is record i+1 whenever A[i] >= 0
        bits, shape = pack_states(S)
        vbits, vshape = pack_states(VM)
        save("traj_n%d.npz" % n, states=bits, shape=shape, actions=A, areas=AR,
             valid=vbits, valid_shape=vshape)


# ------------------------------------------------------------------------- soup
def gen_soup():
    """Positions play never reaches: random stone soups (zero-liberty groups allowed), with the
    INVD plane the reference itself computes for them, plus stale/empty INVD planes and moves on
    finished games."""
    for n, count in ((2, 120), (4, 300), (5, 400), (9, 500), (19, 120)):
        rng = np.random.RandomState(2000 + n)
        S0, A, S1, AR = [], [], [], []
        tries = 0
        while len(S0) < count:
            tries += 1
            density = rng.uniform(0.2, 0.95)
            r = rng.uniform(size=(n, n))
            cut = rng.uniform(0.3, 0.7)
            st = gogame.init_state(n)
            st[govars.BLACK] = (r < density * cut)
            st[govars.WHITE] = (r >= density * cut) & (r < density)
            turn = int(rng.randint(2))
            st[govars.TURN_CHNL] = turn
            mode = rng.randint(4)
            if mode == 0:
                pass  # stale: INVD plane all zero although stones exist
            else:
                # reference's own mask for "player (1-turn) just moved"
                st[govars.INVD_CHNL] = state_utils.compute_invalid_moves(st, 1 - turn, None)
            if mode == 2:
                st[govars.PASS_CHNL] = 1
            if mode == 3:
                st[govars.DONE_CHNL] = 1  # finished game: next_state still works (A.3)
                st[govars.PASS_CHNL] = rng.randint(2)
            legal = np.argwhere(np.append(1 - st[govars.INVD_CHNL].flatten(), 1)).flatten()
            if mode == 0:
                # stale plane: only play on really-empty points (occupied+stale is out of domain)
                occ = (st[0] + st[1]).flatten()
                legal = np.array([a for a in legal if a == n * n or occ[a] == 0])
            a = int(rng.choice(legal))
            S0.append(st)
            A.append(a)
            S1.append(gogame.next_state(st, a))
            AR.append(gogame.areas(st))
        b0, shape = pack_states(np.stack(S0))
        b1, _ = pack_states(np.stack(S1))
        save("soup_n%d.npz" % n, states=b0, next_states=b1, shape=shape,
             actions=np.array(A, dtype=np.int64), areas=np.array(AR, dtype=np.float64))


# --------------------------------------------------------------------- children
def gen_children():
    for n, parents, plies in ((3, 10, 4), (5, 12, 12), (7, 12, 20), (9, 12, 40), (19, 3, 150)):
        rng = np.random.RandomState(3000 + n)
        PS, CH0, CH1 = [], [], []
        for _ in range(parents):
            state = gogame.init_state(n)
            k = int(rng.randint(max(1, plies // 2), plies + 1))
            for _t in range(k):
                vm = gogame.valid_moves(state)
                a = int(rng.choice(np.argwhere(vm).flatten()))
                nxt = gogame.next_state(state, a)
                if gogame.game_ended(nxt):
                    break
                state = nxt
            PS.append(state)
            CH0.append(gogame.children(state, canonical=False, padded=True))
            CH1.append(gogame.children(state, canonical=True, padded=True))
        pb, pshape = pack_states(np.stack(PS))
        c0, cshape = pack_states(np.stack(CH0))
        c1, _ = pack_states(np.stack(CH1))
        save("children_n%d.npz" % n, parents=pb, parents_shape=pshape, children=c0,
             children_canonical=c1, children_shape=cshape)


# ------------------------------------------------------------------- env-level
def gen_env():
    """Full GoEnv.step games on 7x7 (BASELINE.json configs[0]) incl. rewards for both methods."""
    out = {}
    for gi, (method, komi) in enumerate((("real", 0), ("heuristic", 0), ("real", 2.5), ("heuristic", 5.5))):
        rng = np.random.RandomState(4000 + gi)
        for game in range(3):
            env = gym.make("gym_go:go-v0", size=7, komi=komi, reward_method=method)
            env.reset()
            acts, rews, dones, states = [], [], [], [env.state()]
            done = False
            while not done and len(acts) < 400:
                vm = env.valid_moves()
                a = int(rng.choice(np.argwhere(vm).flatten()))
                st, r, done, info = env.step(a)
                acts.append(a)
                rews.append(float(r))
                dones.append(int(done))
                states.append(st)
            key = "g%d_%d" % (gi, game)
            bits, shape = pack_states(np.stack(states))
            out[key + "__states"] = bits
            out[key + "__shape"] = shape
            out[key + "__actions"] = np.array(acts, dtype=np.int64)
            out[key + "__rewards"] = np.array(rews, dtype=np.float64)
            out[key + "__dones"] = np.array(dones, dtype=np.int64)
            out[key + "__komi"] = np.array(komi, dtype=np.float64)
            out[key + "__method"] = np.array(method)
            out[key + "__winning"] = np.array(float(env.winning()), dtype=np.float64)
    out["keys"] = np.array(sorted({k.split("__")[0] for k in out}))
    save("env_n7.npz", **out)


# ------------------------------------------------------------------------ misc
def _midgame(n, plies, rng):
    state = gogame.init_state(n)
    for _ in range(plies):
        a = int(rng.choice(np.argwhere(gogame.valid_moves(state)).flatten()))
        nxt = gogame.next_state(state, a)
        if gogame.game_ended(nxt):
            break
        state = nxt
    return state


def gen_misc():
    out = {}
    # (1) all_symmetries element by element + random_symmetry under seeded streams
    for n, plies in ((5, 14), (9, 50), (19, 220)):
        rng = np.random.RandomState(5000 + n)
        img = _midgame(n, plies, rng)
        syms = np.stack([np.ascontiguousarray(x) for x in gogame.all_symmetries(img)])
        out["sym_n%d__image" % n], out["sym_n%d__image_shape" % n] = pack_states(img)
        out["sym_n%d__all" % n], out["sym_n%d__all_shape" % n] = pack_states(syms)
        picks = []
        for seed in range(24):
            np.random.seed(seed)
            picks.append(np.ascontiguousarray(gogame.random_symmetry(img)))
        out["sym_n%d__random" % n], out["sym_n%d__random_shape" % n] = pack_states(np.stack(picks))
    # (2) str renderings, liberties
    texts, tstates, libs = [], [], []
    for n, plies in ((3, 5), (5, 10), (7, 30), (7, 60), (9, 70)):
        rng = np.random.RandomState(5100 + n + plies)
        st = _midgame(n, plies, rng)
        for extra in ([], [n * n], [n * n, n * n]):
            s2 = st
            for a in extra:
                s2 = gogame.next_state(s2, a)
            texts.append(gogame.str(s2))
            key = "text%d" % (len(texts) - 1)
            out[key + "__state"], out[key + "__shape"] = pack_states(s2)
            lb, lw = gogame.liberties(s2)
            out[key + "__liberties"], out[key + "__liberties_shape"] = pack_states(np.stack([lb, lw]).astype(np.float64))
    out["texts"] = np.array(texts)
    # (3) update_pieces / adj_data: place a stone like next_state does (gogame.py:61-69) and let the reference
    # remove the captured groups; killed groups are stored as a label plane (k-th returned group -> value k+1)
    cases = 0
    for n, count in ((5, 60), (7, 60), (9, 60), (19, 12)):
        rng = np.random.RandomState(5200 + n)
        state = gogame.init_state(n)
        got = 0
        while got < count:
            vm = gogame.valid_moves(state)
            a = int(rng.choice(np.argwhere(vm).flatten()))
            if a < n * n:
                player = gogame.turn(state)
                r, c = a // n, a % n
                placed = np.copy(state)
                placed[player, r, c] = 1
                adj, surrounded = state_utils.adj_data(placed, np.array([r, c]), player)
                after = np.copy(placed)
                killed = state_utils.update_pieces(after, adj, player)
                # keep every capture and a thinning of the quiet moves
                if killed or rng.randint(4) == 0:
                    label = np.zeros((n, n), dtype=np.int64)
                    for k, grp in enumerate(killed):
                        label[grp[:, 0], grp[:, 1]] = k + 1
                    key = "up%d" % cases
                    out[key + "__before"], out[key + "__shape"] = pack_states(placed)
                    out[key + "__after"], _ = pack_states(after)
                    out[key + "__killed"] = label
                    out[key + "__meta"] = np.array([n, r, c, player, int(bool(surrounded)), len(killed)], dtype=np.int64)
                    out[key + "__adj"] = np.asarray(adj, dtype=np.int64)
                    cases += 1
                    got += 1
            state = gogame.next_state(state, a)
            if gogame.game_ended(state):
                state = gogame.init_state(n)
    out["update_cases"] = np.array(cases, dtype=np.int64)
    save("misc.npz", **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["misc"]:
        gen_misc()
        raise SystemExit(0)
    gen_misc()
    gen_kats()
    gen_traj()
    gen_soup()
    gen_children()
    gen_env()
