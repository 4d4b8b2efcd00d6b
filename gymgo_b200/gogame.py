"""Drop-in for the reference's low-level function API `gym_go.gogame` (gym_go/gogame.py:22-468), backed by
the sm_100a kernels.  Same names, argument meaning, return types and error behaviour:

  * numpy in -> numpy float64 out (what the reference returns); torch CUDA tensors in -> torch tensors
    out, left on the device (no host round trip).
  * the rules (next_state, children, valid/invalid moves, areas, canonical form) run on the GPU through
    the C ABI; the 1-bit facts the reference reads off whole planes (turn, pass, done) are read the same
    way here.
  * an illegal move raises AssertionError like gogame.py:59 / :117.
  * `batch_next_states` equals per-board `next_state` - the reference's vectorised variant mis-aligns
    boards when a pass precedes a move in the batch (gym_go/state_utils.py:187-193);
    that bug is not reproduced.

There is no CPU implementation behind these functions: without a CUDA device they raise."""
import numpy as np
import torch

from . import govars
from .engine import engine as _engine

_STATUS_TEXT = {1: "Invalid move", 2: "Action out of range", 3: "Game over"}


# ------------------------------------------------------------------------------- helpers
def _is_torch(x):
    return isinstance(x, torch.Tensor)


def _size_of(state):
    return int(state.shape[-1])


def _to_device(eng, states):
    """[B,6,N,N] numpy/torch -> contiguous CUDA tensor in a dtype gg_pack reads"""
    if _is_torch(states):
        t = states
        if t.dtype not in (torch.uint8, torch.float32, torch.float64):
            t = t.to(torch.float32)
        return t.to(eng.device).contiguous()
    arr = np.ascontiguousarray(states)
    if arr.dtype not in (np.uint8, np.float32, np.float64):
        arr = arr.astype(np.float64)
    return torch.from_numpy(arr).to(eng.device)


def _like(result, template, np_dtype=np.float64):
    """device tensor -> same kind as the caller's input"""
    if _is_torch(template):
        return result
    return result.cpu().numpy().astype(np_dtype, copy=False)


def _out_dtype(template):
    if _is_torch(template) and template.dtype in (torch.uint8, torch.float32, torch.float64):
        return template.dtype
    return torch.float64 if not _is_torch(template) else torch.float32


def _raise_on_status(status, actions):
    bad = torch.nonzero(status).flatten()
    if bad.numel():
        i = int(bad[0])
        code = int(status[i])
        raise AssertionError((_STATUS_TEXT.get(code, "refused"), int(actions[i]), "board %d" % i))


# ------------------------------------------------------------------------------- state type
def init_state(size):
    """gogame.py:22-25"""
    return np.zeros((govars.NUM_CHNLS, size, size))


def batch_init_state(batch_size, board_size):
    """gogame.py:28-31"""
    return np.zeros((batch_size, govars.NUM_CHNLS, board_size, board_size))


def action_size(state=None, board_size=None):
    """gogame.py:189-197"""
    if state is not None:
        m, n = state.shape[1:]
    elif board_size is not None:
        m, n = board_size, board_size
    else:
        raise RuntimeError('No argument passed')
    return m * n + 1


# ------------------------------------------------------------------------------- the step
def batch_next_states(batch_states, batch_action1d, canonical=False):
    """gogame.py:90-150 (per-board semantics, see module docstring)."""
    eng = _engine(_size_of(batch_states))
    dev = _to_device(eng, batch_states)
    acts = torch.as_tensor(np.asarray(batch_action1d) if not _is_torch(batch_action1d) else batch_action1d)
    acts = acts.reshape(-1).to(torch.int64)
    if dev.shape[0] == 0:
        return _like(dev.to(_out_dtype(batch_states)), batch_states)
    rec = eng.pack(dev)
    res = eng.step(rec, acts.to(torch.int32), canonical=canonical)
    _raise_on_status(res["status"].cpu(), acts)
    return _like(eng.unpack(res["rec"], dtype=_out_dtype(batch_states)), batch_states)


def next_state(state, action1d, canonical=False):
    """gogame.py:34-87.  Pure: `state` is not modified."""
    return batch_next_states(state[None], [int(action1d)], canonical)[0]


def children(state, canonical=False, padded=True):
    """gogame.py:175-186: next_state for every valid action; zeros at invalid ones when padded."""
    eng = _engine(_size_of(state))
    rec = eng.pack(_to_device(eng, state[None]))
    res = eng.children(rec, canonical=canonical, obs_dtype=_out_dtype(state) if _out_dtype(state) != torch.float64
                       else torch.float32, want_rec=False)
    if int(res["status"][0]):
        raise AssertionError("Invalid move in children(): finished game with stones on the board")
    kids = res["obs"][0]
    if not padded:
        kids = kids[res["valid"][0].bool()]
    if _out_dtype(state) == torch.float64:
        kids = kids.to(torch.float64)
    return _like(kids, state)


# ------------------------------------------------------------------------------- move masks
def batch_valid_moves(batch_state):
    """gogame.py:171-172 (no ended-game special case in the batch variant)"""
    eng = _engine(_size_of(batch_state))
    rec = eng.pack(_to_device(eng, batch_state))
    return _like(eng.valid_moves(rec, ended_quirk=False, dtype=_out_dtype(batch_state)), batch_state)


def batch_invalid_moves(batch_state):
    """gogame.py:164-168"""
    return 1 - batch_valid_moves(batch_state)


def valid_moves(state):
    """gogame.py:160-161 - all ones once the game has ended (gogame.py:155-156)"""
    eng = _engine(_size_of(state))
    rec = eng.pack(_to_device(eng, state[None]))
    return _like(eng.valid_moves(rec, ended_quirk=True, dtype=_out_dtype(state))[0], state)


def invalid_moves(state):
    """gogame.py:153-157"""
    return 1 - valid_moves(state)


# ------------------------------------------------------------------------------- whole-plane facts
def turn(state):
    """gogame.py:241-246"""
    return int(state[govars.TURN_CHNL].max())


def batch_turn(batch_state):
    """gogame.py:249-250"""
    if _is_torch(batch_state):
        return batch_state[:, govars.TURN_CHNL].amax(dim=(1, 2)).to(torch.int64)
    return np.max(batch_state[:, govars.TURN_CHNL], axis=(1, 2)).astype(int)


def prev_player_passed(state):
    """gogame.py:200-201"""
    return (state[govars.PASS_CHNL] == 1).max() == 1


def batch_prev_player_passed(batch_state):
    """gogame.py:204-205"""
    if _is_torch(batch_state):
        return batch_state[:, govars.PASS_CHNL].amax(dim=(1, 2)) == 1
    return np.max(batch_state[:, govars.PASS_CHNL], axis=(1, 2)) == 1


def game_ended(state):
    """gogame.py:208-214 - int 0/1"""
    m, n = state.shape[1:]
    return int(int((state[govars.DONE_CHNL] == 1).sum()) == m * n)


def batch_game_ended(batch_state):
    """gogame.py:217-222"""
    if _is_torch(batch_state):
        return batch_state[:, govars.DONE_CHNL].amax(dim=(1, 2))
    return np.max(batch_state[:, govars.DONE_CHNL], axis=(1, 2))


# ------------------------------------------------------------------------------- scoring
def batch_areas(batch_state):
    """gogame.py:303-310 -> (black_areas [B], white_areas [B])"""
    eng = _engine(_size_of(batch_state))
    ar = eng.areas(eng.pack(_to_device(eng, batch_state)))
    if _is_torch(batch_state):
        return ar[:, 0], ar[:, 1]
    ar = ar.cpu().numpy().astype(np.float64)
    return ar[:, 0], ar[:, 1]


def areas(state):
    """gogame.py:275-300 -> (black_area, white_area)"""
    b, w = batch_areas(state[None])
    return b[0], w[0]


def winning(state, komi=0):
    """gogame.py:225-230"""
    b, w = areas(state)
    return torch.sign(b - w - komi) if _is_torch(state) else np.sign(b - w - komi)


def batch_winning(state, komi=0):
    """gogame.py:233-238"""
    b, w = batch_areas(state)
    return torch.sign(b - w - komi) if _is_torch(state) else np.sign(b - w - komi)


# ------------------------------------------------------------------------------- canonical form
def batch_canonical_form(batch_state):
    """gogame.py:324-337"""
    eng = _engine(_size_of(batch_state))
    if batch_state.shape[0] == 0:
        return batch_state.clone() if _is_torch(batch_state) else np.copy(batch_state)
    rec = eng.canonical(eng.pack(_to_device(eng, batch_state)))
    return _like(eng.unpack(rec, dtype=_out_dtype(batch_state)), batch_state)


def canonical_form(state):
    """gogame.py:313-321"""
    return batch_canonical_form(state[None])[0]


# ------------------------------------------------------------------------------- host-side helpers
# (SURVEY.md section 2 marks these as off the hot path: plain array manipulation, kept for API coverage)
def liberties(state):
    """gogame.py:253-264: per-COLOUR union of liberties (debug helper)."""
    st = np.asarray(state.cpu() if _is_torch(state) else state)
    occ = (st[govars.BLACK] + st[govars.WHITE]) > 0
    out = []
    for colour in (govars.BLACK, govars.WHITE):
        s = st[colour] > 0
        grow = np.zeros_like(s)
        grow[1:] |= s[:-1]
        grow[:-1] |= s[1:]
        grow[:, 1:] |= s[:, :-1]
        grow[:, :-1] |= s[:, 1:]
        out.append(grow & ~occ)
    return out[0], out[1]


def num_liberties(state):
    """gogame.py:267-272"""
    b, w = liberties(state)
    return np.count_nonzero(b), np.count_nonzero(w)


def _flip(image, axis):
    return torch.flip(image, dims=(axis,)) if _is_torch(image) else np.flip(image, axis)


def _rot90(image):
    return torch.rot90(image, 1, dims=(-2, -1)) if _is_torch(image) else np.rot90(image, 1, axes=(-2, -1))


def symmetry(image, orientation):
    """Element `orientation` (0..7) of all_symmetries: bit 0 mirrors the columns, bit 1 mirrors the rows, bit 2 turns
    the result a quarter counter-clockwise - applied in that order (gogame.py:347-355, :366-380).  The board axes are
    the last two (the reference names axes 1 and 2 of a [C,N,N] image; leading batch axes are allowed here).
    gg_symmetry / GoEngine.symmetry apply the same element to packed records."""
    orientation = int(orientation)
    if not 0 <= orientation < 8:
        raise ValueError("orientation must be in 0..7")
    if orientation & 1:
        image = _flip(image, -1)
    if orientation & 2:
        image = _flip(image, -2)
    if orientation & 4:
        image = _rot90(image)
    return image


def random_symmetry(image):
    """gogame.py:340-355: one uniformly drawn element of all_symmetries (one np.random.randint(0, 8) draw, so a
    seeded numpy stream picks the same orientation as the reference)."""
    return symmetry(image, np.random.randint(0, 8))


def all_symmetries(image):
    """gogame.py:358-382: the 8 dihedral transforms, in the reference's order (numpy or torch)."""
    return [symmetry(image, i) for i in range(8)]


def random_weighted_action(move_weights):
    """gogame.py:385-392: sample an action index with probability proportional to its weight."""
    w = np.asarray(move_weights, dtype=np.float64)
    return np.random.choice(np.arange(len(w)), p=w / w.sum())


def random_action(state):
    """gogame.py:395-404: uniform over valid moves incl. pass."""
    return random_weighted_action(np.asarray(valid_moves(state).cpu() if _is_torch(state) else valid_moves(state)))


_STONES = {0: "\u25cb", 1: "\u25cf"}                       # glyph of a black / white stone
_EDGE_LINK, _INNER_LINK = "\u2550", "\u2500"               # horizontal link on the first/last row, elsewhere
_CORNERS = ("\u2554\u2557", "\u255f\u2562", "\u255a\u255d")   # (left, right) end glyphs: first row, inner rows, last row
_JOINTS = ("\u2564", "\u253c", "\u2567")                        # inner-column glyph: first row, inner rows, last row


def str(state):  # noqa: A001 - the reference exports this name (gogame.py:407-468)
    """Text rendering, character for character the reference's: a tab-indented column header, one line per row
    (row index, tab, box-drawing grid with stones), then the turn / game-state line and the area line."""
    st = np.asarray(state.cpu() if _is_torch(state) else state)
    n = st.shape[1]
    last = n - 1
    lines = ["\t" + "".join("{:<2d}".format(c) for c in range(n))]
    for r in range(n):
        band = 0 if r == 0 else (2 if r == last else 1)
        link = _INNER_LINK if band == 1 else _EDGE_LINK
        cells = []
        for c in range(n):
            colour = 0 if st[govars.BLACK, r, c] == 1 else (1 if st[govars.WHITE, r, c] == 1 else None)
            if colour is not None:
                glyph = _STONES[colour]
            elif c == 0 or c == last:
                glyph = _CORNERS[band][0 if c == 0 else 1]
            else:
                glyph = _JOINTS[band]
            cells.append(glyph if c == last else glyph + link)
        lines.append("{}\t{}".format(r, "".join(cells)))
    black_area, white_area = areas(st)
    phase = "END" if game_ended(st) else ("PASSED" if prev_player_passed(st) else "ONGOING")
    lines.append("\tTurn: {}, Game State (ONGOING|PASSED|END): {}".format("WHITE" if turn(st) else "BLACK", phase))
    lines.append("\tBlack Area: {}, White Area: {}".format(int(black_area), int(white_area)))
    return "\n".join(lines) + "\n"
