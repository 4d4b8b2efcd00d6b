"""Drop-in for the parts of the reference's `gym_go/state_utils.py` that make sense outside `next_state`.

In the reference these L1 primitives (group labelling, liberties, capture, legality) are separate numpy/scipy
passes (`state_utils.py:24-250`); here they are fused inside the step kernel (gg_step).  What is exposed:

  * compute_invalid_moves / batch_compute_invalid_moves (state_utils.py:24-156) - evaluated by the same device
    code path: the legality mask "after `player` moved" is exactly what a pass by `player` recomputes
    (gogame.py:48-53,78), so we set the turn plane to `player`, pass on the GPU and read the INVD plane.
  * adj_data / batch_adj_data (state_utils.py:214-232), set_turn / batch_set_turn (:235-250): tiny array helpers.
  * update_pieces / batch_update_pieces (:159-211): the capture routine of the step kernel as a stand-alone launch
    (gg_update_pieces); the state array is updated in place and the killed groups come back as coordinate arrays
    in the reference's order (groups by their first stone in raster order, stones in raster order)."""
import numpy as np

from . import govars
from .engine import engine as _engine

neighbor_deltas = np.array([[-1, 0], [1, 0], [0, -1], [0, 1]])


def batch_compute_invalid_moves(batch_state, batch_player, batch_ko_protect):
    batch_state = np.asarray(batch_state)
    b, n = len(batch_state), batch_state.shape[-1]
    if b == 0:
        return np.zeros(batch_state.shape[:1] + batch_state.shape[2:], dtype=bool)
    st = np.array(batch_state, dtype=np.float64, copy=True)
    st[:, govars.TURN_CHNL] = np.asarray(batch_player, dtype=np.float64).reshape(b, 1, 1)
    st[:, govars.PASS_CHNL] = 0
    st[:, govars.DONE_CHNL] = 0
    eng = _engine(n)
    import torch
    rec = eng.pack(torch.from_numpy(st).to(eng.device))
    res = eng.step(rec, np.full(b, n * n, dtype=np.int32), obs_dtype=torch.uint8)
    out = res["obs"][:, govars.INVD_CHNL].cpu().numpy().astype(bool)
    for i, ko in enumerate(batch_ko_protect):
        if ko is not None:
            out[i, ko[0], ko[1]] = True
    return out


def compute_invalid_moves(state, player, ko_protect=None):
    return batch_compute_invalid_moves(np.asarray(state)[None], [player], [ko_protect])[0]


def adj_data(state, action2d, player):
    n = state.shape[1]
    neighbors = neighbor_deltas + np.asarray(action2d)
    neighbors = neighbors[((neighbors >= 0) & (neighbors < n)).all(axis=1)]
    surrounded = (state[1 - player][neighbors[:, 0], neighbors[:, 1]] > 0).all()
    return neighbors, surrounded


def batch_adj_data(batch_state, batch_action2d, batch_player):
    pairs = [adj_data(s, a, p) for s, a, p in zip(batch_state, batch_action2d, batch_player)]
    return [p[0] for p in pairs], [p[1] for p in pairs]


def set_turn(state):
    state[govars.TURN_CHNL] = 1 - state[govars.TURN_CHNL]


def batch_set_turn(batch_state):
    batch_state[:, govars.TURN_CHNL] = 1 - batch_state[:, govars.TURN_CHNL]


def _split_groups(dead):
    """bool [N,N] plane of removed stones -> list of [k,2] coordinate arrays, one per 4-connected group, ordered like
    ndimage.label numbers them (by first stone in raster order); stones inside a group in raster order (np.argwhere)"""
    groups = []
    left = dead.copy()
    n = dead.shape[0]
    while left.any():
        r, c = np.argwhere(left)[0]
        member = np.zeros_like(left)
        frontier = [(int(r), int(c))]
        member[r, c] = True
        while frontier:
            y, x = frontier.pop()
            for dy, dx in neighbor_deltas:
                v, u = y + dy, x + dx
                if 0 <= v < n and 0 <= u < n and left[v, u] and not member[v, u]:
                    member[v, u] = True
                    frontier.append((v, u))
        groups.append(np.argwhere(member))
        left &= ~member
    return groups


def batch_update_pieces(batch_non_pass, batch_state, batch_adj_locs, batch_player):
    """state_utils.py:183-211: for the boards `batch_non_pass` of `batch_state` (in place), remove the groups of the
    opponent of `batch_player[i]` that touch `batch_adj_locs[i]` and have no liberty; -> list (per listed board) of
    lists of killed groups."""
    import torch
    idx = np.asarray(batch_non_pass, dtype=np.int64).reshape(-1)
    if len(idx) == 0:
        return []
    n = batch_state.shape[-1]
    eng = _engine(n)
    players = np.asarray(batch_player, dtype=np.int64).reshape(-1)
    if len(players) != len(idx):
        players = players[idx]                      # the reference passes the full-batch vector (state_utils.py:184)
    touch = np.zeros((len(idx), govars.NUM_CHNLS, n, n), dtype=np.uint8)
    for i, locs in enumerate(batch_adj_locs):
        locs = np.asarray(locs, dtype=np.int64).reshape(-1, 2)
        touch[i, govars.BLACK, locs[:, 0], locs[:, 1]] = 1
    sub = np.ascontiguousarray(np.asarray(batch_state)[idx] != 0).astype(np.uint8)
    rec = eng.pack(torch.from_numpy(sub).to(eng.device))
    killed = eng.update_pieces(rec, eng.pack(torch.from_numpy(touch).to(eng.device)), players)
    dead = eng.unpack(killed, dtype=torch.uint8)[:, govars.BLACK].cpu().numpy().astype(bool)
    out = []
    for i, b in enumerate(idx):
        batch_state[b, 1 - int(players[i])][dead[i]] = 0
        out.append(_split_groups(dead[i]))
    return out


def update_pieces(state, adj_locs, player):
    """state_utils.py:159-180 (state modified in place; returns the list of killed groups)"""
    return batch_update_pieces([0], state[None], [adj_locs], [player])[0]
