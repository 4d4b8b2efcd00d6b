"""Drop-in for the parts of the reference's `gym_go/state_utils.py` that make sense outside `next_state`.

In the reference these L1 primitives (group labelling, liberties, capture, legality) are separate numpy/scipy
passes (`state_utils.py:24-250`); here they are fused inside the step kernel (gg_step).  What is exposed:

  * compute_invalid_moves / batch_compute_invalid_moves (state_utils.py:24-156) - evaluated by the same device
    code path: the legality mask "after `player` moved" is exactly what a pass by `player` recomputes
    (gogame.py:48-53,78), so we set the turn plane to `player`, pass on the GPU and read the INVD plane.
  * adj_data / batch_adj_data (state_utils.py:214-232), set_turn / batch_set_turn (:235-250): tiny array helpers.
  * update_pieces / batch_update_pieces (:159-211) have no standalone equivalent - capture removal happens
    inside gg_step; calling them raises NotImplementedError with that pointer."""
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


def update_pieces(*args, **kwargs):
    raise NotImplementedError("capture removal is fused into the step kernel (gg_step / gogame.next_state)")


batch_update_pieces = update_pieces
