"""Reader for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py from the
real reference).  Pure numpy; usable on the GPU box (never touches /root/reference)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _unpack(bits, shape):
    shape = tuple(int(x) for x in shape)
    count = int(np.prod(shape))
    return np.unpackbits(bits)[:count].reshape(shape).astype(np.float64)


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def kat_cases():
    z = load("kat.npz")
    out = []
    for name in z["names"]:
        name = str(name)
        out.append(dict(
            name=name,
            states=_unpack(z[name + "__states"], z[name + "__shape"]),
            actions=z[name + "__actions"], rewards=z[name + "__rewards"], dones=z[name + "__dones"],
            turns=z[name + "__turns"], prev_pass=z[name + "__prev_pass"],
            raises=int(z[name + "__raises"]), komi=float(z[name + "__komi"]), method=str(z[name + "__method"])))
    return out


def trajectory(n):
    """-> states [T,6,N,N], actions [T] (-1 = no transition out of this record), areas [T,2], valid [T,A]."""
    z = load("traj_n%d.npz" % n)
    return (_unpack(z["states"], z["shape"]), z["actions"], z["areas"], _unpack(z["valid"], z["valid_shape"]))


def transitions(n):
    """-> (s0 [M,6,N,N], a [M], s1 [M,6,N,N]) extracted from the trajectory fixture."""
    S, A, _, _ = trajectory(n)
    idx = np.flatnonzero(A >= 0)
    return S[idx], A[idx], S[idx + 1]


def soup(n):
    z = load("soup_n%d.npz" % n)
    return (_unpack(z["states"], z["shape"]), z["actions"], _unpack(z["next_states"], z["shape"]), z["areas"])


def children(n):
    z = load("children_n%d.npz" % n)
    return (_unpack(z["parents"], z["parents_shape"]), _unpack(z["children"], z["children_shape"]),
            _unpack(z["children_canonical"], z["children_shape"]))


def env_games():
    z = load("env_n7.npz")
    out = []
    for key in z["keys"]:
        key = str(key)
        out.append(dict(key=key, states=_unpack(z[key + "__states"], z[key + "__shape"]),
                        actions=z[key + "__actions"], rewards=z[key + "__rewards"], dones=z[key + "__dones"],
                        komi=float(z[key + "__komi"]), method=str(z[key + "__method"]),
                        winning=float(z[key + "__winning"])))
    return out


def symmetries(n):
    """-> (image [6,N,N], all [8,6,N,N] = reference all_symmetries(image), random [24,6,N,N] = reference
    random_symmetry(image) after np.random.seed(0..23))"""
    z = load("misc.npz")
    k = "sym_n%d__" % n
    return (_unpack(z[k + "image"], z[k + "image_shape"]), _unpack(z[k + "all"], z[k + "all_shape"]),
            _unpack(z[k + "random"], z[k + "random_shape"]))


def texts():
    """-> list of (state [6,N,N], reference gogame.str(state), liberties [2,N,N])"""
    z = load("misc.npz")
    out = []
    for i, text in enumerate(z["texts"]):
        k = "text%d__" % i
        out.append((_unpack(z[k + "state"], z[k + "shape"]), str(text),
                    _unpack(z[k + "liberties"], z[k + "liberties_shape"])))
    return out


def update_pieces_cases():
    """-> list of dict(before, after [6,N,N], killed [N,N] label plane, n, point, player, surrounded, groups, adj)"""
    z = load("misc.npz")
    out = []
    for i in range(int(z["update_cases"])):
        k = "up%d__" % i
        n, r, c, player, surrounded, groups = (int(x) for x in z[k + "meta"])
        out.append(dict(before=_unpack(z[k + "before"], z[k + "shape"]), after=_unpack(z[k + "after"], z[k + "shape"]),
                        killed=z[k + "killed"], n=n, point=(r, c), player=player, surrounded=bool(surrounded),
                        groups=groups, adj=z[k + "adj"]))
    return out


SYM_SIZES = (5, 9, 19)
TRAJ_SIZES = (3, 5, 7, 9, 13, 19)
SOUP_SIZES = (2, 4, 5, 9, 19)
CHILDREN_SIZES = (3, 5, 7, 9, 19)
