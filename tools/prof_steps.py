"""Developer probe (not part of the product): per-ply kernel time as the game phase evolves, C-side vs Python
launch loops, host launch overhead."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from gymgo_b200.engine import GoEngine  # noqa: E402


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3  # us


out = {}
for size, boards in ((9, 65536), (19, 16384)):
    eng = GoEngine(size, "cuda:0")
    rec = eng.new_records(boards)
    ring = eng.empty((3, boards, 6, size, size), dtype=torch.float32)
    done = eng.empty((boards,))
    rew = eng.empty((boards,), dtype=torch.float32)
    chunks = []
    t = 0
    for _ in range(16):
        us = timed(lambda: eng.rollout(rec, 0, 0, t, 50, obs_ring=ring))
        t += 50
        chunks.append(round(us / 50, 2))
    ppl = {}
    for k in (1, 2, 4, 8, 16, 50):
        ppl[k] = round(timed(lambda: eng.rollout(rec, 0, 0, t, 50, plies_per_launch=k, obs_ring=ring)) / 50, 2)
        t += 50
    os.environ["GG_ROLLOUT_VARIANT"] = "1"
    tpb = {}
    for k in (1, 16, 50):
        tpb[k] = round(timed(lambda: eng.rollout(rec, 0, 0, t, 50, plies_per_launch=k, obs_ring=ring)) / 50, 2)
        t += 50
    tpb["no_obs_16"] = round(timed(lambda: eng.rollout(rec, 0, 0, t, 48, plies_per_launch=16, obs_ring=None)) / 48, 2)
    t += 48
    sliced = {}
    if size > 9:
        os.environ["GG_ROLLOUT_VARIANT"] = "2"
        for k in (2, 3, 4):
            os.environ["GG_ROLLOUT_K"] = str(k)
            a = round(timed(lambda: eng.rollout(rec, 0, 0, t, 48, plies_per_launch=16, obs_ring=ring)) / 48, 2)
            b = round(timed(lambda: eng.rollout(rec, 0, 0, t, 48, plies_per_launch=16, obs_ring=None)) / 48, 2)
            os.environ["GG_ROLLOUT_VARIANT"] = "0"
            c = round(timed(lambda: eng.rollout(rec, 0, 0, t, 48, plies_per_launch=16, obs_ring=ring)) / 48, 2)
            d = round(timed(lambda: eng.rollout(rec, 0, 0, t, 48, plies_per_launch=16, obs_ring=None)) / 48, 2)
            os.environ["GG_ROLLOUT_VARIANT"] = "2"
            sliced[k] = dict(obs=a, no_obs=b, default_same_phase_obs=c, default_same_phase_no_obs=d)
            t += 48
    os.environ["GG_ROLLOUT_VARIANT"] = "0"
    noobs = timed(lambda: eng.rollout(rec, 0, 0, t, 50, obs_ring=None)) / 50
    t += 50
    u8ring = eng.empty((3, boards, 6, size, size), dtype=torch.uint8)
    u8 = timed(lambda: eng.rollout(rec, 0, 0, t, 50, obs_ring=u8ring)) / 50
    t += 50

    def pyloop():
        for k in range(100):
            eng.rollout_step(rec, 0, 0, t + k, obs=ring[k % 3], done=done, reward=rew, reward_mode=1)
    py = timed(pyloop) / 100
    t += 100
    small = eng.new_records(64)
    t0 = time.time()
    for k in range(2000):
        eng.rollout_step(small, 0, 0, k)
    torch.cuda.synchronize()
    host = (time.time() - t0) / 2000 * 1e6
    out["%dx%d" % (size, size)] = dict(us_per_ply_by_50=chunks, us_per_ply_by_plies_per_launch=ppl, thread_per_board_variant=tpb, sliced_variant=sliced, no_obs_us=round(noobs, 2), u8_us=round(u8, 2),
                                       python_loop_us=round(py, 2), host_call_us_tiny_batch=round(host, 2))
print(json.dumps(out))
