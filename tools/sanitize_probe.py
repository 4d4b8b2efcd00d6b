"""Small workload touching every kernel, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymgo_b200.engine import GoEngine  # noqa: E402

for n, boards in ((9, 83), (19, 21), (7, 50), (5, 130)):
    e = GoEngine(n, "cuda:0")
    rec = e.new_records(boards)
    ring = e.empty((3, boards, 6, n, n), dtype=torch.float32)
    ring8 = e.empty((3, boards, 6, n, n), dtype=torch.uint8)
    acts = torch.empty((12, boards), dtype=torch.int32, device="cuda")
    done = torch.empty((12, boards), dtype=torch.uint8, device="cuda")
    rew = torch.empty((12, boards), dtype=torch.float32, device="cuda")
    e.rollout(rec, 1, 0, 0, 12, plies_per_launch=5, actions_log=acts, obs_ring=ring, done_log=done, reward_log=rew,
              reward_mode=2, komi=0.5)
    e.rollout(rec, 1, 0, 12, 7, plies_per_launch=7, obs_ring=ring8)
    # both persistent kernels, static and dynamically scheduled ((tile, 2-ply) tickets), every observation dtype
    ring16 = e.empty((3, boards, 6, n, n), dtype=torch.bfloat16)
    for kernel in (0, 1):
        for dyn in (False, True):
            for r in (ring, ring8, ring16):
                e.rollout(rec, 1, 0, 19, 9, plies_per_launch=9, actions_log=acts, obs_ring=r, done_log=done, reward_log=rew,
                          reward_mode=1, kernel=kernel, dynamic=dyn, block_plies=2)
    e.step(rec, e.sample_legal(rec, 2, 0, 5), out=rec, auto_reset=True, obs_dtype=torch.bfloat16)
    for kernel in ("lanes", "thread"):                      # both single-ply kernels, every output
        for dt in (torch.float32, torch.uint8, torch.float16):
            e.step(rec, e.sample_legal(rec, 2, 0, 6), out=rec, auto_reset="skip", obs_dtype=dt, want_done=True, want_areas=True,
                   reward_mode=2, komi=0.5, kernel=kernel)
    a = e.sample_legal(rec, 2, 0, 0)
    res = e.step(rec, a, obs_dtype=torch.float32, want_done=True, want_areas=True, reward_mode=1)
    res = e.step(res["rec"], e.sample_legal(res["rec"], 2, 0, 1), out=res["rec"], obs_dtype=torch.uint8)
    kids = e.children(rec, canonical=True, obs_dtype=torch.uint8, want_rec=True)
    e.areas(rec)
    e.valid_moves(rec, ended_quirk=True)
    e.canonical(rec)
    e.unpack(rec, dtype=torch.float64)
    e.pack(e.unpack(rec, dtype=torch.uint8))
    e.reset(rec, e.flags(rec) & 4 != 0)
    torch.cuda.synchronize()
print("sanitize probe done")
