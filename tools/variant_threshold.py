"""Developer probe: lane-sliced vs thread-per-board rollout kernel as a function of the batch size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymgo_b200.engine import GoEngine  # noqa: E402


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


out = {}
for size in (9, 7, 5):
    eng = GoEngine(size, "cuda:0")
    for boards in (4096, 8192, 16384, 32768, 65536, 131072):
        ring = eng.empty((2, boards, 6, size, size), dtype=torch.float32)
        row = {}
        for variant in ("0", "1"):
            os.environ["GG_ROLLOUT_VARIANT"] = variant
            rec = eng.new_records(boards)
            eng.rollout(rec, 0, 0, 0, 96, plies_per_launch=32, obs_ring=ring)
            us = timed(lambda: eng.rollout(rec, 0, 0, 96, 96, plies_per_launch=32, obs_ring=ring)) / 96
            row["tpb" if variant == "1" else "lane"] = round(us, 2)
        out["%dx%d/%d" % (size, size, boards)] = row
print(json.dumps(out))
