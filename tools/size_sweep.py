"""Developer probe: rollout throughput by board size (f32 observations, 32 plies per launch, steady state)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymgo_b200.engine import GoEngine  # noqa: E402

out = {}
for size, boards in ((5, 131072), (7, 65536), (9, 65536), (11, 32768), (13, 32768), (15, 16384), (19, 16384)):
    eng = GoEngine(size, "cuda:0")
    rec = eng.new_records(boards)
    ring = eng.empty((3, boards, 6, size, size), dtype=torch.float32)
    eng.rollout(rec, 0, 0, 0, 256, obs_ring=ring)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.rollout(rec, 0, 0, 256, 512, obs_ring=ring)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 512
    byt = boards * 6 * size * size * 4
    out["%dx%d" % (size, size)] = dict(boards=boards, us_per_ply=round(us, 2), env_steps_per_s=round(boards / us * 1e6),
                                       obs_write_GBps=round(byt / us / 1e3), kernel=eng.lib.gg_rollout_kernel(size, boards).decode())
print(json.dumps(out))
