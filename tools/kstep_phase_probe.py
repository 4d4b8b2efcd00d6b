"""Developer probe (GPU box): one host-driven ply (gg_step, one launch) with no / u8 / bf16 / f32 observation: if the
rules and the stores of ONE launch do not overlap, t(f32) ~ t(none) + bytes / HBM write rate."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymgo_b200.engine import GoEngine  # noqa: E402

out = []
for n, b in ((9, 65536), (19, 16384)):
    e = GoEngine(n, "cuda:0")
    rec = e.new_records(b)
    e.rollout(rec, 0, 0, 0, 256, plies_per_launch=32)
    seq, r3 = [], rec.clone()
    for t in range(60):                                   # a legal action sequence for the boards (auto-reset on)
        a = e.sample_legal(r3, 7, 0, t)
        e.step(r3, a, out=r3, auto_reset=True, want_status=False)
        seq.append(a)
    for name, dt, elem in (("none", None, 0), ("u8", torch.uint8, 1), ("bf16", torch.bfloat16, 2), ("f32", torch.float32, 4)):
        r2 = rec.clone()
        obs = None if dt is None else e.empty((b, 6, n, n), dtype=dt)
        st = e.empty((b,))
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for t in range(60):
            if t == 10:
                ev0.record()
            e.step(r2, seq[t], out=r2, auto_reset=True, obs=obs, want_status=False)
        ev1.record()
        torch.cuda.synchronize()
        assert torch.equal(r2, r3)
        us = ev0.elapsed_time(ev1) * 1e3 / 50
        out.append({"size": n, "boards": b, "obs": name, "us_per_step": round(us, 2),
                    "obs_bytes_over_7200_gbs_us": round(b * 6 * n * n * elem / 7.2e12 * 1e6, 2)})
        print(json.dumps(out[-1]), flush=True)
