"""Developer probe for BASELINE.json configs[3]: 9x9 children() full legal-move expansion, 4,096 parents."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymgo_b200.engine import GoEngine  # noqa: E402

e = GoEngine(9, "cuda:0")
parents = e.new_records(4096)
e.rollout(parents, 0, 0, 0, 40, plies_per_launch=8)
e.reset(parents, e.flags(parents) & 4 != 0)
out = {}
for name, kw in (("packed_only", dict(want_rec=True, want_obs=False)), ("f32_dense", dict(want_rec=False, want_obs=True)),
                 ("u8_dense", dict(want_rec=False, want_obs=True, obs_dtype=torch.uint8))):
    for _ in range(3):
        e.children(parents, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        e.children(parents, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    out[name] = dict(us_per_call=round(us, 1), parents_per_s=round(4096 / us * 1e6), child_states_per_s=round(4096 * 82 / us * 1e6))
print(json.dumps(out))
