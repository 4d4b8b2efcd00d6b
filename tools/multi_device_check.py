"""Developer check: the library (own static CUDA runtime) works on a non-zero device ordinal and on two
devices from one process."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gymgo_b200.engine import GoEngine
from oracle import c_oracle as co

assert torch.cuda.device_count() >= 2
for dev in ("cuda:1", "cuda:0", "cuda:1"):
    e = GoEngine(9, dev)
    rec = e.new_records(500)
    acts = e.empty((500,), dtype=torch.int32)
    obs = e.empty((500, 6, 9, 9), dtype=torch.uint8)
    dense = np.zeros((500, 6, 9, 9), dtype=np.uint8)
    for t in range(40):
        e.rollout_step(rec, 3, 0, t, actions=acts, obs=obs)
        dense[dense[:, 5, 0, 0] == 1] = 0
        dense, st = co.batch_next_states(dense, acts.cpu().numpy())
        assert not st.any() and np.array_equal(obs.cpu().numpy(), dense)
    assert rec.device == torch.device(dev)
    assert torch.cuda.current_device() == int(dev[-1])      # engines switch the current device (like set_device)
    print(dev, "ok")
