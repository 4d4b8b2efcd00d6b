"""Where the time of one host-buffer step goes (GPU box): wall-clock of the phases of HostStepper.step() for the packed
transport - action hand-over, graph launch, wait for the GPU (H2D + kernel + D2H of the records), host codec.

    python tools/e2e_phases.py [--size 9] [--boards 65536] [--obs f32]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from gymgo_b200 import _cabi  # noqa: E402
from gymgo_b200.envs import BatchedGoEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=9)
ap.add_argument("--boards", type=int, default=65536)
ap.add_argument("--obs", default="f32")
args = ap.parse_args()
dt = {"f32": torch.float32, "u8": torch.uint8, "bf16": torch.bfloat16}[args.obs]
env = BatchedGoEnv(args.boards, args.size, obs_dtype=dt)
for _ in range(100):
    env.random_step()
hs = env.host_stepper(returns="obs", transport="packed", follow_current_stream=False)
acts = env.engine.sample_legal(env.rec, 1, 0, 0).cpu().numpy()
act_np = hs.actions.numpy()
torch.cuda.synchronize()
for _ in range(10):
    hs.step()
lib = env.engine.lib
T = {k: [] for k in ("copy_actions", "enter", "replay", "gpu_wait", "codec", "whole_step_call")}
for it in range(200):
    t0 = time.perf_counter()
    np.copyto(act_np, acts)
    t1 = time.perf_counter()
    env.engine._enter()
    t2 = time.perf_counter()
    with torch.cuda.stream(hs._stream):
        hs._graph.replay()
    t3 = time.perf_counter()
    hs._stream.synchronize()
    t4 = time.perf_counter()
    _cabi.check(lib.gg_host_unpack(hs.rec.data_ptr(), env.batch_size, env.size, hs._gg_dtype, hs.obs.data_ptr(), hs.threads))
    t5 = time.perf_counter()
    hs.step()
    t6 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
        T[k].append(v)
out = {"size": args.size, "boards": args.boards, "obs": args.obs, "threads": hs.threads,
       "median_us": {k: round(1e6 * sorted(v)[len(v) // 2], 1) for k, v in T.items()},
       "mean_us": {k: round(1e6 * sum(v) / len(v), 1) for k, v in T.items()}}
print(json.dumps(out, indent=1))
