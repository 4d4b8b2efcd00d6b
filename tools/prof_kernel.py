"""Workload for ncu captures of ONE rollout kernel at steady state (GPU box).

    ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 9 -c 1 -o gpurun_out/prof \
        python tools/prof_kernel.py --size 19 --boards 16384 --kernel 0

Pre-rolls 256 plies (8 launches of 32), then launches the chosen kernel twice more with the observation ring: the
capture above skips the 8 pre-roll launches and the first ring launch and takes the second one."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymgo_b200.engine import GoEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=19)
ap.add_argument("--boards", type=int, default=16384)
ap.add_argument("--kernel", type=int, default=-1)
ap.add_argument("--obs", default="f32", choices=["f32", "u8", "bf16", "none"])
ap.add_argument("--ppl", type=int, default=32)
args = ap.parse_args()
dt = {"f32": torch.float32, "u8": torch.uint8, "bf16": torch.bfloat16, "none": None}[args.obs]
e = GoEngine(args.size, "cuda:0")
rec = e.new_records(args.boards)
ring = None if dt is None else e.empty((args.ppl, args.boards, 6, args.size, args.size), dtype=dt)
acts = e.empty((args.ppl, args.boards), dtype=torch.int32)
done = e.empty((args.ppl, args.boards))
rew = e.empty((args.ppl, args.boards), dtype=torch.float32)
e.rollout(rec, 0, 0, 0, 256, plies_per_launch=32, kernel=args.kernel)
for i in range(2):
    e.rollout(rec, 0, 0, 256 + i * args.ppl, args.ppl, plies_per_launch=args.ppl, obs_ring=ring, actions_log=acts,
              done_log=done, reward_log=rew, reward_mode=1, kernel=args.kernel)
torch.cuda.synchronize()
print("ok")
