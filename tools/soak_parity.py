"""Soak test: millions of fresh random-legal transitions from the CUDA rollout, every one replayed through the
plain-C oracle (bit-exact next state, legal sampled action, done flag).  Run on the B200 box:
    python tools/soak_parity.py [million_transitions_per_size]"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gymgo_b200.engine import GoEngine  # noqa: E402
from oracle import c_oracle as co  # noqa: E402


def check(args):
    prev, acts, nxt = args
    prev = prev.copy()
    prev[prev[:, 5, 0, 0] == 1] = 0                        # auto-reset precedes the ply
    want, status = co.batch_next_states(prev, acts)
    return int(status.any()), int((want != nxt).any(axis=(1, 2, 3)).sum())


def main():
    target = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    co.build()
    pool = mp.get_context("fork").Pool(min(16, os.cpu_count() or 1))
    for n, boards in ((9, 65536), (9, 32768), (19, 16384), (8, 65536), (7, 65536), (13, 8192), (6, 65536), (5, 32768), (4, 65536),
                      (3, 32768), (16, 4096)):     # 65,536 boards of side <= 9 select k_rollout_tpb, the rest k_rollout
        e = GoEngine(n, "cuda:0")
        chunk = 16
        rec = e.new_records(boards)
        ring = e.empty((chunk, boards, 6, n, n), dtype=torch.uint8)
        acts = torch.empty((chunk, boards), dtype=torch.int32, device="cuda")
        prev = np.zeros((boards, 6, n, n), dtype=np.uint8)
        done = bad_status = mismatches = 0
        t = 0
        t0 = time.time()
        want_total = int(target * 1e6 * (1.0 if n in (9, 19) else 0.25))
        kernel = e.lib.gg_rollout_kernel(n, boards).decode()
        while done < want_total:
            e.rollout(rec, 20240925, 0, t, chunk, plies_per_launch=chunk, actions_log=acts, obs_ring=ring)
            obs = ring.cpu().numpy()
            a = acts.cpu().numpy()
            order = [(t + p) % chunk for p in range(chunk)]
            jobs = []
            for p in range(chunk):
                cur = obs[order[p]]
                for lo in range(0, boards, 4096):
                    jobs.append((prev[lo:lo + 4096], a[p, lo:lo + 4096], cur[lo:lo + 4096]))
                prev = cur
            for st, mm in pool.map(check, jobs):
                bad_status += st
                mismatches += mm
            done += chunk * boards
            t += chunk
        print("N=%2d x %6d  %9d transitions  illegal-action batches %d  state mismatches %d  (%.1f s)  %s"
              % (n, boards, done, bad_status, mismatches, time.time() - t0, kernel), flush=True)
        assert bad_status == 0 and mismatches == 0
    print("soak parity OK")


if __name__ == "__main__":
    main()
