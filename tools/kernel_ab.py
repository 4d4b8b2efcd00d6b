"""Developer probe (GPU box): the two persistent rollout kernels side by side.

    python tools/kernel_ab.py [--out gpurun_out/kernel_ab.json] [--quick]

For every (board size, batch, observation dtype) the same rollout (seed 0, 256-ply pre-roll so that game phases are
de-synchronised) is continued with each kernel for `plies` plies, `ppl` plies per launch into a ring of `ppl` observation
slots (every observation of a launch stays readable), timed with CUDA events.  Reports us per ply and the fraction of
the measured HBM peak.  Results are bit-identical across kernels (tests/test_gpu_parity.py); this only measures."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gymgo_b200 import _cabi  # noqa: E402
from gymgo_b200.engine import GoEngine  # noqa: E402

NAMES = {0: "lanes", 1: "thread"}


def bytes_per_ply(n, elem):
    p = 4 * ((n * n + 31) // 32)
    return 2 * (3 * p + 4) + 4 + 6 * n * n * elem


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernel_ab.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--ppl", type=int, default=32)
    ap.add_argument("--lib", default=None, help="developer A/B: load this build of libgymgo_b200.so instead of the in-tree one")
    args = ap.parse_args()
    if args.lib:
        from gymgo_b200 import build as _b
        _b.LIB = os.path.abspath(args.lib)
        _b.up_to_date = lambda: True
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        peak = 6650.0
    cases = [(9, 65536), (19, 16384), (13, 32768), (7, 65536), (9, 16384), (9, 32768), (9, 131072), (19, 4096), (19, 65536), (5, 131072), (11, 32768), (15, 16384)]
    if args.quick:
        cases = cases[:2]
    dtypes = (("f32", torch.float32, 4), ("u8", torch.uint8, 1), ("bf16", torch.bfloat16, 2), ("none", None, 0))
    out = {"peak_gbs": peak, "plies_per_launch": args.ppl, "rows": []}
    for n, boards in cases:
        e = GoEngine(n, "cuda:0")
        start = e.new_records(boards)
        e.rollout(start, 0, 0, 0, 256, plies_per_launch=32)
        for dname, dt, elem in dtypes:
            ring = None if dt is None else e.empty((args.ppl, boards, 6, n, n), dtype=dt)
            row = {"size": n, "boards": boards, "obs": dname}
            for k in (0, 1):
                if k == 1 and n > 9 and boards * n * n > 16384 * 361 // 2 and dname != "f32":
                    continue                                    # thread-per-board on big boards: slow, sample f32 only
                for mode, dyn, bp in (("_static", False, 0), ("", True, 0)):
                    rec = start.clone()
                    plies = 10 * args.ppl
                    e.rollout(rec, 0, 0, 256, args.ppl, plies_per_launch=args.ppl, obs_ring=ring, kernel=k, dynamic=dyn, block_plies=bp)
                    torch.cuda.synchronize()
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record()
                    e.rollout(rec, 0, 0, 256 + args.ppl, plies, plies_per_launch=args.ppl, obs_ring=ring, kernel=k, dynamic=dyn, block_plies=bp)
                    ev1.record()
                    torch.cuda.synchronize()
                    us = ev0.elapsed_time(ev1) * 1e3 / plies
                    row[NAMES[k] + mode] = round(us, 2)
                    if elem:
                        row[NAMES[k] + mode + "_frac"] = round(boards * bytes_per_ply(n, elem) / (us * 1e-6) / 1e9 / peak, 3)
            print(json.dumps(row), flush=True)
            out["rows"].append(row)
            del ring
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
