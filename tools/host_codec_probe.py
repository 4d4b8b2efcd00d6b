"""gg_host_unpack throughput on this box: GB/s of dense output vs worker threads, pinned and pageable destinations,
next to a plain torch copy of the same size (the host-memory write ceiling a single stream of stores reaches).

    python tools/host_codec_probe.py [--size 9] [--boards 65536]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from gymgo_b200 import _cabi, hostmem  # noqa: E402
from gymgo_b200.engine import GoEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=9)
    ap.add_argument("--boards", type=int, default=65536)
    args = ap.parse_args()
    n, b = args.size, args.boards
    eng = GoEngine(n)
    rec = eng.new_records(b)
    eng.rollout(rec, 0, 0, 0, 100, plies_per_launch=20)
    want = eng.unpack(rec, dtype=torch.float32).cpu()
    rec_h = rec.cpu()
    lib = _cabi.lib()
    out = {"size": n, "boards": b, "path": lib.gg_host_unpack_path().decode(), "usable_cores": hostmem.usable_cores(), "rows": []}
    dests = {"pinned": hostmem.pinned_empty((b, 6, n, n), torch.float32, 0), "pageable": torch.empty((b, 6, n, n))}
    nbytes = want.numel() * 4
    for name, dst in dests.items():
        for threads in (1, 2, 4, 8, 12, 16, 24, 32):
            if threads > 2 * hostmem.usable_cores():
                break
            best = 1e9
            for _ in range(6):
                t0 = time.perf_counter()
                _cabi.check(lib.gg_host_unpack(rec_h.data_ptr(), b, n, _cabi.GG_F32, dst.data_ptr(), threads))
                best = min(best, time.perf_counter() - t0)
            assert torch.equal(dst, want)
            out["rows"].append({"dest": name, "threads": threads, "ms": best * 1e3, "dense_gbs": nbytes / best / 1e9})
    src = torch.ones((b, 6, n, n))
    best = 1e9
    for _ in range(6):
        t0 = time.perf_counter()
        dests["pageable"].copy_(src)
        best = min(best, time.perf_counter() - t0)
    out["torch_copy_gbs_written"] = nbytes / best / 1e9
    # what a stepping loop sees: the call repeated with other host work in between (mean, not best)
    dst = dests["pinned"]
    small_src, small_dst = torch.zeros(b, dtype=torch.int32), torch.zeros(b, dtype=torch.int32)
    threads = hostmem.codec_threads()

    def loop(between, reps=60):
        ts = []
        for _ in range(reps):
            between()
            t0 = time.perf_counter()
            _cabi.check(lib.gg_host_unpack(rec_h.data_ptr(), b, n, _cabi.GG_F32, dst.data_ptr(), threads))
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return {"mean_ms": 1e3 * sum(ts) / len(ts), "median_ms": 1e3 * ts[len(ts) // 2], "best_ms": 1e3 * ts[0]}

    ev = torch.cuda.Event()

    def gpu_wait():
        rec.add_(0)
        ev.record()
        ev.synchronize()

    out["loop_threads"] = threads
    out["loop"] = {"back_to_back": loop(lambda: None), "sleep_200us": loop(lambda: time.sleep(200e-6)),
                   "sleep_2ms": loop(lambda: time.sleep(2e-3)),
                   "torch_copy_256KB": loop(lambda: small_dst.copy_(small_src)),
                   "gpu_kernel_and_sync": loop(gpu_wait)}
    torch.set_num_threads(1)
    out["loop"]["torch_copy_256KB_1_torch_thread"] = loop(lambda: small_dst.copy_(small_src))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
