import sys, os, torch
sys.path.insert(0, os.getcwd())
from gymgo_b200.engine import GoEngine
from gymgo_b200 import _cabi
e = GoEngine(9, "cuda:0")
buf = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
s = e._enter()
for run in (512, 2048, 4096, 16384, 65536, 262144, 1 << 20, 4 << 20):
    best = 0
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _cabi.check(e.lib.gg_probe_write(buf.data_ptr(), buf.numel(), run, s)); b.record(); torch.cuda.synchronize()
        best = max(best, buf.numel() / a.elapsed_time(b) / 1e6)
    print(run, round(best, 1), "GB/s")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); buf.zero_(); b.record(); torch.cuda.synchronize(); print("torch zero_", round(buf.numel() / a.elapsed_time(b) / 1e6, 1))
src = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
a.record(); buf[:4 << 30].copy_(src); b.record(); torch.cuda.synchronize(); print("torch copy (r+w bytes)", round(2 * src.numel() / a.elapsed_time(b) / 1e6, 1))
