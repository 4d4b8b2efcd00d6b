"""Developer probe: does a device->host DMA of dense observations run concurrently with the host codec without slowing
it down?  (If yes, a hybrid transport - part of the batch dense over PCIe, the rest packed + codec - adds the two rates.)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from gymgo_b200 import _cabi, hostmem  # noqa: E402
from gymgo_b200.envs import BatchedGoEnv  # noqa: E402

n, b = 9, 65536
env = BatchedGoEnv(b, n)
for _ in range(100):
    env.random_step()
lib = env.engine.lib
rec_h = hostmem.pinned_empty(tuple(env.rec.shape), torch.uint8, 0)
rec_h.copy_(env.rec)
obs_h = hostmem.pinned_empty(tuple(env.obs.shape), torch.float32, 0)
dma_h = hostmem.pinned_empty(tuple(env.obs.shape), torch.float32, 0)
st = torch.cuda.Stream()
torch.cuda.synchronize()
threads = hostmem.codec_threads()
out = {}
for frac in (0.0, 0.1, 0.2, 0.3, 0.4):
    k = int(b * frac) // 32 * 32          # boards sent dense over PCIe; the codec expands the other b - k
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        if k:
            with torch.cuda.stream(st):
                dma_h[:k].copy_(env.obs[:k], non_blocking=True)
        _cabi.check(lib.gg_host_unpack(rec_h[k:].data_ptr(), b - k, n, _cabi.GG_F32, obs_h[k:].data_ptr(), threads))
        t1 = time.perf_counter()
        st.synchronize()
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t0))
    ts.sort(key=lambda x: x[1])
    med = ts[len(ts) // 2]
    out["dense_fraction_%.1f" % frac] = {"codec_ms": round(med[0] * 1e3, 3), "both_done_ms": round(med[1] * 1e3, 3),
                                         "total_gbs": round(b * 6 * n * n * 4 / med[1] / 1e9, 1)}
print(json.dumps(out, indent=1))
