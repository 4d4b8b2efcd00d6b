"""Developer probe: loop-trip statistics of the device algorithm (via the host simulator) on steady-state
rollouts, incl. the max over groups of `bpw` boards (what a warp pays)."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import hostsim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 9
boards = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 200
bpw = hostsim.layout(n)["bpw"]
boards = boards // bpw * bpw
lib = hostsim.lib()
recs = hostsim.pack(np.zeros((boards, 6, n, n), dtype=np.uint8))
acts = np.zeros(boards, dtype=np.int32)
stats = np.zeros((boards, 8), dtype=np.int32)
for t in range(warm):
    hostsim.rollout_step(recs, n, 0, 0, t)
acc = []
for t in range(warm, warm + 40):
    lib.hs_rollout_step_stats(n, ctypes.c_void_p(recs.ctypes.data), boards, ctypes.c_uint64(0), ctypes.c_uint64(0),
                              ctypes.c_uint64(t), ctypes.c_void_p(acts.ctypes.data), ctypes.c_void_p(stats.ctypes.data))
    acc.append(stats.copy())
a = np.stack(acc)                      # [T, boards, 2]
w = a.reshape(a.shape[0], -1, bpw, 8).max(axis=2)
print("n=%d boards/warp=%d" % (n, bpw))
print("flood iterations per board-ply: mean %.2f  p90 %d  max %d | per warp (max over boards, lockstep lower bound): mean %.2f"
      % (a[..., 0].mean(), np.percentile(a[..., 0], 90), a[..., 0].max(), w[..., 0].mean()))
print("pocket-loop trips per board-ply: mean %.2f  p90 %d  max %d | per warp (max over %d boards): mean %.2f"
      % (a[..., 1].mean(), np.percentile(a[..., 1], 90), a[..., 1].max(), bpw, w[..., 1].mean()))
names = ["capture: groups at the move", "capture: alive subset", "big groups (to move)", "big groups (just moved)",
         "pocket loop (all trips)"]
for k, name in enumerate(names):
    col = a[..., 2 + k]
    print("  flood iterations %-28s per board mean %.2f p90 %d max %d | warp-max mean %.2f"
          % (name, col.mean(), np.percentile(col, 90), col.max(), w[..., 2 + k].mean()))
