"""Developer probe (GPU box): us per ply of the headline rollouts vs the scheduling block size (bp0 = static
scheduling, one CTA per tile for the whole launch), all outputs on, 32 and 20 plies per launch."""
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from gymgo_b200.engine import GoEngine
PPLS = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [32, 20]
BPS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 5, 6, 8, 10]
for n, boards in ((9, 65536), (19, 16384)):
    e = GoEngine(n, "cuda:0")
    start = e.new_records(boards)
    e.rollout(start, 0, 0, 0, 256, plies_per_launch=32)
    for dname, dt in (("f32", torch.float32), ("u8", torch.uint8)):
        for ppl in PPLS:
            ring = e.empty((ppl, boards, 6, n, n), dtype=dt)
            acts = e.empty((ppl, boards), dtype=torch.int32); done = e.empty((ppl, boards)); rew = e.empty((ppl, boards), dtype=torch.float32)
            row = {"size": n, "obs": dname, "ppl": ppl}
            for bp in BPS:
                if bp and ppl < 2 * bp: continue
                rec = start.clone()
                kw = dict(plies_per_launch=ppl, obs_ring=ring, actions_log=acts, done_log=done, reward_log=rew, reward_mode=1, dynamic=bool(bp), block_plies=bp)
                e.rollout(rec, 0, 0, 256, ppl, **kw)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for i in range(max(3, 384 // ppl)):
                    e.rollout(rec, 0, 0, 256 + ppl * (i + 1), ppl, **kw)
                ev1.record(); torch.cuda.synchronize()
                row["bp%d" % bp] = round(ev0.elapsed_time(ev1) * 1e3 / (max(3, 384 // ppl) * ppl), 2)
            print(json.dumps(row), flush=True)
            del ring
