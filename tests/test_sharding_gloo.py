"""N>1 path on CPU: world_size-2 gloo processes run the sharded rollout with the host simulator of the device
algorithm (board offsets, Philox keyed by global board index) and the end-of-run all-gather; the shards must
reproduce the single-process run bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, boards, plies, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hostsim
    from gymgo_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(boards, rank, world)
    recs = hostsim.pack(np.zeros((hi - lo, 6, n, n), dtype=np.uint8))
    acts = [hostsim.rollout_step(recs, n, 7, lo, t) for t in range(plies)]
    np.save(os.path.join(tmpdir, "rec%d.npy" % rank), recs)
    np.save(os.path.join(tmpdir, "act%d.npy" % rank), np.stack(acts))
    counters = sharding.gather_counters([(hi - lo) * plies, 1.0 + rank])
    if rank == 0:
        assert counters.shape == (world, 2)
        assert sharding.throughput(counters) == boards * plies / float(world)     # slowest rank: `world` seconds
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from gymgo_b200 import sharding
    for boards in (1, 7, 64, 65536, 131072):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(boards, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == boards
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


@pytest.mark.parametrize("n", (5, 9))
def test_two_rank_rollout_matches_single_process(tmp_path, n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hostsim
    boards, plies, world = 37, 60, 2
    port = 29500 + (os.getpid() % 2000) + n
    mp.spawn(_worker, args=(world, port, n, boards, plies, str(tmp_path)), nprocs=world, join=True)
    whole = hostsim.pack(np.zeros((boards, 6, n, n), dtype=np.uint8))
    acts = np.stack([hostsim.rollout_step(whole, n, 7, 0, t) for t in range(plies)])
    rec = np.concatenate([np.load(tmp_path / ("rec%d.npy" % r)) for r in range(world)])
    act = np.concatenate([np.load(tmp_path / ("act%d.npy" % r)) for r in range(world)], axis=1)
    assert np.array_equal(rec, whole) and np.array_equal(act, acts)
