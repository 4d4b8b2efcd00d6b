"""Multi-GPU plumbing for the batched environment: boards are independent, so the batch is block-sharded over
ranks (one process per GPU) and the ONLY collective is an end-of-run all-gather of per-rank counters.
Pure torch.distributed - works with NCCL (GPU ranks) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(global_boards, rank, world):
    """contiguous block sharding: rank g owns boards [g*B/G, (g+1)*B/G) (remainder to the first ranks)"""
    base, rem = divmod(int(global_boards), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_counters(values, device=None):
    """all-gather a small float64 vector from every rank -> tensor [world, len(values)] (on the CPU)"""
    mine = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return mine.cpu().reshape(1, -1)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return torch.stack(out).cpu()


def throughput(counters, plies_col=0, secs_col=1):
    """whole-job env-steps/s = units all ranks processed / the slowest rank's time"""
    return float(counters[:, plies_col].sum()) / float(counters[:, secs_col].max())
