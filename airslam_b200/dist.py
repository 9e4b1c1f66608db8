"""Multi-GPU plumbing: stereo pairs are independent, so the path shards pairs over ranks with NO data-path
collective; torch.distributed (NCCL on the GPU box, gloo in the CPU tests) only provides the barrier, the max-over-ranks
reduction of timings and -- for the batched-relocalization config -- one all-gather of query features."""
import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(total, rank, world):
    """Contiguous block partition of `total` units (pairs / keyframes): returns (begin, end) of this rank; sizes differ by <= 1."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_owner(index, total, world):
    """Inverse of shard_range: which rank owns unit `index`."""
    base, rem = divmod(total, world)
    cut = rem * (base + 1)
    return index // (base + 1) if index < cut else rem + (index - cut) // max(base, 1)


def max_over_ranks(x, device=None):
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device=None):
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, seconds_this_rank, device=None):
    """Whole-job throughput: all units processed by all ranks / the slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)


def all_gather_features(feat, counts):
    """Relocalization exchange (config 5): every rank contributes its query feature sets, padded to a common capacity.
    feat: torch tensor [Q_local, cap, 259]; counts: int32 tensor [Q_local].  Returns (feat_all [world*Q_local, cap, 259], counts_all)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return feat, counts
    w = dist.get_world_size()
    out_f = [torch.empty_like(feat) for _ in range(w)]
    out_c = [torch.empty_like(counts) for _ in range(w)]
    dist.all_gather(out_f, feat)
    dist.all_gather(out_c, counts)
    return torch.cat(out_f, 0), torch.cat(out_c, 0)
