"""Batch sharding across the GPUs of one box.

Images are independent (no state is shared between CWavelet2D instances), so a batch is split
into contiguous chunks of ceil(N/ngpu) images, one process per GPU, with NO data-path collective.
The only cross-rank communication is the max-over-ranks of the device-timed duration.
"""
import torch
import torch.distributed as dist


def shard(n_images: int, rank: int, world: int):
    """Contiguous [begin, end) image range of `rank` out of `world` ranks."""
    per = (n_images + world - 1) // world
    b = min(rank * per, n_images)
    return b, min(b + per, n_images)


def max_over_ranks(value_ms: float, device="cpu") -> float:
    """Max of a per-rank duration over the default process group (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank: int, local_ms: float, device="cpu") -> float:
    """Whole-job units per second: all ranks' units / slowest rank's time."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    total = torch.tensor([float(units_per_rank)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return float(total.item()) / (max_over_ranks(local_ms, device) * 1e-3)
