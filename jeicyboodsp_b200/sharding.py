"""Multi-GPU plumbing.  Streams / sources / utterances are independent, so the hot path
shards with NO data-path collective: rank r owns a contiguous slice of the leading index and its own
per-stream state.  The only collective anywhere is the optional all-gather of per-GPU MFCC feature blocks
when a caller wants a single feature matrix (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of `n_units` owned by `rank`; the first n_units % world ranks get one extra."""
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Job time is the slowest rank's device time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allgather_features(local, n_units_total: int):
    """Gather per-rank MFCC blocks [units_r, frames, n_cep] into [n_units_total, frames, n_cep] on every rank.
    Shards may differ by one unit (shard_range), so blocks are padded to the largest shard for the fixed-size
    all_gather_into_tensor and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_units_total, r, world) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * biggest: r * biggest + (e - b)] for r, (b, e) in enumerate(sizes)]
    return torch.cat(parts, dim=0)
