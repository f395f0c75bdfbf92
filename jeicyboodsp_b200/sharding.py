"""Multi-GPU plumbing.  Streams / sources / utterances are independent, so the hot path
shards with NO data-path collective: rank r owns a contiguous slice of the leading index and its own
per-stream state.  The only collective anywhere is the optional all-gather of per-GPU MFCC feature blocks
when a caller wants a single feature matrix (NCCL over NVLink on GPUs, gloo in the CPU tests) -- or, without any
collective, the scatter form of the MFCC kernel writing straight into every GPU's copy of the matrix (PeerMatrix)."""
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


class PeerMatrix:
    """One [n_units_total, row_floats] float32 matrix per rank, every rank's copy mapped into all the other ranks of the box
    (CUDA IPC over NVLink).  `dests(unit0)` are the addresses at which unit `unit0` lies in each copy: handed to
    MfccPlan.run_scatter, every rank's kernel writes its own block into ALL copies, so that once all kernels have finished
    every rank holds the whole matrix -- the fused replacement of kernel + all-gather.  Creation and close() are collective
    calls (every rank of the default process group)."""

    def __init__(self, ctx, n_units_total: int, row_floats: int):
        import torch.distributed as dist
        self.ctx, self.n_units, self.row = ctx, n_units_total, row_floats
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.nbytes = n_units_total * row_floats * 4
        self.local = ctx.malloc(self.nbytes)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, ctx.peer_export(self.local))
        self.addrs = [self.local if r == self.rank else ctx.peer_open(handles[r]) for r in range(self.world)]

    def dests(self, unit0: int) -> list[int]:
        off = unit0 * self.row * 4
        # own copy first: its stores stay on this GPU and do not queue behind the NVLink ones
        order = [self.rank] + [r for r in range(self.world) if r != self.rank]
        return [self.addrs[r] + off for r in order]

    def tensor(self):
        """This rank's copy as a torch tensor [n_units_total, row_floats] (no copy)."""
        import torch

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.n_units, self.row), "typestr": "<f4", "data": (self.local, False), "version": 3, "strides": None}
        return torch.as_tensor(raw, device=f"cuda:{torch.cuda.current_device()}")

    def close(self, collective: bool = True) -> None:
        """collective=False: tear down this rank's side only (construction failed on another rank; nothing was written)."""
        import torch.distributed as dist
        self.ctx.sync()              # this rank's kernels (the writers into the peers' copies) have finished
        if not collective:
            self.world = 1
        if self.world > 1:
            dist.barrier()           # nobody unmaps or frees while a peer may still be writing
        for r, a in enumerate(self.addrs):
            if r != self.rank:
                self.ctx.peer_close(a)
        if self.world > 1:
            dist.barrier()           # every mapping of this rank's copy is gone before it is freed
        self.ctx.free(self.local)
        self.addrs, self.local = [], 0


class MulticastMatrix:
    """The same matrix in torch's symmetric memory, which also binds all copies to ONE NVSwitch multicast address: `dest(unit0)`
    is handed to MfccPlan.run_multicast, whose kernel stores each feature row once and lets the switch replicate it into every
    GPU's copy.  `available()` is False where the platform has no multicast (then use PeerMatrix).  Collective constructor."""

    def __init__(self, n_units_total: int, row_floats: int, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.n_units, self.row = n_units_total, row_floats
        self.t = symm_mem.empty(n_units_total * row_floats, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.t, dist.group.WORLD.group_name)
        self.mc = int(self.hdl.multicast_ptr)

    def available(self) -> bool:
        return self.mc != 0

    def dest(self, unit0: int) -> int:
        return self.mc + unit0 * self.row * 4

    def tensor(self):
        return self.t.view(self.n_units, self.row)
