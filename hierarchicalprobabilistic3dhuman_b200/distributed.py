"""Multi-GPU plumbing for the hot path: images shard contiguously over ranks, every rank's
kernels write straight into its slice of pre-allocated gather buffers, and ONE in-place all-gather per
output tensor (NCCL over NVLink on GPUs; gloo in the CPU tests) assembles the global result.
The reference has no distributed code at all; this is new."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [start, stop) of `total` items owned by `rank`; the first `total % world` ranks get one extra."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class GatherBuffers:
    """Per-output gather buffers of shape (world * per_rank, *item_shape); `local(name)` is this rank's slice
    (hand it to the kernels as their output), `all_gather()` completes the other slices in place."""

    def __init__(self, per_rank, specs, device, rank=None, world=None, dtype=torch.float32):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.per_rank = per_rank
        self.full = {k: torch.empty(self.world * per_rank, *shape, device=device, dtype=dtype) for k, shape in specs.items()}

    def local(self, name):
        return self.full[name][self.rank * self.per_rank:(self.rank + 1) * self.per_rank]

    def all_gather(self, names=None):
        if self.world == 1:
            return
        for k in (names or self.full.keys()):
            # in place: the send buffer is this rank's slice of the receive buffer (NCCL in-place semantics)
            dist.all_gather_into_tensor(self.full[k], self.local(k))

    def bytes_received_per_rank(self, names=None):
        return sum(self.full[k][0].numel() * self.per_rank * (self.world - 1) * self.full[k].element_size()
                   for k in (names or self.full.keys()))
