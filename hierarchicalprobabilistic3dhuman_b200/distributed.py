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


class PeerPush:
    """EXPERIMENTAL (opt-in, `bench.py --transport p2p`; crashed with SIGSEGV in its first 2-GPU trial and is not on any
    default path): all-gather by peer-to-peer PUSH over NVLink with the copy engines instead of SM-resident NCCL kernels.

    Every rank owns an identical buffer `full` (same shape on every GPU). The buffers are exchanged once as CUDA IPC
    handles; `push(view_fn, stream)` then copies this rank's slice straight into the same slice of every peer's buffer
    (`cudaMemcpyPeerAsync` -> copy engines), so the transfer neither needs nor blocks SMs -- the hot-path kernels are
    persistent one-CTA-per-SM kernels that an NCCL kernel cannot co-run with. `fence(stream)` enqueues a tiny NCCL
    all-reduce behind the copies: when it completes on a rank, every rank's preceding pushes have landed.
    Falls back (ok == False) if IPC / peer access is unavailable; callers then use `dist.all_gather_into_tensor`."""

    def __init__(self, full, rank, world):
        self.full, self.rank, self.world, self.ok, self.peers = full, rank, world, False, []
        self.flag = torch.zeros(1, device=full.device)
        try:
            from torch.multiprocessing.reductions import reduce_tensor
            handle = reduce_tensor(full)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            for r in range(world):
                if r == rank:
                    self.peers.append(None)
                else:
                    fn, fargs = handles[r]
                    self.peers.append(fn(*fargs))
            ok = torch.ones(1, device=full.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            self.ok = bool(ok.item() == 1)
        except Exception as e:           # noqa: BLE001 -- any failure means "use NCCL instead"
            self.error = repr(e)
            try:
                ok = torch.zeros(1, device=full.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            except Exception:
                pass
            self.ok = False

    def push(self, view_fn, stream):
        """view_fn(buffer) -> the slice this rank produced (same indexing applied to the local and the peer buffers)."""
        src = view_fn(self.full)
        with torch.cuda.stream(stream):
            for r, peer in enumerate(self.peers):
                if peer is not None:
                    view_fn(peer).copy_(src, non_blocking=True)

    def fence(self, stream):
        with torch.cuda.stream(stream):
            dist.all_reduce(self.flag)
