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


class SymmPush:
    """All-gather by peer-to-peer PUSH on the copy engines over torch's symmetric memory (CUDA VMM buffers mapped into
    every rank's address space): `full` has the same shape on every rank; `push(view_fn, stream)` copies this rank's
    slice into the same slice of every peer's buffer with plain device-to-device copies (no SM-resident kernels --
    NCCL's all-gather kernels cannot co-run with the hot path's kernels, which fill every SM: measured at 2 GPUs the
    NCCL gather serialises with the compute, 5.5 -> 8.9 ms per step). `fence(stream)` enqueues a one-element NCCL
    all-reduce behind the copies: when it completes on a rank, every rank's preceding pushes have landed.
    `ok == False` (rendezvous unavailable) means: use `dist.all_gather_into_tensor` instead."""

    def __init__(self, shape, device, rank, world, dtype=torch.float32, mode=None, ctas=None, fence_group=None):
        import os
        # `fence_group`: process group for the completion all-reduce. It must NOT be the group the caller uses for collectives
        # on its main stream: ProcessGroupNCCL runs all collectives of a group on ONE internal stream, so a fence queued
        # behind the pushes would hold up every later collective of that group -- and the stream that waits for it (this is
        # what serialised the vertices gather with the next step's compute in round 1, whatever the transport).
        self.fence_group = fence_group
        # "kernel": hp3d_peer_push -- SM stores over NVLink from tiny CTAs that co-reside with the compute kernels (default);
        # "ce": plain device-to-device copies on the copy engines (no SMs, but ~430 GB/s per rank measured at 4 GPUs)
        # "mc": the same through the NVSwitch multicast mapping (one multimem.st reaches every GPU; falls back to "kernel"
        # when the symmetric-memory handle has no multicast pointer)
        self.mode = mode or os.environ.get("HP3D_PUSH", "mc")
        self.ctas = int(ctas if ctas is not None else os.environ.get("HP3D_PUSH_CTAS", "0"))
        self.rank, self.world, self.ok, self.peers, self.error = rank, world, False, [], None
        self.full, self.mc_ptr = None, 0
        good = 0.0
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.full = symm_mem.empty(*shape, dtype=dtype, device=device)
            self.hdl = symm_mem.rendezvous(self.full, dist.group.WORLD)
            for r in range(world):
                self.peers.append(None if r == rank else self.hdl.get_buffer(r, tuple(shape), dtype))
            self.mc_ptr = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
            good = 1.0
        except Exception as e:           # noqa: BLE001 -- any failure means "use NCCL instead"
            self.error = repr(e)
        ok = torch.tensor([good, 1.0 if self.mc_ptr else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self.ok = bool(ok[0].item() == 1)
        if self.mode == "mc" and ok[1].item() != 1:
            self.mode = "kernel"                    # no NVLS multicast mapping on some rank
        if not self.ok:
            self.peers = []
            if self.full is None:
                self.full = torch.empty(*shape, dtype=dtype, device=device)
        self.flag = torch.zeros(1, device=device)
        self.num_streams, self._streams = 4, None

    def push(self, view_fn, stream):
        """Copies ordered after the work already enqueued on `stream`; `stream` then waits for them. One device-to-device
        copy runs on one copy engine (~320 GB/s measured), so the (peer, sub-slice) copies are spread over
        `num_streams` streams to keep several engines -- and all NVLink lanes -- busy."""
        if not self.ok:
            return                                  # fallback mode: the caller gathers with NCCL / gloo
        src = view_fn(self.full)
        if self.mode == "mc":
            import ctypes
            from . import _lib
            assert src.is_contiguous()
            off = src.data_ptr() - self.full.data_ptr()
            with torch.cuda.device(self.full.device):
                _lib.check(_lib.lib().hp3d_peer_push_multicast(src.data_ptr(), ctypes.c_void_p(self.mc_ptr + off),
                                                               src.numel() * src.element_size(), self.ctas,
                                                               ctypes.c_void_p(stream.cuda_stream)), "hp3d_peer_push_multicast")
            return
        if self.mode == "kernel":
            import ctypes
            from . import _lib
            dsts = [view_fn(p) for p in self.peers if p is not None]
            assert src.is_contiguous() and all(d.is_contiguous() for d in dsts)
            arr = (ctypes.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts])
            with torch.cuda.device(self.full.device):
                _lib.check(_lib.lib().hp3d_peer_push(src.data_ptr(), arr, len(dsts), src.numel() * src.element_size(), self.ctas,
                                                     ctypes.c_void_p(stream.cuda_stream)), "hp3d_peer_push")
            return
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.full.device) for _ in range(self.num_streams)]
        start = torch.cuda.Event()
        start.record(stream)
        peers = [p for p in self.peers if p is not None]
        parts = max(1, -(-self.num_streams // max(1, len(peers))))
        parts = min(parts, src.shape[0])
        jobs = []
        for peer in peers:
            dst = view_fn(peer)
            for sp, dp in zip(src.chunk(parts, dim=0), dst.chunk(parts, dim=0)):
                jobs.append((sp, dp))
        used = set()
        for i, (sp, dp) in enumerate(jobs):
            st = self._streams[i % self.num_streams]
            if i < self.num_streams:
                st.wait_event(start)
            with torch.cuda.stream(st):
                dp.copy_(sp, non_blocking=True)
            used.add(i % self.num_streams)
        for i in used:
            done = torch.cuda.Event()
            done.record(self._streams[i])
            stream.wait_event(done)

    def fence(self, stream):
        with torch.cuda.stream(stream):
            dist.all_reduce(self.flag, group=self.fence_group)
