"""N>1 host logic on CPU: world_size-2 gloo run of the shard + in-place all-gather plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hierarchicalprobabilistic3dhuman_b200.distributed import shard_range, GatherBuffers, SymmPush


def test_shard_range_partitions_exactly():
    for total in (1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, N = 3, 4                                   # images per rank, samples per image
        gb = GatherBuffers(B, {"rotmats": (N, 23, 3, 3), "betas": (10,), "vertices": (N, 50, 3)}, "cpu")
        # stand-in for the kernels: each rank fills ITS slice only, as a function of the global image index
        g0 = rank * B
        for name in gb.full:
            gb.full[name].fill_(float("nan"))
            loc = gb.local(name)
            for i in range(B):
                loc[i] = (g0 + i) + 0.001 * torch.arange(loc[i].numel(), dtype=torch.float32).view(loc[i].shape)
        gb.all_gather()
        ok = True
        for name, t in gb.full.items():
            for i in range(world * B):
                exp = i + 0.001 * torch.arange(t[i].numel(), dtype=torch.float32).view(t[i].shape)
                ok &= bool(torch.equal(t[i], exp))
        # the local slice aliases the gather buffer (no pack/copy)
        ok &= gb.local("betas").data_ptr() == gb.full["betas"][rank * B:].data_ptr()
        # copy-engine transport: without CUDA symmetric memory every rank must agree on the fallback (ok == False on all
        # ranks, a plain buffer of the requested shape, push() a no-op) instead of dead-locking or diverging
        sp = SymmPush((2, world, 3, 5), "cpu", rank, world)
        ok &= (not sp.ok) and tuple(sp.full.shape) == (2, world, 3, 5) and sp.error is not None
        sp.push(lambda buf: buf[0, rank], None)
        # the completion fence runs on a communicator of its OWN (bench.py: collectives issued on the comm stream must not share
        # ProcessGroupNCCL's internal stream with the main stream's collectives): the group is plumbed through and used
        own = dist.new_group(backend="gloo")
        sp2 = SymmPush((2, world, 3, 5), "cpu", rank, world, mode="ce", fence_group=own)
        sp2.flag.fill_(float(rank + 1))
        dist.all_reduce(sp2.flag, group=sp2.fence_group)           # what SymmPush.fence() issues (on a CUDA stream there)
        ok &= sp2.fence_group is own and float(sp2.flag) == sum(range(1, world + 1)) and sp2.mode == "ce"
        q.put((rank, ok, gb.bytes_received_per_rank(["vertices"])))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_inplace_all_gather_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok, _ in res), res
    assert all(b == 3 * 4 * 50 * 3 * 4 for _, _, b in res)
