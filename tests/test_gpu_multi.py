"""Multi-GPU checks (need >= 2 visible GPUs; skipped otherwise -- run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu`):
  * SURVEY.md §4 last row / §8e: the tensors gathered from a 2-rank sharded run are BIT-IDENTICAL to a single-GPU run on
    the concatenated batch (Philox keyed by the global image index, batch-invariant encoder tiles);
  * one process driving two devices (per-device shared-memory opt-in, per-device handles and workspaces)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]
B_PER_RANK, N = 4, 8


def _pipeline(dev, batch, image_offset):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from conftest import reference_config
    net = hp.PoseMFShapeGaussianNet(syn.SMPL_PARENTS.tolist(), reference_config())
    net.load_state_dict(syn.synthetic_state_dict(0))
    net = net.to(dev).eval()
    smpl = hp.SMPL(model=syn.synthetic_smpl_model()).to(dev)
    return hp.HotPathPipeline(net, smpl, batch, N, dev, image_offset=image_offset)


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    x_all = torch.from_numpy(syn.synthetic_proxy_rep(world * B_PER_RANK, seed=21))
    pipe = _pipeline(dev, B_PER_RANK, rank * B_PER_RANK)
    torch.manual_seed(99)                                   # the SAME generator state on every rank
    res = pipe.run_device(x_all[rank * B_PER_RANK:(rank + 1) * B_PER_RANK].to(dev))
    gathered = {}
    for k in ("rotmats", "vertices", "joints", "uncertainty", "mode_vertices"):
        t = res[k].contiguous()
        full = torch.empty(world * t.shape[0], *t.shape[1:], device=dev)
        dist.all_gather_into_tensor(full, t)
        gathered[k] = full
    if rank == 0:
        ref_pipe = _pipeline(dev, world * B_PER_RANK, 0)
        torch.manual_seed(99)
        ref = ref_pipe.run_device(x_all.to(dev))
        same = {k: bool(torch.equal(gathered[k], ref[k].reshape(gathered[k].shape))) for k in gathered}
        torch.save(same, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_is_bit_identical_to_single_gpu(built_lib, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "same.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    same = torch.load(out)
    assert all(same.values()), same


def test_one_process_two_devices(built_lib):
    """cudaFuncSetAttribute is per device -- the > 48 KB kernels must launch on the second device too."""
    xs = torch.from_numpy(syn.synthetic_proxy_rep(2, seed=5))
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        pipe = _pipeline(dev, 2, 0)
        with torch.cuda.device(dev):
            torch.manual_seed(5)
            res = pipe.run_device(xs.to(dev))
            torch.cuda.synchronize(dev)
        outs.append({k: res[k].cpu() for k in ("rotmats", "vertices", "uncertainty")})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
