"""SMPL forward + per-vertex statistics on B x N meshes: the fused kernel group vs the staged three-kernel path + statistics
kernel (HP3D_SMPL=staged), CUDA-event timed alone; optionally on a vertex-shuffled model (argv: shuffle block size)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import hierarchicalprobabilistic3dhuman_b200 as hp
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn, _lib

B, N = int(os.environ.get("B", 256)), int(os.environ.get("N", 100))
dev = torch.device("cuda", 0)
models = {"synthetic (part-ordered vertices)": syn.synthetic_smpl_model()}
for blk in [int(a) for a in sys.argv[1:]]:
    models[f"shuffled, blocks of {blk}"] = syn.shuffle_smpl_vertices(syn.synthetic_smpl_model(), seed=3, block=blk)
M = B * N
R = hp.rot6d_to_rotmat(torch.randn(M * 23, 6, device=dev)).view(M, 23, 3, 3)
gR = hp.rot6d_to_rotmat(torch.randn(B, 6, device=dev))
betas = torch.randn(B, 10, device=dev)
verts = torch.empty(M, 6890, 3, device=dev); joints = torch.empty(M, 90, 3, device=dev); unc = torch.empty(B, 6890, device=dev)
L = _lib.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, model in models.items():
    smpl = hp.SMPL(model=model).to(dev)
    h = smpl._handle(dev)
    for mode in ("fused", "staged"):
        os.environ["HP3D_SMPL"] = mode
        ws = torch.empty(L.hp3d_smpl_workspace_bytes(h, M, B), dtype=torch.uint8, device=dev)
        run = lambda: _lib.check(L.hp3d_smpl_forward_stats(h, betas.data_ptr(), B, gR.data_ptr(), B, R.data_ptr(), M, N, verts.data_ptr(),
                                                           joints.data_ptr(), unc.data_ptr(), None, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"model": name, "path": mode, "B": B, "N": N, "ms": ms, "GB/s fused accounting (84,664 B/mesh)": 84664 * M / ms / 1e6,
                          "checksum": float(verts.double().sum()), "unc": float(unc.double().sum())}))
        del ws
