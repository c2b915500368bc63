"""Micro-benchmark of the SMPL-LBS kernel alone (FK + skinning + joints) at the bench size.
HP3D_LBS / HP3D_LBS_MODE select the kernel variant; prints ms, GB/s (algorithmic 167,592 B/mesh) and frac of measured peak."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import hierarchicalprobabilistic3dhuman_b200 as hp
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn, _lib

B, N = int(os.environ.get("B", 256)), int(os.environ.get("N", 100))
M = B * N
dev = torch.device("cuda", 0)
smpl = hp.SMPL(model=syn.synthetic_smpl_model()).to(dev)
L = _lib.lib(); h = smpl._handle(dev)
vp = torch.empty(M, 20672, device=dev).normal_()
J = torch.randn(B, 24, 3, device=dev)
gR = hp.rot6d_to_rotmat(torch.randn(B, 6, device=dev))
R = hp.rot6d_to_rotmat(torch.randn(M * 23, 6, device=dev)).view(M, 23, 3, 3)
verts = torch.empty(M, 6890, 3, device=dev); joints = torch.empty(M, 90, 3, device=dev)
call = lambda: _lib.check(L.hp3d_smpl_lbs(h, vp.data_ptr(), J.data_ptr(), B, gR.data_ptr(), B, R.data_ptr(), M, verts.data_ptr(), joints.data_ptr(), None))
for _ in range(3): call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) else 6650.0
gbs = 167592 * M / (ms * 1e-3) / 1e9
print(json.dumps({"variant": os.environ.get("HP3D_LBS", "tile") + ":" + os.environ.get("HP3D_LBS_MODE", "default"), "ms": ms, "GBps": gbs, "frac": gbs / peak}))
