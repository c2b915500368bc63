"""Parity + timing of the experimental CTA-pair blend kernel (HP3D_BLEND=pair) against the default blend kernel.
Run on a B200: `HP3D_BLEND=pair python tools/check_blend_pair.py` (the variant is chosen when the SMPL handle is created)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import hierarchicalprobabilistic3dhuman_b200 as hp
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
from oracle.smpl_oracle import SMPLOracle

variant = os.environ.get("HP3D_BLEND", "default")
model = syn.synthetic_smpl_model()
smpl = hp.SMPL(model=model).cuda()
out = {"variant": variant}
for B, N in ((3, 7), (5, 100)):          # 21 meshes (one partial pair) and 500 meshes (odd number of 128-mesh tiles)
    torch.manual_seed(B)
    R = hp.rot6d_to_rotmat(torch.randn(B * N * 23, 6, device="cuda")).view(B * N, 23, 3, 3)
    gR = hp.rot6d_to_rotmat(torch.randn(B, 6, device="cuda"))
    betas = torch.randn(B, 10, device="cuda") * 1.25
    o = smpl(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
    ref = SMPLOracle(model, torch.float64).forward(betas.cpu().repeat_interleave(N, 0), R.cpu(), gR.cpu().repeat_interleave(N, 0)[:, None])
    err = ((o.vertices.double().cpu() - ref["vertices"]).abs().max() / ref["vertices"].abs().max()).item()
    out[f"rel_err_M{B * N}"] = err
    assert err < 1e-4, out
M = 25600
R = hp.rot6d_to_rotmat(torch.randn(M * 23, 6, device="cuda")).view(M, 23, 3, 3)
gR = hp.rot6d_to_rotmat(torch.randn(256, 6, device="cuda"))
betas = torch.randn(256, 10, device="cuda")
for _ in range(3):
    smpl(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    smpl(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
e1.record(); torch.cuda.synchronize()
out["smpl_forward_ms_M25600"] = e0.elapsed_time(e1) / 5
print(json.dumps(out))
