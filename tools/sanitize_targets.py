"""Small invocations of the kernels SURVEY.md §5 names as compute-sanitizer targets: the sampler's ballot compaction
(`mf_sample_kernel`) and the LBS shared-memory tree walk (`lbs_tile_kernel`), plus the statistics kernel.
  compute-sanitizer --tool racecheck python tools/sanitize_targets.py
  compute-sanitizer --tool memcheck  python tools/sanitize_targets.py all     # + head, encoder (split), blend GEMM
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch
import hierarchicalprobabilistic3dhuman_b200 as hp
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn

everything = len(sys.argv) > 1 and sys.argv[1] == "all"
dev = torch.device("cuda", 0)
rs = np.random.RandomState(0)
B, N = 3, 40
t = lambda a: torch.from_numpy(a.astype(np.float32)).to(dev)
U = t(np.linalg.qr(rs.normal(size=(B, 23, 3, 3)))[0]); V = t(np.linalg.qr(rs.normal(size=(B, 23, 3, 3)))[0])
S = t(np.sort(np.exp(rs.uniform(-3, 4, size=(B, 23, 3))), axis=-1)[..., ::-1].copy())
R = hp.pose_matrix_fisher_sampling_torch(U, S, V, N)                                   # Philox mode
eps, w = torch.randn(B, 23, 8 * N, 4, device=dev), torch.rand(B, 23, 8 * N, device=dev)
R2 = hp.pose_matrix_fisher_sampling_torch(U, S, V, N, noise=(eps, w))                  # injected-noise mode
smpl = hp.SMPL(model=syn.synthetic_smpl_model()).to(dev)
if not everything:
    os.environ["HP3D_BLEND"] = "fp32"            # racecheck run: keep the tcgen05/TMA kernels out (tool support), LBS is the target
gR = hp.rot6d_to_rotmat(torch.randn(B, 6, device=dev))
out = smpl(body_pose=R.view(B * N, 23, 3, 3), global_orient=gR.unsqueeze(1), betas=torch.randn(B, 10, device=dev), pose2rot=False)
mean, unc = hp.vertex_uncertainty(out.vertices.view(B, N, 6890, 3))
if everything:
    from test_gpu_net import make_model
    net = make_model("split")
    x = torch.from_numpy(syn.synthetic_proxy_rep(2, seed=0)).to(dev)
    F, U2, S2, V2, mode, dist, glob, cam = net(x)
torch.cuda.synchronize()
hp.check_sampler_status()
print("sanitize targets done:", float(out.vertices.abs().max()), float(unc.mean()))
