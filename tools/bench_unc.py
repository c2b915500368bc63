"""Micro-benchmark of the per-vertex sample-statistics kernel (HP3D_UNC selects the variant) at the bench size,
with a correctness check against torch."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import hierarchicalprobabilistic3dhuman_b200 as hp
B, N = int(os.environ.get("B", 256)), int(os.environ.get("N", 100))
v = torch.randn(B, N, 6890, 3, device="cuda")
mean, dist = hp.vertex_uncertainty(v)
ref_mean = v[:8].mean(1); ref = (v[:8] - ref_mean[:, None]).norm(dim=-1).mean(1)
err = ((dist[:8] - ref).abs().max() / ref.abs().max()).item(); errm = (mean[:8] - ref_mean).abs().max().item()
for _ in range(3): hp.vertex_uncertainty(v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): hp.vertex_uncertainty(v)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"variant": os.environ.get("HP3D_UNC", "default"), "ms": ms, "GBps": v.numel() * 4 / ms / 1e6, "rel_err": err, "mean_err": errm}))
