"""BASELINE configs[4]: batch 512 x N_samples=500 matrix-Fisher sampler sweep, high-kappa vs low-kappa.
Reports accept rate (= rejection-loop lane occupancy: every proposal occupies one lane-slot, accepted ones are
useful work), proposals per accepted sample, rotations/s and kernel time."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import hierarchicalprobabilistic3dhuman_b200 as hp
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn

B, N = int(os.environ.get("B", 512)), int(os.environ.get("N", 500))
out = []
for name, lo, hi in (("low-kappa S<1", 1e-2, 1.0), ("mid-kappa 1<S<50", 1.0, 50.0), ("high-kappa S>50", 50.0, 500.0), ("full sweep", 1e-2, 5e2)):
    U, S, V = (torch.from_numpy(a).cuda() for a in syn.synthetic_usv(B, seed=7, s_lo=lo, s_hi=hi))
    R = torch.empty(B, N, 23, 3, 3, device="cuda")
    for _ in range(2):
        hp.pose_matrix_fisher_sampling_torch(U, S, V, N, out=R)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _, stats = hp.pose_matrix_fisher_sampling_torch(U, S, V, N, out=R, return_stats=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    prop, acc, fail = (int(v) for v in stats)
    out.append({"regime": name, "B": B, "N": N, "ms": ms, "rotations_per_s": B * N * 23 / (ms * 1e-3), "accept_rate": acc / prop,
                "proposals_per_sample": prop / (B * N * 23), "exhausted_chains": fail, "out_GBps": B * N * 23 * 36 / (ms * 1e-3) / 1e9})
    print(json.dumps(out[-1]))
