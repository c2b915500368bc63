"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
build container only) on seeded synthetic inputs, and pin the oracle restatements against it.

  python -m oracle.make_golden

Inputs are NOT stored (they are regenerated bit-identically from seeds by
hierarchicalprobabilistic3dhuman_b200.synthetic); fp64 checksums of inputs/weights are stored so a
drifting generator is detected. Outputs stored are a few hundred KB.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import reference_import, net_oracle, sampler_oracle   # noqa: E402
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn   # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def checksum(t):
    t = torch.as_tensor(t).double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def sd_checksum(sd):
    return np.sum([checksum(v.float())[1] for k, v in sorted(sd.items())])


def main():
    torch.set_num_threads(os.cpu_count())
    ref = reference_import.import_reference()
    os.makedirs(GOLD, exist_ok=True)
    parents = syn.SMPL_PARENTS.tolist()

    # ---- config[0]: 4 synthetic 256x256 inputs through the reference network
    sd = syn.synthetic_state_dict(0)
    model = ref.PoseMFShapeGaussianNet(parents, ref.config).eval()
    model.load_state_dict(sd)
    x = torch.from_numpy(syn.synthetic_proxy_rep(4, seed=0))
    with torch.no_grad():
        feats = model.image_encoder(x)
        F, U, S, V, mode, dist, glob, cam = model(x)
        glob_R = ref.rot6d_to_rotmat(glob)
    # oracle restatement must reproduce the reference bit for bit
    with torch.no_grad():
        f2 = net_oracle.encoder_forward(sd, x)
        h2 = net_oracle.head_forward(sd, f2, parents)
    assert torch.equal(feats, f2) and torch.equal(F, h2["F"]) and torch.equal(U, h2["U"]) and torch.equal(mode, h2["mode"])
    assert torch.equal(glob_R, net_oracle.rot6d_to_rotmat(glob))
    np.savez_compressed(os.path.join(GOLD, "net_b4.npz"), feats=feats.numpy(), F=F.numpy(), U=U.numpy(), S=S.numpy(),
                        V=V.numpy(), mode=mode.numpy(), shape_loc=dist.loc.numpy(), shape_scale=dist.scale.numpy(),
                        glob=glob.numpy(), cam=cam.numpy(), glob_rotmats=glob_R.numpy(),
                        x_checksum=checksum(x), sd_checksum=sd_checksum(sd), weights_seed=0, input_seed=0)

    # ---- head alone at B=64 on synthetic features (input_feats bypass, reference :90-91)
    rs = np.random.RandomState(7)
    feats64 = torch.from_numpy(np.abs(rs.normal(0, 1.0, size=(64, 512))).astype(np.float32))
    with torch.no_grad():
        r = model(None, input_feats=feats64)
    np.savez_compressed(os.path.join(GOLD, "head_b64.npz"), F=r[0].numpy(), U=r[1].numpy(), S=r[2].numpy(),
                        V=r[3].numpy(), mode=r[4].numpy(), shape_loc=r[5].loc.numpy(), shape_scale=r[5].scale.numpy(),
                        glob=r[6].numpy(), cam=r[7].numpy(), feats_seed=7)

    # ---- sampler: reference draws from torch's CPU generator; replayable from the seed
    for name, (Un, Sn, Vn), N, seed in [
            ("sampler_usv_b4_n8", syn.synthetic_usv(4, seed=1), 8, 123),
            ("sampler_head_b4_n8", (U.numpy(), S.numpy(), V.numpy()), 8, 321),
            ("sampler_lowk_b2_n100", syn.synthetic_usv(2, seed=3, s_lo=1e-2, s_hi=1.0), 100, 11),
            ("sampler_highk_b2_n100", syn.synthetic_usv(2, seed=4, s_lo=50.0, s_hi=500.0), 100, 12)]:
        Ut, St, Vt = torch.from_numpy(Un), torch.from_numpy(Sn), torch.from_numpy(Vn)
        torch.manual_seed(seed)
        R = ref.pose_matrix_fisher_sampling_torch(Ut, St, Vt, N)
        torch.manual_seed(seed)
        eps, w = sampler_oracle.draw_noise(Ut.shape[0], Ut.shape[1], N)
        R2, acc = sampler_oracle.sample_with_noise(Ut, St, Vt, N, eps, w)
        assert torch.equal(R, R2), name
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), R=R.numpy(), U=Un, S=Sn, V=Vn, N=N, seed=seed,
                            accepted=acc.numpy(), noise_checksum=checksum(eps) + checksum(w))
    print("golden fixtures written to", os.path.normpath(GOLD))
    for f in sorted(os.listdir(GOLD)):
        print(" ", f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
