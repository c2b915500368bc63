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
from oracle import reference_import, net_oracle, sampler_oracle, proxy_oracle, crop_oracle, mf_loss_oracle   # noqa: E402
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
    # ---- sample-ranking helpers: heat-map arg-max and projection, the reference's own functions
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        from utils.label_conversions import (convert_heatmaps_to_2Djoints_coordinates_torch,
                                             convert_2Djoints_to_gaussian_heatmaps_torch)
        from utils.cam_utils import orthographic_project_torch
        from utils.joints2d_utils import undo_keypoint_normalisation
        from models.canny_edge_detector import CannyEdgeDetector
    xr = torch.from_numpy(syn.synthetic_proxy_rep(3, seed=5))
    j2d_r, vis_r = convert_heatmaps_to_2Djoints_coordinates_torch(xr[:, 1:], eps=1e-6)
    j2d_o, vis_o = sampler_oracle.heatmaps_to_joints2d(xr[:, 1:])
    assert torch.equal(j2d_r, j2d_o) and torch.equal(vis_r, vis_o)
    rs = np.random.RandomState(3)
    Jc = torch.from_numpy(rs.normal(0, 0.4, size=(5, 17, 3)).astype(np.float32))
    camr = torch.tensor([[0.87, 0.05, -0.1]])
    flipped = Jc * torch.tensor([1.0, -1.0, -1.0])      # pytorch3d 180-degree flip about x is absent here: restated
    px_r = undo_keypoint_normalisation(orthographic_project_torch(flipped, camr), 256)
    px_o = sampler_oracle.project_joints_to_pixels(Jc, camr, 256)
    assert torch.equal(px_r, px_o)
    np.savez_compressed(os.path.join(GOLD, "rank_helpers.npz"), joints2d=j2d_r.numpy(), vis=vis_r.numpy(), J=Jc.numpy(),
                        cam=camr.numpy(), pixels=px_r.numpy(), proxy_seed=5)

    # ---- proxy-representation generation: the reference's CannyEdgeDetector and heat-maps
    rgb, j2d, vis = (torch.from_numpy(a) for a in syn.synthetic_images(2, seed=5))
    store = {"image_seed": 5, "rgb_checksum": checksum(rgb), "joints2d": j2d.numpy(), "vis": vis.numpy()}
    for tag, thr, nms in (("cfg", 0.0, True), ("thr", 0.2, True), ("nonms", 0.1, False)):   # cfg = configs/...:21-24
        det = CannyEdgeDetector(non_max_suppression=nms, gaussian_filter_std=1.0, gaussian_filter_size=5, threshold=thr)
        with torch.no_grad():
            r_ = det(rgb)
        o_ = proxy_oracle.canny_edges(rgb, thr, nms)
        for k in r_:
            assert torch.equal(r_[k], o_[k]), (tag, k)             # restatement is bit-identical to the reference
        key = "thresholded_thin_edges" if nms else "thresholded_grad_magnitude"
        store[f"edges_{tag}"] = r_[key].numpy()
        if tag == "cfg":
            store["grad_orientation"] = r_["grad_orientation"].numpy().astype(np.uint16)   # multiples of 45 <= 360
            store["blurred_checksum"] = checksum(r_["blurred_img"])
            store["grad_magnitude_checksum"] = checksum(r_["grad_magnitude"])
    heat_r = convert_2Djoints_to_gaussian_heatmaps_torch(j2d, 256, std=4)
    assert torch.equal(heat_r, proxy_oracle.joints2d_to_heatmaps(j2d, 256, 4))
    heat_m = heat_r * vis[:, :, None, None]
    store["heat_checksum"] = checksum(heat_m)
    store["heat_rows"] = heat_m[:, :, ::37, :].numpy()           # every 37th row of every map (7 rows): small, exact
    jj, vv = convert_heatmaps_to_2Djoints_coordinates_torch(heat_m, eps=1e-6)
    store["heat_argmax"], store["heat_argmax_vis"] = jj.numpy(), vv.numpy()
    # the full network on this image-space input (reference predict/...:91-104)
    proxy = torch.cat([torch.from_numpy(store["edges_cfg"]), heat_m], dim=1).float()
    assert torch.equal(proxy, proxy_oracle.proxy_representation(rgb, j2d, vis))
    with torch.no_grad():
        store["feats"] = model.image_encoder(proxy).numpy()
    np.savez_compressed(os.path.join(GOLD, "proxy_b2.npz"), **store)
    # ---- crop / affine resample and HRNet key-point arg-max: the reference's own functions
    with contextlib.redirect_stdout(io.StringIO()):
        from utils.image_utils import batch_crop_pytorch_affine
        from predict.predict_hrnet import get_kp_locations_confs_from_heatmaps
    crgb, cj2d, cc, ch, cw = (torch.from_numpy(a) for a in syn.synthetic_crop_inputs(3, seed=8))
    store = {"crop_seed": 8, "rgb_checksum": checksum(crgb)}
    for tag, scale in (("s10", 1.0), ("s12", 1.2)):            # predict/...:93 uses 1.0, the function's default is 1.2
        r_ = batch_crop_pytorch_affine(input_wh=(288, 384), output_wh=(256, 256), num_to_crop=3, device="cpu", joints2D=cj2d,
                                       rgb=crgb, bbox_centres=cc.clone(), bbox_heights=ch.clone(), bbox_widths=cw.clone(),
                                       orig_scale_factor=scale)
        o_ = crop_oracle.batch_crop_affine((288, 384), (256, 256), cj2d, crgb, cc, ch, cw, scale)
        assert torch.equal(r_["joints2D"], o_["joints2D"]) and torch.equal(r_["rgb"], o_["rgb"]), tag
        store[f"joints2D_{tag}"] = r_["joints2D"].numpy()
        store[f"rgb_rows_{tag}"] = r_["rgb"][:, :, ::31, :].numpy()          # every 31st row (9 rows), exact
        store[f"rgb_checksum_{tag}"] = checksum(r_["rgb"])
    rs = np.random.RandomState(4)
    hm = torch.from_numpy(rs.normal(size=(2, 17, 96, 72)).astype(np.float32))
    hm[0, 3] = -1.0                                                          # never positive: key point zeroed
    kps, confs = get_kp_locations_confs_from_heatmaps(hm)
    k2, c2 = crop_oracle.keypoints_from_heatmaps(hm)
    assert torch.equal(kps, k2) and torch.equal(confs, c2)
    store["hrnet_kps"], store["hrnet_confs"] = kps.numpy(), confs.numpy()
    np.savez_compressed(os.path.join(GOLD, "crop_b3.npz"), **store)
    # ---- matrix-Fisher normalising constant and its gradient: the reference's LogMFNormConstant
    with contextlib.redirect_stdout(io.StringIO()):
        from losses.matrix_fisher_loss import LogMFNormConstant
    rs = np.random.RandomState(6)
    Sn = np.exp(rs.uniform(np.log(1e-2), np.log(5e2), size=(600, 3))).astype(np.float32)
    Sn = -np.sort(-Sn, axis=1)
    Sn[::3, 2] *= -1.0                                                       # proper s3 is negative when det(U V^T) = -1
    St = torch.from_numpy(Sn).requires_grad_(True)
    lc = LogMFNormConstant.apply(St)
    lc.sum().backward()
    lo, go = mf_loss_oracle.log_mf_norm_constant(torch.from_numpy(Sn))
    assert torch.equal(lo, lc.detach()) and torch.equal(go, St.grad)
    np.savez_compressed(os.path.join(GOLD, "mf_norm.npz"), S=Sn, log_c=lc.detach().numpy(), dlogc_ds=St.grad.numpy())
    print("golden fixtures written to", os.path.normpath(GOLD))
    for f in sorted(os.listdir(GOLD)):
        print(" ", f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
