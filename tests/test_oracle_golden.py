"""The oracle restatements reproduce the reference's own outputs (golden fixtures written by
oracle/make_golden.py from the unmodified reference). CPU only."""
import numpy as np
import torch

from conftest import load_golden, rel_err
from oracle import net_oracle, sampler_oracle
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn


def _checksum(t):
    t = torch.as_tensor(t).double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def test_synthetic_inputs_reproduce_checksums():
    g = load_golden("net_b4")
    x = syn.synthetic_proxy_rep(4, seed=0)
    np.testing.assert_allclose(_checksum(x), g["x_checksum"], rtol=1e-12)
    sd = syn.synthetic_state_dict(0)
    s = np.sum([_checksum(v.float())[1] for k, v in sorted(sd.items())])
    np.testing.assert_allclose(s, g["sd_checksum"], rtol=1e-12)


def test_net_oracle_matches_reference_golden():
    g = load_golden("net_b4")
    sd = syn.synthetic_state_dict(0)
    x = torch.from_numpy(syn.synthetic_proxy_rep(4, seed=0))
    with torch.no_grad():
        feats = net_oracle.encoder_forward(sd, x)
        h = net_oracle.head_forward(sd, feats, syn.SMPL_PARENTS)
    # same code path / same LAPACK in the same image: bit-exact here; tolerance covers thread-count effects
    assert rel_err(feats, g["feats"]) < 1e-5
    for k in ("F", "S", "mode"):
        assert rel_err(h[k], g[k]) < 1e-4, k
    assert rel_err(h["shape_params"][:, :10], g["shape_loc"]) < 1e-5
    assert rel_err(torch.exp(h["shape_params"][:, 10:]), g["shape_scale"]) < 1e-5
    assert rel_err(h["glob"], g["glob"]) < 1e-5 and rel_err(h["cam"], g["cam"]) < 1e-5
    assert rel_err(net_oracle.rot6d_to_rotmat(h["glob"]), g["glob_rotmats"]) < 1e-5


def test_head_oracle_matches_reference_golden_b64():
    g = load_golden("head_b64")
    sd = syn.synthetic_state_dict(0)
    rs = np.random.RandomState(7)
    feats = torch.from_numpy(np.abs(rs.normal(0, 1.0, size=(64, 512))).astype(np.float32))
    with torch.no_grad():
        h = net_oracle.head_forward(sd, feats, syn.SMPL_PARENTS)
    for k in ("F", "U", "S", "V", "mode"):
        assert rel_err(h[k], g[k]) < 1e-5, k


def test_sampler_oracle_matches_reference_golden():
    for name in ("sampler_usv_b4_n8", "sampler_head_b4_n8", "sampler_lowk_b2_n100", "sampler_highk_b2_n100"):
        g = load_golden(name)
        U, S, V = (torch.from_numpy(g[k]) for k in ("U", "S", "V"))
        N = int(g["N"])
        torch.manual_seed(int(g["seed"]))
        eps, w = sampler_oracle.draw_noise(U.shape[0], U.shape[1], N)
        np.testing.assert_allclose(_checksum(eps) + _checksum(w), g["noise_checksum"], rtol=1e-10)
        R, acc = sampler_oracle.sample_with_noise(U, S, V, N, eps, w)
        assert rel_err(R, g["R"]) < 1e-6, name
        assert np.array_equal(acc.numpy(), g["accepted"])
        # rotation validity
        assert (torch.det(R) - 1).abs().max() < 1e-5
        eye = torch.eye(3)
        assert (R.transpose(-1, -2) @ R - eye).abs().max() < 1e-5


def test_sample_ranking_helpers_match_reference_golden():
    """heat-map arg-max and projection helpers vs outputs of the reference's own functions."""
    g = load_golden("rank_helpers")
    x = torch.from_numpy(syn.synthetic_proxy_rep(3, seed=int(g["proxy_seed"])))
    j2d, vis = sampler_oracle.heatmaps_to_joints2d(x[:, 1:])
    assert torch.equal(j2d, torch.from_numpy(g["joints2d"])) and torch.equal(vis, torch.from_numpy(g["vis"]))
    px = sampler_oracle.project_joints_to_pixels(torch.from_numpy(g["J"]), torch.from_numpy(g["cam"]), 256)
    assert torch.equal(px, torch.from_numpy(g["pixels"]))
    # ranking is a permutation sorted by error
    rs = np.random.RandomState(1)
    joints = torch.from_numpy(rs.normal(0, 0.4, size=(3, 11, 90, 3)).astype(np.float32))
    cam = torch.tensor([[0.9, 0.0, 0.0]]).expand(3, -1)
    order, err = sampler_oracle.rank_samples(joints, x[:, 1:], cam)
    assert all(sorted(o.tolist()) == list(range(11)) for o in order)
    assert (torch.gather(err, 1, order).diff(dim=1) >= 0).all()


def test_proxy_representation_oracle_matches_reference_golden():
    """Canny edges + joint heat-maps vs outputs of the reference's own CannyEdgeDetector /
    convert_2Djoints_to_gaussian_heatmaps_torch: the restatement is bit-identical."""
    from oracle import proxy_oracle
    g = load_golden("proxy_b2")
    rgb, j2d, vis = (torch.from_numpy(a) for a in syn.synthetic_images(2, seed=int(g["image_seed"])))
    np.testing.assert_allclose(_checksum(rgb), g["rgb_checksum"], rtol=1e-12)
    assert np.array_equal(j2d.numpy(), g["joints2d"]) and np.array_equal(vis.numpy(), g["vis"])
    for tag, thr, nms in (("cfg", 0.0, True), ("thr", 0.2, True), ("nonms", 0.1, False)):
        o = proxy_oracle.canny_edges(rgb, thr, nms)
        key = "thresholded_thin_edges" if nms else "thresholded_grad_magnitude"
        assert np.array_equal(o[key].numpy(), g[f"edges_{tag}"]), tag
        if tag == "cfg":
            assert np.array_equal(o["grad_orientation"].numpy().astype(np.uint16), g["grad_orientation"])
            np.testing.assert_allclose(_checksum(o["blurred_img"]), g["blurred_checksum"], rtol=1e-12)
            np.testing.assert_allclose(_checksum(o["grad_magnitude"]), g["grad_magnitude_checksum"], rtol=1e-12)
    heat = proxy_oracle.joints2d_to_heatmaps(j2d, 256, 4, vis)
    assert np.array_equal(heat[:, :, ::37, :].numpy(), g["heat_rows"])
    np.testing.assert_allclose(_checksum(heat), g["heat_checksum"], rtol=1e-12)
    jj, vv = sampler_oracle.heatmaps_to_joints2d(heat)
    assert np.array_equal(jj.numpy(), g["heat_argmax"]) and np.array_equal(vv.numpy(), g["heat_argmax_vis"])
    # edge maps are sparse, thin and non-negative; border pixels carry the zero-padding gradient
    e = torch.from_numpy(g["edges_cfg"])
    assert (e >= 0).all() and 0.05 < (e > 0).float().mean() < 0.5


def test_proxy_oracle_small_known_answers():
    from oracle import proxy_oracle
    # a constant image has zero gradient away from the (zero-padded) border: no interior edges
    e = proxy_oracle.canny_edges(torch.full((1, 3, 32, 32), 0.5), 0.0, True)
    assert e["thresholded_thin_edges"][0, 0, 6:-6, 6:-6].abs().max() == 0
    # a vertical step edge survives thinning as a 1-pixel-wide line with orientation 0/180 degrees
    img = torch.zeros(1, 1, 32, 32); img[..., 16:] = 1.0
    e = proxy_oracle.canny_edges(img, 0.0, True)
    row = e["thresholded_thin_edges"][0, 0, 16]
    assert (row[8:24] > 0).sum() == 1
    assert set(e["grad_orientation"][0, 0, 16, 14:18].tolist()) <= {0.0, 180.0, 360.0}
    # heat-map: peak 1 at an integer joint, (u, v) = (column, row)
    h = proxy_oracle.joints2d_to_heatmaps(torch.tensor([[[5.0, 9.0]]]), 16, 4)
    assert h[0, 0, 9, 5] == 1.0 and h[0, 0].argmax().item() == 9 * 16 + 5


def test_crop_oracle_matches_reference_golden():
    """Crop / affine resample and HRNet key-point arg-max (groundwork: oracle only) vs outputs of
    the reference's own batch_crop_pytorch_affine / get_kp_locations_confs_from_heatmaps: bit-identical."""
    from oracle import crop_oracle
    g = load_golden("crop_b3")
    rgb, j2d, c, h, w = (torch.from_numpy(a) for a in syn.synthetic_crop_inputs(3, seed=int(g["crop_seed"])))
    np.testing.assert_allclose(_checksum(rgb), g["rgb_checksum"], rtol=1e-12)
    for tag, scale in (("s10", 1.0), ("s12", 1.2)):
        o = crop_oracle.batch_crop_affine((288, 384), (256, 256), j2d, rgb, c, h, w, scale)
        assert np.array_equal(o["joints2D"].numpy(), g[f"joints2D_{tag}"]), tag
        assert np.array_equal(o["rgb"][:, :, ::31, :].numpy(), g[f"rgb_rows_{tag}"]), tag
        np.testing.assert_allclose(_checksum(o["rgb"]), g[f"rgb_checksum_{tag}"], rtol=1e-12)
    # the predict path's call (whole image as the box, scale 1.0) is a pure resize: joints scale by 256/384 about the centre
    o = crop_oracle.batch_crop_affine((288, 384), (256, 256), j2d[:1], None, c[:1], h[:1], w[:1], 1.0)
    exp = (j2d[:1] - torch.tensor([144.0, 192.0])) * (256.0 / 384.0) + 128.0
    assert (o["joints2D"] - exp).abs().max() < 1e-4
    rs = np.random.RandomState(4)
    hm = torch.from_numpy(rs.normal(size=(2, 17, 96, 72)).astype(np.float32))
    hm[0, 3] = -1.0
    kps, confs = crop_oracle.keypoints_from_heatmaps(hm)
    assert np.array_equal(kps.numpy(), g["hrnet_kps"]) and np.array_equal(confs.numpy(), g["hrnet_confs"])
    assert kps[0, 3].tolist() == [0.0, 0.0] and confs[0, 3] == -1.0
