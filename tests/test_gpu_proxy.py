"""GPU parity of the proxy-representation generation against the oracle and the golden
outputs of the reference's own CannyEdgeDetector / heat-map functions (tests/golden/proxy_b2.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, reference_config

pytestmark = pytest.mark.gpu


def _inputs(n, seed):
    from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
    return tuple(torch.from_numpy(a) for a in syn.synthetic_images(n, seed=seed))


def _mismatch(a, b, tol=1e-6):
    """fraction of elements differing by more than tol * max|b| (non-max suppression is discontinuous: an ulp in
    atan2f can flip a pixel between 'edge' and 0, so edge maps are judged by mismatch fraction + value parity)."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs() > tol * b.abs().max()).double().mean().item()


@pytest.mark.parametrize("thr,nms,tag", [(0.0, True, "cfg"), (0.2, True, "thr"), (0.1, False, "nonms")])
def test_canny_matches_reference_golden(built_lib, thr, nms, tag):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from oracle import proxy_oracle
    g = load_golden("proxy_b2")
    rgb, j2d, vis = _inputs(2, int(g["image_seed"]))
    det = hp.CannyEdgeDetector(non_max_suppression=nms, gaussian_filter_std=1.0, gaussian_filter_size=5, threshold=thr)
    out = det(rgb.cuda())
    ref = proxy_oracle.canny_edges(rgb, thr, nms)
    assert set(out) == set(ref)
    # linear stages: same fp32 operation order as the reference's oneDNN convolutions -> exact
    assert torch.equal(out["blurred_img"].cpu(), ref["blurred_img"])
    # magnitude: the oracle's `(gx**2 + gy**2) ** 0.5` runs on the GPU box's HOST cpu through whatever vector pow that cpu
    # dispatches to (AVX-512 sqrt: 0.6 % of results 1 ulp off IEEE; on one box, r02m, 3e-4 absolute off on near-zero
    # magnitudes). The strict check is therefore made against the magnitude formed in float64 from the oracle's own Sobel
    # responses (host-independent); the oracle's fp32 magnitude only bounds the result loosely.
    refg = proxy_oracle.canny_edges(rgb, thr, nms, return_gradients=True)
    mag64 = torch.sqrt(refg["grad_x"].double() ** 2 + refg["grad_y"].double() ** 2).float()
    assert rel_err(out["grad_magnitude"], mag64) < 1e-6
    assert rel_err(out["grad_magnitude"], ref["grad_magnitude"]) < 1e-3
    thr64 = torch.where(mag64 < thr, torch.zeros_like(mag64), mag64)
    assert _mismatch(out["thresholded_grad_magnitude"], thr64) < 1e-5
    # orientation bins / thinning: identical except where atan2f differs by an ulp at a bin boundary
    assert _mismatch(out["grad_orientation"], ref["grad_orientation"]) < 1e-4
    key = "thresholded_thin_edges" if nms else "thresholded_grad_magnitude"
    assert _mismatch(out[key], g[f"edges_{tag}"]) < 1e-4
    assert rel_err(out["blurred_img"], ref["blurred_img"]) == 0


def test_canny_odd_sizes_and_filter_sizes(built_lib):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from oracle import proxy_oracle
    rs = np.random.RandomState(3)
    for (B, C, H, W), size, std in (((3, 1, 45, 70), 3, 0.8), ((1, 4, 33, 31), 7, 1.5), ((2, 3, 64, 64), 9, 2.0)):
        img = torch.from_numpy(rs.uniform(0, 1, size=(B, C, H, W)).astype(np.float32))
        out = hp.CannyEdgeDetector(True, std, size, 0.05)(img.cuda())
        ref = proxy_oracle.canny_edges(img, 0.05, True, std, size)
        refg = proxy_oracle.canny_edges(img, 0.05, True, std, size, return_gradients=True)
        mag64 = torch.sqrt(refg["grad_x"].double() ** 2 + refg["grad_y"].double() ** 2).float()
        assert rel_err(out["grad_magnitude"], mag64) < 1e-6 and rel_err(out["grad_magnitude"], ref["grad_magnitude"]) < 1e-3
        assert _mismatch(out["thresholded_thin_edges"], ref["thresholded_thin_edges"]) < 2e-3   # tiny images: 1 pixel ~ 5e-4


def test_heatmaps_and_proxy_rep(built_lib):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from oracle import proxy_oracle
    g = load_golden("proxy_b2")
    rgb, j2d, vis = _inputs(2, int(g["image_seed"]))
    h = hp.convert_2Djoints_to_gaussian_heatmaps_torch(j2d.cuda(), 256, std=4)
    assert rel_err(h, proxy_oracle.joints2d_to_heatmaps(j2d, 256, 4)) < 1e-6
    hm = hp.convert_2Djoints_to_gaussian_heatmaps_torch(j2d.cuda(), 256, std=4, visibility=vis.cuda())
    assert rel_err(hm[:, :, ::37, :], g["heat_rows"]) < 1e-6
    h2 = hp.convert_2Djoints_to_gaussian_heatmaps_torch(j2d[:, :5].cuda(), 48, std=2.5)
    assert rel_err(h2, proxy_oracle.joints2d_to_heatmaps(j2d[:, :5], 48, 2.5)) < 1e-6
    x = hp.proxy_representation(rgb.cuda(), j2d.cuda(), vis.cuda())
    ref = proxy_oracle.proxy_representation(rgb, j2d, vis)
    assert x.shape == (2, 18, 256, 256)
    assert rel_err(x[:, 1:], ref[:, 1:]) < 1e-6
    assert _mismatch(x[:, :1], ref[:, :1]) < 1e-4
    assert (x[:, 0] >= 0).all()


def test_joints2d_argmax_without_heatmaps(built_lib):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from oracle import proxy_oracle, sampler_oracle
    g = load_golden("proxy_b2")
    _, j2d, vis = _inputs(2, int(g["image_seed"]))
    px, v = hp.joints2d_heatmap_argmax(j2d.cuda(), vis.cuda(), 256, 4.0)
    assert np.array_equal(px.cpu().numpy(), g["heat_argmax"]) and np.array_equal(v.cpu().numpy() != 0, g["heat_argmax_vis"])
    # stress: half-integer ties, joints outside the image, far outside (invisible)
    rs = np.random.RandomState(9)
    j = rs.uniform(-30, 286, size=(64, 17, 2)).astype(np.float32)
    j[:8] = np.round(j[:8]) + 0.5
    j[8:16] = np.round(j[8:16])
    jt = torch.from_numpy(j)
    ref_px, ref_v = sampler_oracle.heatmaps_to_joints2d(proxy_oracle.joints2d_to_heatmaps(jt, 256, 4))
    px, v = hp.joints2d_heatmap_argmax(jt.cuda(), None, 256, 4.0)
    v = v.cpu() != 0
    assert (v == ref_v).float().mean() > 0.995            # max ~ eps only for joints ~21 px outside the image
    both = v & ref_v
    assert ((px.cpu() - ref_px).abs().amax(-1)[both] > 0).float().mean() < 0.01   # exp-ulp ties only


def test_fused_image_encoder_equals_proxy_then_encode(built_lib):
    """hp3d_encoder_forward_image (proxy representation written straight into the stem's fp16 NHWC records) must give
    the features of proxy_representation -> encode, and match the reference network on the same image input."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
    g = load_golden("proxy_b2")
    rgb, j2d, vis = _inputs(2, int(g["image_seed"]))
    sd = syn.synthetic_state_dict(0)
    for mode, tol in (("fast", 5e-3), ("parity", 1e-4)):
        net = hp.PoseMFShapeGaussianNet(syn.SMPL_PARENTS.tolist(), reference_config(), encoder_mode=mode)
        net.load_state_dict(sd)
        net = net.cuda().eval()
        f_img = net.encode_image(rgb.cuda(), j2d.cuda(), vis.cuda())
        f_two = net.encode(hp.proxy_representation(rgb.cuda(), j2d.cuda(), vis.cuda()))
        assert rel_err(f_img, f_two) < 1e-6, mode
        assert rel_err(f_img, g["feats"]) < tol, mode


def test_pipeline_from_images_matches_proxy_path(built_lib):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
    B, N = 4, 6
    rgb, j2d, vis = _inputs(B, 11)
    net = hp.PoseMFShapeGaussianNet(syn.SMPL_PARENTS.tolist(), reference_config(), encoder_mode="fast")
    net.load_state_dict(syn.synthetic_state_dict(0))
    net = net.cuda().eval()
    smpl = hp.SMPL(model=syn.synthetic_smpl_model()).cuda()
    pipe = hp.HotPathPipeline(net, smpl, B, N, "cuda:0")
    torch.manual_seed(5)
    a = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in pipe.run_device_images(rgb.cuda(), j2d.cuda(), vis.cuda()).items()}
    torch.manual_seed(5)
    x = hp.proxy_representation(rgb.cuda(), j2d.cuda(), vis.cuda())
    b = pipe.run_device(x)
    for k in ("mode_vertices", "rotmats", "uncertainty", "sample_error"):
        assert rel_err(a[k], b[k]) < 1e-5, k
    assert torch.equal(a["sample_order"], b["sample_order"])
    # host-streaming variant
    out, ev = pipe.run_host_images(rgb.pin_memory(), j2d.pin_memory(), vis.to(torch.uint8).pin_memory())
    ev.synchronize()
    assert rel_err(out["mode_vertices"], b["mode_vertices"]) < 1e-5


@pytest.mark.parametrize("mode", ["fast", "parity"])
def test_encoder_argmax_byproduct(built_lib, mode):
    """encode(..., return_joints2d=True): heat-map arg-max fused into the input pass == the reference's
    convert_heatmaps_to_2Djoints_coordinates_torch (pinned oracle), features unchanged."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
    from oracle import sampler_oracle
    x = torch.from_numpy(syn.synthetic_proxy_rep(5, seed=5))
    x[1, 3] = 0.0                                   # an all-zero heat-map: invisible joint
    x[2, 4] = 1e-7                                  # below eps everywhere: invisible too
    x[3, 5] = 0.0
    x[3, 5, 10:12, 20:22] = 0.75                    # a 4-way tie: first index wins
    net = hp.PoseMFShapeGaussianNet(syn.SMPL_PARENTS.tolist(), reference_config(), encoder_mode=mode)
    net.load_state_dict(syn.synthetic_state_dict(0))
    net = net.cuda().eval()
    feats, j2d, vis = net.encode(x.cuda(), return_joints2d=True)
    ref_j2d, ref_vis = sampler_oracle.heatmaps_to_joints2d(x[:, 1:])
    assert torch.equal(j2d.cpu(), ref_j2d) and torch.equal(vis.cpu() != 0, ref_vis)
    assert not ref_vis[1, 2] and not ref_vis[2, 3] and ref_j2d[3, 4].tolist() == [20.0, 10.0]
    assert torch.equal(feats, net.encode(x.cuda()))
