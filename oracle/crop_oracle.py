"""CPU restatement of the crop / resample step in front of the proxy-representation generator
-- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* `batch_crop_affine`: reference utils/image_utils.py:234-378 (`batch_crop_pytorch_affine`) for the call the predict
  path makes (predict/predict_poseMF_shapeGaussian_net.py:84-93: given bbox centre / height / width, RGB + 2D joints,
  no augmentation): bounding box → aspect-ratio fix → scale → forward affine (joints) and normalised inverse affine →
  `F.affine_grid` + bilinear `F.grid_sample` (align_corners=False, zero padding), restated with explicit index
  arithmetic.
* `keypoints_from_heatmaps`: reference predict/predict_hrnet.py:7-30 (`get_kp_locations_confs_from_heatmaps`).

PINNED against the imported reference by `oracle/make_golden.py` (tests/golden/crop_b3.npz): joints, affine matrices and
resampled pixels are BIT-IDENTICAL to the reference as run in the build container (ATen's vectorised grid_sample forms
the source coordinate with a fused multiply-subtract and accumulates the four corners with fused multiply-adds; the
restatement reproduces that order).
"""
import torch


def _fma(a, b, c):
    """fused multiply-add emulated in float64 (the 24x24-bit product is exact there)"""
    return (a.double() * b.double() + c.double()).float()


def crop_affine_matrices(input_wh, output_wh, bbox_centres, bbox_heights, bbox_widths, orig_scale_factor=1.2):
    """-> (affine (B,2,3) forward pixel transform, theta (B,2,3) normalised inverse for affine_grid), both float32.
    bbox_centres (B,2) as (vertical, horizontal); image_utils.py:305-349 without the random augmentations."""
    in_wh = torch.tensor(input_wh, dtype=torch.float32)
    out_wh = torch.tensor(output_wh, dtype=torch.float32)
    c = bbox_centres.clone().float()
    h = bbox_heights.clone().float()
    w = bbox_widths.clone().float()
    aspect = (out_wh[1] / out_wh[0]).item()
    m = h > w * aspect
    w[m] = h[m] / aspect
    m = h < w * aspect
    h[m] = w[m] * aspect
    h = h * orig_scale_factor
    w = w * orig_scale_factor
    B = c.shape[0]
    affine = torch.zeros(B, 2, 3)
    affine[:, 0, 0] = out_wh[0] / w
    affine[:, 1, 1] = out_wh[1] / h
    whs = torch.stack([w, h], dim=-1)
    affine[:, :, 2] = out_wh * 0.5 - (out_wh / whs) * c[:, [1, 0]]
    theta = torch.zeros(B, 2, 3)
    theta[:, 0, 0] = w / in_wh[0]
    theta[:, 1, 1] = h / in_wh[1]
    theta[:, :, 2] = -affine[:, :, 2] / (out_wh / whs)
    theta[:, :, 2] = theta[:, :, 2] / (in_wh * 0.5) + (whs / in_wh) - 1
    return affine, theta


def affine_grid_sample_bilinear(img, theta, out_h, out_w):
    """F.grid_sample(img, F.affine_grid(theta, [B,1,out_h,out_w], align_corners=False), 'bilinear', 'zeros', False)
    with explicit arithmetic. img (B,C,H,W) float32 -> (B,C,out_h,out_w)."""
    B, C, H, W = img.shape
    # affine_grid, align_corners=False (ATen AffineGridGenerator): base grid = linspace(-1, 1, n) * (n - 1) / n per axis with
    # a homogeneous 1, multiplied by theta^T as a batched matrix product
    base = torch.empty(B, out_h, out_w, 3, dtype=torch.float32)
    base[..., 0] = torch.linspace(-1, 1, out_w) * (out_w - 1) / out_w
    base[..., 1] = (torch.linspace(-1, 1, out_h) * (out_h - 1) / out_h)[:, None]
    base[..., 2] = 1
    grid = base.view(B, out_h * out_w, 3).bmm(theta.transpose(1, 2)).view(B, out_h, out_w, 2)
    gx, gy = grid[..., 0], grid[..., 1]
    # grid_sample unnormalisation, align_corners=False, then the bilinear corner weights as ATen's GridSamplerKernel forms
    # them: w = x - floor(x), e = 1 - w, n = y - floor(y), s = 1 - n
    ix = _fma(gx + 1, torch.tensor(W / 2.0), torch.tensor(-0.5))    # vectorised kernel: fmsub((x + 1), size / 2, 0.5)
    iy = _fma(gy + 1, torch.tensor(H / 2.0), torch.tensor(-0.5))
    x0 = torch.floor(ix); y0 = torch.floor(iy)
    x1 = x0 + 1; y1 = y0 + 1
    w_ = ix - x0; e_ = 1 - w_; n_ = iy - y0; s_ = 1 - n_
    w_nw = s_ * e_; w_ne = s_ * w_; w_sw = n_ * e_; w_se = n_ * w_

    def tap(xi, yi):
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        xc = xi.clamp(0, W - 1).long(); yc = yi.clamp(0, H - 1).long()
        flat = (yc * W + xc).view(B, 1, -1).expand(B, C, -1)
        v = torch.gather(img.reshape(B, C, H * W), 2, flat).view(B, C, out_h, out_w)
        return v * ok[:, None].to(img.dtype)

    # accumulated as one multiply followed by three fused multiply-adds, in corner order nw, ne, sw, se
    acc = tap(x0, y0) * w_nw[:, None]
    acc = _fma(tap(x1, y0), w_ne[:, None], acc)
    acc = _fma(tap(x0, y1), w_sw[:, None], acc)
    return _fma(tap(x1, y1), w_se[:, None], acc)


def batch_crop_affine(input_wh, output_wh, joints2D, rgb, bbox_centres, bbox_heights, bbox_widths, orig_scale_factor=1.2):
    """-> dict(rgb (B,C,out_h,out_w), joints2D (B,K,2)) like the reference's `cropped_dict` (image_utils.py:352-378)."""
    affine, theta = crop_affine_matrices(input_wh, output_wh, bbox_centres, bbox_heights, bbox_widths, orig_scale_factor)
    out = {}
    if joints2D is not None:
        homo = torch.cat([joints2D.float(), torch.ones(joints2D.shape[0], joints2D.shape[1], 1)], dim=-1)
        out["joints2D"] = torch.einsum("bij,bkj->bki", affine, homo)
    if rgb is not None:
        out["rgb"] = affine_grid_sample_bilinear(rgb.float(), theta, int(output_wh[1]), int(output_wh[0]))
    return out


def keypoints_from_heatmaps(batch_heatmaps):
    """(B,K,h,w) -> (pred_kps (B,K,2) as (x, y) of the arg-max, zeroed where the maximum is not positive; max_confs (B,K))."""
    B, K, h, w = batch_heatmaps.shape
    confs, idx = torch.max(batch_heatmaps.reshape(B, K, -1), dim=2)
    kps = torch.zeros(B, K, 2, dtype=torch.float32)
    kps[:, :, 0] = (idx % w).float()
    kps[:, :, 1] = torch.floor(idx / float(w))
    kps = kps * (confs > 0.0)[:, :, None]
    return kps, confs
