"""CPU restatement of the proxy-representation generation that feeds the hot path
-- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* `canny_edges`  : reference models/canny_edge_detector.py:104-166 (CannyEdgeDetector.forward): separable Gaussian
  blur per channel (zero padding at every stage), Sobel gradients summed over channels and averaged, magnitude,
  orientation binned to 45 degrees, threshold, directional non-maximum suppression.
* `joints2d_to_heatmaps` : reference utils/label_conversions.py:105-124 (+ the visibility mask of
  predict/predict_poseMF_shapeGaussian_net.py:97-99).
* `proxy_representation` : the concatenation predict/...:91-100 feeds to the network.

The reference's nn.Conv2d calls run in oneDNN, whose fp32 kernels accumulate the filter taps in row-major tap order
with fused multiply-adds starting from zero; `_conv_taps` reproduces exactly that (an FMA is emulated in float64:
the 24x24-bit product is exact there), which makes every intermediate BIT-IDENTICAL to the reference on the
fixtures (`oracle/make_golden.py` asserts it). PINNED against the imported reference.
"""
import math

import numpy as np
import torch


def gaussian_taps(size=5, std=1.0):
    """scipy.signal.windows.gaussian(size, std) normalised to sum 1, as float32 (canny_edge_detector.py:23-24,31)."""
    n = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-0.5 * (n / std) ** 2)
    return torch.from_numpy((g / g.sum())).float()


def _shift(x, dy, dx):
    """y[..., i, j] = x[..., i + dy, j + dx], zero outside the image (Conv2d zero padding)."""
    H, W = x.shape[-2:]
    y = torch.zeros_like(x)
    yd, xd = slice(max(0, -dy), min(H, H - dy)), slice(max(0, -dx), min(W, W - dx))
    ys, xs = slice(max(0, dy), min(H, H + dy)), slice(max(0, dx), min(W, W + dx))
    y[..., yd, xd] = x[..., ys, xs]
    return y


def _fma(a, b, c):
    return (a.double() * b.double() + c.double()).float()


def _conv_taps(x, taps):
    """Cross-correlation of x (..., H, W) with `taps` = [(weight, dy, dx)] in row-major filter order; zero-weight taps
    contribute fma(0, v, acc) = acc and are skipped."""
    acc = torch.zeros_like(x)
    for w, dy, dx in taps:
        if w != 0.0:
            acc = _fma(torch.tensor(float(w), dtype=torch.float32), _shift(x, dy, dx), acc)
    return acc


SOBEL = ((1, 0, -1), (2, 0, -2), (1, 0, -1))                         # canny_edge_detector.py:41-43
# directional filters: index -> (dy, dx) of the neighbour subtracted from the centre (canny_edge_detector.py:62-100)
NMS_NEIGHBOUR = ((0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1))


def canny_edges(img, threshold=0.0, non_max_suppression=True, gaussian_filter_std=1.0, gaussian_filter_size=5,
                return_gradients=False):
    """img (B,C,H,W) float32 -> dict with the reference's keys (each (B,1,H,W) except blurred_img (B,C,H,W)).
    return_gradients=True adds the channel-averaged Sobel responses "grad_x", "grad_y" (B,1,H,W) (not reference outputs: test
    helpers, so a checker can form the magnitude in float64 instead of relying on the host's vector pow)."""
    img = img.float()
    B, C, H, W = img.shape
    g = gaussian_taps(gaussian_filter_size, gaussian_filter_std)
    r = gaussian_filter_size // 2
    gx = torch.zeros(B, H, W)
    gy = torch.zeros(B, H, W)
    blurred_img = torch.zeros_like(img)
    for c in range(C):
        bh = _conv_taps(img[:, c], [(g[k].item(), 0, k - r) for k in range(gaussian_filter_size)])
        bv = _conv_taps(bh, [(g[k].item(), k - r, 0) for k in range(gaussian_filter_size)])
        blurred_img[:, c] = bv
        gx = gx + _conv_taps(bv, [(SOBEL[i][j], i - 1, j - 1) for i in range(3) for j in range(3)])
        gy = gy + _conv_taps(bv, [(SOBEL[j][i], i - 1, j - 1) for i in range(3) for j in range(3)])
    gx, gy = gx / C, gy / C
    mag = (gx ** 2 + gy ** 2) ** 0.5
    ori = torch.atan2(gy, gx) * (180.0 / np.pi) + 180.0
    ori = torch.round(ori / 45.0) * 45.0
    thr_mag = mag.clone()
    thr_mag[mag < threshold] = 0.0
    out = {"blurred_img": blurred_img, "grad_magnitude": mag[:, None], "grad_orientation": ori[:, None],
           "thresholded_grad_magnitude": thr_mag[:, None]}
    if return_gradients:
        out["grad_x"], out["grad_y"] = gx[:, None], gy[:, None]
    if non_max_suppression:
        idx = (ori / 45) % 8
        thin = mag.clone()
        diff = [mag - _shift(mag, dy, dx) for dy, dx in NMS_NEIGHBOUR]
        for p in range(4):
            oriented = (idx == p) | (idx == p + 4)
            is_max = torch.minimum(diff[p], diff[p + 4]) > 0.0
            thin[oriented & ~is_max] = 0.0
        thr_thin = thin.clone()
        thr_thin[thin < threshold] = 0.0
        out["thin_edges"] = thin[:, None]
        out["thresholded_thin_edges"] = thr_thin[:, None]
    return out


def joints2d_to_heatmaps(joints2D, img_wh, std=4, visibility=None):
    """joints2D (B,K,2) as (u = column, v = row) -> (B,K,img_wh,img_wh); the reference's meshgrid is 'ij', so its `xx`
    is the ROW index and is paired with v (label_conversions.py:115-123). Optional (B,K) visibility mask."""
    joints2D = joints2D.float()
    rows = torch.arange(img_wh).float()[None, None, :, None]
    cols = torch.arange(img_wh).float()[None, None, None, :]
    u = joints2D[:, :, 0, None, None]
    v = joints2D[:, :, 1, None, None]
    h = torch.exp(-(((rows - v) / std) ** 2) / 2 - (((cols - u) / std) ** 2) / 2)
    if visibility is not None:
        h = h * visibility[:, :, None, None]
    return h


def proxy_representation(rgb, joints2D, visibility, threshold=0.0, non_max_suppression=True, std=4):
    """(B,3,S,S) image in [0,1], (B,17,2) joints, (B,17) visibility -> (B,18,S,S) (predict/...:91-100)."""
    e = canny_edges(rgb, threshold, non_max_suppression)
    edge = e["thresholded_thin_edges"] if non_max_suppression else e["thresholded_grad_magnitude"]
    heat = joints2d_to_heatmaps(joints2D, rgb.shape[-1], std, visibility)
    return torch.cat([edge, heat], dim=1).float()
