"""Drop-ins for the proxy-representation generation that precedes the network:

  CannyEdgeDetector(non_max_suppression, gaussian_filter_std, gaussian_filter_size, threshold).forward(img)
      -- reference models/canny_edge_detector.py:11-166, same constructor, same output dict keys;
  convert_2Djoints_to_gaussian_heatmaps_torch(joints2D, img_wh, std)
      -- reference utils/label_conversions.py:105-124;
  proxy_representation(rgb, joints2D, visibility, ...) -- predict/predict_poseMF_shapeGaussian_net.py:91-100 in
      one kernel launch.
All arithmetic runs in libhp3d (csrc/proxy.cu); CUDA tensors only, no fallback."""
import torch
from torch import nn

from . import _lib


def _f32(t, name):
    return _lib.require_cuda(t, name).detach().to(torch.float32).contiguous()


def _vis_bytes(visibility, dev):
    if visibility is None:
        return None
    return visibility.detach().to(device=dev).ne(0).to(torch.uint8).contiguous()


class CannyEdgeDetector(nn.Module):
    def __init__(self, non_max_suppression=True, gaussian_filter_std=1.0, gaussian_filter_size=5, threshold=0.2):
        super().__init__()
        self.threshold = threshold
        self.non_max_suppression = non_max_suppression
        self.gaussian_filter_std = float(gaussian_filter_std)
        self.gaussian_filter_size = int(gaussian_filter_size)

    def forward(self, img):
        """img (B,C,H,W) -> dict(blurred_img, grad_magnitude, grad_orientation, thresholded_grad_magnitude
        [, thin_edges, thresholded_thin_edges]) exactly like the reference (:154-166)."""
        x = _f32(img, "img")
        B, C, H, W = x.shape
        new = lambda *s: torch.empty(*s, device=x.device, dtype=torch.float32)
        out = {"blurred_img": new(B, C, H, W), "grad_magnitude": new(B, 1, H, W), "grad_orientation": new(B, 1, H, W),
               "thresholded_grad_magnitude": new(B, 1, H, W)}
        if self.non_max_suppression:
            out["thin_edges"] = new(B, 1, H, W)
            out["thresholded_thin_edges"] = new(B, 1, H, W)
        ptr = lambda k: out[k].data_ptr() if k in out else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().hp3d_canny_edges(
                x.data_ptr(), B, C, H, W, self.gaussian_filter_std, self.gaussian_filter_size, float(self.threshold),
                int(bool(self.non_max_suppression)), ptr("blurred_img"), ptr("grad_magnitude"), ptr("grad_orientation"),
                ptr("thresholded_grad_magnitude"), ptr("thin_edges"), ptr("thresholded_thin_edges"), None, 0,
                _lib.stream_ptr()), "hp3d_canny_edges")
        return out


def convert_2Djoints_to_gaussian_heatmaps_torch(joints2D, img_wh, std=4, visibility=None):
    """joints2D (B,K,2) -> (B,K,img_wh,img_wh); `visibility` (B,K) optionally applies predict/...:97-99's mask."""
    j = _f32(joints2D, "joints2D")
    B, K = j.shape[:2]
    vis = _vis_bytes(visibility, j.device)
    out = torch.empty(B, K, img_wh, img_wh, device=j.device, dtype=torch.float32)
    with torch.cuda.device(j.device):
        _lib.check(_lib.lib().hp3d_joints2d_to_heatmaps(j.data_ptr(), vis.data_ptr() if vis is not None else None, B, K,
                                                        int(img_wh), float(std), out.data_ptr(), K * img_wh * img_wh,
                                                        _lib.stream_ptr()), "hp3d_joints2d_to_heatmaps")
    return out


def proxy_representation(rgb, joints2D, visibility=None, threshold=0.0, non_max_suppression=True, gaussian_filter_std=1.0,
                         gaussian_filter_size=5, heatmap_std=4.0, out=None):
    """(B,C,S,S) image in [0,1], (B,K,2) joints, (B,K) visibility -> (B,K+1,S,S) proxy representation
    (edge map | masked joint heat-maps), predict/...:91-100 with DATA.EDGE_* / HEATMAP_GAUSSIAN_STD defaults."""
    x = _f32(rgb, "rgb")
    j = _f32(joints2D, "joints2D")
    B, C, S, S2 = x.shape
    K = j.shape[1]
    assert S == S2 and j.shape[0] == B
    vis = _vis_bytes(visibility, x.device)
    if out is None:
        out = torch.empty(B, K + 1, S, S, device=x.device, dtype=torch.float32)
    assert out.is_contiguous() and out.shape == (B, K + 1, S, S)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().hp3d_proxy_rep(x.data_ptr(), j.data_ptr(), vis.data_ptr() if vis is not None else None, B, C, K,
                                             S, float(gaussian_filter_std), int(gaussian_filter_size), float(threshold),
                                             int(bool(non_max_suppression)), float(heatmap_std), out.data_ptr(),
                                             _lib.stream_ptr()), "hp3d_proxy_rep")
    return out


def joints2d_heatmap_argmax(joints2D, visibility=None, img_wh=256, std=4.0, eps=1e-6):
    """What utils/label_conversions.py:127-155 would return for the heat-maps of these joints, without building them:
    (joints2D_px (B,K,2) float, vis (B,K) int32)."""
    j = _f32(joints2D, "joints2D")
    B, K = j.shape[:2]
    vis = _vis_bytes(visibility, j.device)
    px = torch.empty(B, K, 2, device=j.device, dtype=torch.float32)
    vo = torch.empty(B, K, device=j.device, dtype=torch.int32)
    with torch.cuda.device(j.device):
        _lib.check(_lib.lib().hp3d_joints2d_heatmap_argmax(j.data_ptr(), vis.data_ptr() if vis is not None else None, B, K,
                                                           int(img_wh), float(std), float(eps), px.data_ptr(), vo.data_ptr(),
                                                           _lib.stream_ptr()), "hp3d_joints2d_heatmap_argmax")
    return px, vo
