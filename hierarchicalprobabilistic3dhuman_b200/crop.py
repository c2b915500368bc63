"""Drop-ins for the crop / resample step between the 2D-pose network and the proxy representation:

  batch_crop_pytorch_affine(input_wh, output_wh, num_to_crop, device, joints2D=, rgb=, bbox_centres=, bbox_heights=,
                            bbox_widths=, orig_scale_factor=)      -- reference utils/image_utils.py:234-378
  get_kp_locations_confs_from_heatmaps(batch_heatmaps)            -- reference predict/predict_hrnet.py:7-30

Only the call shape of the predict path is implemented (given bounding box, RGB + 2D joints, no augmentation); the
training-time options (IUV / segmentation inputs, bounding boxes derived from them, random scale / centre jitter) raise
NotImplementedError. STATUS: the arithmetic (csrc/crop_math.h) is verified on the host bit for bit against the
reference-pinned oracle; the CUDA kernels have not yet run on hardware."""
import torch

from . import _lib


def batch_crop_pytorch_affine(input_wh, output_wh, num_to_crop, device=None, iuv=None, joints2D=None, rgb=None, seg=None,
                              bbox_determiner=None, bbox_centres=None, bbox_heights=None, bbox_widths=None, joints2D_vis=None,
                              orig_scale_factor=1.2, delta_scale_range=None, delta_centre_range=None, out_of_frame_pad_val=0):
    if iuv is not None or seg is not None or bbox_determiner is not None or delta_scale_range is not None \
            or delta_centre_range is not None or bbox_centres is None:
        raise NotImplementedError("libhp3d implements batch_crop_pytorch_affine for a given bounding box with rgb / joints2D "
                                  "inputs (the predict path, reference predict/...:84-93)")
    ref = rgb if rgb is not None else joints2D
    _lib.require_cuda(ref, "rgb / joints2D")
    dev = ref.device
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    B = int(num_to_crop)
    c, h, w = f32(bbox_centres), f32(bbox_heights), f32(bbox_widths)
    assert c.shape == (B, 2) and h.shape == (B,) and w.shape == (B,)
    in_w, in_h = int(input_wh[0]), int(input_wh[1])
    out_w, out_h = int(output_wh[0]), int(output_wh[1])
    x = j = xo = jo = None
    C = K = 0
    if rgb is not None:
        x = f32(rgb)
        C = x.shape[1]
        assert x.shape == (B, C, in_h, in_w)
        xo = torch.empty(B, C, out_h, out_w, device=dev, dtype=torch.float32)
    if joints2D is not None:
        j = f32(joints2D)
        K = j.shape[1]
        assert j.shape == (B, K, 2)
        jo = torch.empty(B, K, 2, device=dev, dtype=torch.float32)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().hp3d_crop_affine(ptr(x), ptr(j), B, C, in_h, in_w, K, c.data_ptr(), h.data_ptr(), w.data_ptr(),
                                               float(orig_scale_factor), out_w, out_h, ptr(xo), ptr(jo), _lib.stream_ptr()),
                   "hp3d_crop_affine")
    out = {}
    if jo is not None:
        out["joints2D"] = jo
    if xo is not None:
        out["rgb"] = xo
    return out


def get_kp_locations_confs_from_heatmaps(batch_heatmaps):
    """(B,K,h,w) -> (pred_kps (B,K,2), max_confs (B,K)) like the reference."""
    hm = _lib.require_cuda(batch_heatmaps, "batch_heatmaps").detach().to(torch.float32).contiguous()
    B, K, h, w = hm.shape
    kps = torch.empty(B, K, 2, device=hm.device, dtype=torch.float32)
    confs = torch.empty(B, K, device=hm.device, dtype=torch.float32)
    with torch.cuda.device(hm.device):
        _lib.check(_lib.lib().hp3d_heatmap_keypoints(hm.data_ptr(), B, K, h, w, kps.data_ptr(), confs.data_ptr(),
                                                     _lib.stream_ptr()), "hp3d_heatmap_keypoints")
    return kps, confs
