"""ctypes binding of libhp3d.so (C ABI in include/hp3d.h). There is NO fallback: if the CUDA
library is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os
from ctypes import c_int, c_void_p, c_size_t, c_float, c_uint64, c_char_p, POINTER, Structure, c_double, c_int32

from ._build import LIB_PATH

_lib = None

EXPORTS = [
    "hp3d_version", "hp3d_last_error",
    "hp3d_smpl_create", "hp3d_smpl_destroy", "hp3d_smpl_workspace_bytes", "hp3d_smpl_forward", "hp3d_smpl_forward_stats", "hp3d_smpl_layout_info",
    "hp3d_smpl_shape_blend", "hp3d_smpl_pose_blend_workspace_bytes", "hp3d_smpl_pose_blend", "hp3d_smpl_lbs", "hp3d_rodrigues", "hp3d_rot6d_to_rotmat",
    "hp3d_vertex_uncertainty", "hp3d_rank_samples_by_joints2d", "hp3d_mf_sample", "hp3d_mf_sample_sharded",
    "hp3d_head_create", "hp3d_head_destroy", "hp3d_head_workspace_bytes", "hp3d_head_forward",
    "hp3d_encoder_create", "hp3d_encoder_destroy", "hp3d_encoder_workspace_bytes", "hp3d_encoder_forward", "hp3d_encoder_forward_taps",
    "hp3d_encoder_forward_image", "hp3d_encoder_forward_argmax", "hp3d_encoder_forward_f16in", "hp3d_crop_affine", "hp3d_heatmap_keypoints", "hp3d_mf_log_norm_constant", "hp3d_canny_edges", "hp3d_joints2d_to_heatmaps", "hp3d_proxy_rep", "hp3d_joints2d_heatmap_argmax", "hp3d_peer_push", "hp3d_peer_push_multicast",
]


class SmplModel(Structure):
    _fields_ = [("v_template", c_void_p), ("shapedirs", c_void_p), ("posedirs", c_void_p), ("J_regressor", c_void_p),
                ("lbs_weights", c_void_p), ("parents", c_void_p), ("extra_vertex_ids", c_void_p),
                ("joint_regressors_extra", c_void_p)]


class HeadWeights(Structure):
    _fields_ = [("fc1_w", c_void_p), ("fc1_b", c_void_p), ("fc_shape_w", c_void_p), ("fc_shape_b", c_void_p),
                ("fc_glob_w", c_void_p), ("fc_glob_b", c_void_p), ("fc_cam_w", c_void_p), ("fc_cam_b", c_void_p),
                ("fc_embed_w", c_void_p), ("fc_embed_b", c_void_p),
                ("fc_pose0_w", POINTER(c_void_p)), ("fc_pose0_b", POINTER(c_void_p)),
                ("fc_pose2_w", POINTER(c_void_p)), ("fc_pose2_b", POINTER(c_void_p)),
                ("init_glob", c_void_p), ("init_cam", c_void_p), ("parents", c_void_p), ("delta_i_weight", c_float)]


class ConvBn(Structure):
    _fields_ = [("w", c_void_p), ("cout", c_int), ("cin", c_int), ("k", c_int), ("stride", c_int), ("pad", c_int),
                ("bn_w", c_void_p), ("bn_b", c_void_p), ("bn_mean", c_void_p), ("bn_var", c_void_p)]


class EncoderWeights(Structure):
    _fields_ = [("stem", ConvBn), ("conv", ConvBn * 2 * 2 * 4), ("down", ConvBn * 4), ("bn_eps", c_float)]


def lib():
    """Load libhp3d.so or fail loudly (the product has no CPU / PyTorch fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"libhp3d.so not found at {LIB_PATH}: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no fallback path.")
    L = ctypes.CDLL(LIB_PATH)
    L.hp3d_last_error.restype = c_char_p
    L.hp3d_smpl_workspace_bytes.restype = c_size_t
    L.hp3d_head_workspace_bytes.restype = c_size_t
    L.hp3d_encoder_workspace_bytes.restype = c_size_t
    L.hp3d_smpl_destroy.restype = None
    L.hp3d_head_destroy.restype = None
    L.hp3d_encoder_destroy.restype = None
    L.hp3d_smpl_create.argtypes = [POINTER(SmplModel), POINTER(c_void_p)]
    L.hp3d_smpl_destroy.argtypes = [c_void_p]
    L.hp3d_smpl_workspace_bytes.argtypes = [c_void_p, c_int, c_int]
    L.hp3d_smpl_forward.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_size_t, c_void_p]
    L.hp3d_smpl_forward_stats.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.hp3d_smpl_layout_info.argtypes = [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]
    L.hp3d_smpl_shape_blend.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.hp3d_smpl_pose_blend.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.hp3d_smpl_pose_blend_workspace_bytes.argtypes = [c_int]
    L.hp3d_smpl_pose_blend_workspace_bytes.restype = c_size_t
    L.hp3d_smpl_lbs.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                c_void_p, c_void_p]
    L.hp3d_rodrigues.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
    L.hp3d_rot6d_to_rotmat.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
    L.hp3d_vertex_uncertainty.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.hp3d_mf_sample.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_uint64, c_uint64,
                                 c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.hp3d_mf_sample_sharded.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_uint64, c_uint64, c_uint64,
                                         c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.hp3d_head_create.argtypes = [POINTER(HeadWeights), POINTER(c_void_p)]
    L.hp3d_head_destroy.argtypes = [c_void_p]
    L.hp3d_head_workspace_bytes.argtypes = [c_void_p, c_int]
    L.hp3d_head_forward.argtypes = [c_void_p, c_void_p, c_int] + [c_void_p] * 8 + [c_void_p] * 3 + [c_void_p, c_size_t, c_void_p]
    L.hp3d_encoder_create.argtypes = [POINTER(EncoderWeights), c_int, POINTER(c_void_p)]
    L.hp3d_encoder_destroy.argtypes = [c_void_p]
    L.hp3d_encoder_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int]
    L.hp3d_encoder_forward.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.hp3d_encoder_forward_taps.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]
    L.hp3d_rank_samples_by_joints2d.argtypes = [c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    LL = ctypes.c_longlong
    L.hp3d_canny_edges.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_float, c_int] + [c_void_p] * 7 + [LL, c_void_p]
    L.hp3d_joints2d_to_heatmaps.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, LL, c_void_p]
    L.hp3d_proxy_rep.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_float, c_int, c_float,
                                 c_void_p, c_void_p]
    L.hp3d_joints2d_heatmap_argmax.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]
    L.hp3d_encoder_forward_image.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_float, c_int,
                                             c_float, c_void_p, c_void_p, c_size_t, c_void_p]
    L.hp3d_crop_affine.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p]
    L.hp3d_heatmap_keypoints.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.hp3d_mf_log_norm_constant.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.hp3d_encoder_forward_argmax.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_float,
                                              c_void_p, c_void_p, c_void_p]
    L.hp3d_encoder_forward_f16in.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_float,
                                             c_void_p, c_void_p, c_void_p]
    L.hp3d_peer_push.argtypes = [c_void_p, POINTER(c_void_p), c_int, c_size_t, c_int, c_void_p]
    L.hp3d_peer_push_multicast.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_void_p]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().hp3d_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libhp3d {what} failed (status {rc}): {msg}")


def require_cuda(t, name):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the hot path runs only on the GPU (no CPU fallback)")
    return t


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class nvtx:
    """NVTX range around one stage of the path (`with _lib.nvtx("hp3d.encoder"):`), visible in nsys / ncu --nvtx timelines;
    HP3D_NVTX=0 turns the ranges off (they cost ~0.1 us each without a profiler attached)."""
    enabled = os.environ.get("HP3D_NVTX", "1") != "0"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if nvtx.enabled:
            import torch
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if nvtx.enabled:
            import torch
            torch.cuda.nvtx.range_pop()


class Workspace:
    """Grow-only per-device scratch buffer handed to the C ABI (caller-owned, torch storage)."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device):
        import torch
        key = str(device)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b
