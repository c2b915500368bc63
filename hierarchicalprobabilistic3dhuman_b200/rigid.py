"""Rotation helpers on the hot path, same names/semantics as the reference's
utils/rigid_transform_utils.py (rot6d_to_rotmat :80-94) but executed by libhp3d kernels."""
import torch
from . import _lib


def rot6d_to_rotmat(x):
    """(B,6) or (B, K*6) 6D rotations -> (B*K,3,3). Unlike the reference (torch.cross without dim, wrong at
    B==3 -- SURVEY.md §7.6) the cross product is always taken per row."""
    _lib.require_cuda(x, "x")
    x = x.detach().to(torch.float32).contiguous().reshape(-1, 6)
    out = torch.empty(x.shape[0], 3, 3, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().hp3d_rot6d_to_rotmat(x.data_ptr(), x.shape[0], out.data_ptr(), _lib.stream_ptr()),
                   "hp3d_rot6d_to_rotmat")
    return out
