"""The crop / affine-resample arithmetic of csrc/crop_math.h (shared by the CUDA kernel), compiled for the host, against
the oracle that is bit-pinned to the reference's batch_crop_pytorch_affine (utils/image_utils.py:234-378)."""
import ctypes

import numpy as np
import torch

from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn


def host_crop(built_lib, rgb, j2d, c, h, w, scale, out_wh=(256, 256)):
    lib = ctypes.CDLL(built_lib.HOST_SHIM_PATH)
    B, C, H, W = rgb.shape
    K = j2d.shape[1]
    out = np.empty((B, C, out_wh[1], out_wh[0]), np.float32)
    jo = np.empty((B, K, 2), np.float32)
    p = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(ctypes.c_void_p)
    keep = [np.ascontiguousarray(a, np.float32) for a in (rgb, j2d, c, h, w)]
    q = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.hp3d_host_crop(q(keep[0]), q(keep[1]), B, C, H, W, K, q(keep[2]), q(keep[3]), q(keep[4]), ctypes.c_float(scale),
                       out_wh[0], out_wh[1], q(out), q(jo))
    return out, jo


def test_crop_math_is_bit_identical_to_the_pinned_oracle(built_lib):
    from oracle import crop_oracle
    rgb, j2d, c, h, w = syn.synthetic_crop_inputs(3, seed=8)
    for scale in (1.0, 1.2):
        out, jo = host_crop(built_lib, rgb, j2d, c, h, w, scale)
        o = crop_oracle.batch_crop_affine((288, 384), (256, 256), *(torch.from_numpy(a) for a in (j2d, rgb, c, h, w)), scale)
        assert np.array_equal(jo, o["joints2D"].numpy()), scale
        assert np.array_equal(out, o["rgb"].numpy()), scale
    # boxes far outside the image and non-square outputs
    c2 = np.array([[-50.0, 400.0], [500.0, -100.0], [192.0, 144.0]], np.float32)
    out, jo = host_crop(built_lib, rgb, j2d, c2, h, w, 1.2, out_wh=(192, 128))
    o = crop_oracle.batch_crop_affine((288, 384), (192, 128), *(torch.from_numpy(a) for a in (j2d, rgb, c2, h, w)), 1.2)
    assert np.array_equal(out, o["rgb"].numpy()) and np.array_equal(jo, o["joints2D"].numpy())
