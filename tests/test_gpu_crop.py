"""GPU parity of the crop / affine-resample kernels against the reference-pinned oracle
(first run on a B200 in round 2: `profiles/r02a_unverified.log`); the arithmetic (csrc/crop_math.h) is additionally
verified on the host by tests/test_crop_host.py."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def test_crop_matches_reference_golden(built_lib):
    from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
    from hierarchicalprobabilistic3dhuman_b200.crop import batch_crop_pytorch_affine
    g = load_golden("crop_b3")
    rgb, j2d, c, h, w = (torch.from_numpy(a).cuda() for a in syn.synthetic_crop_inputs(3, seed=int(g["crop_seed"])))
    for tag, scale in (("s10", 1.0), ("s12", 1.2)):
        o = batch_crop_pytorch_affine((288, 384), (256, 256), 3, "cuda", joints2D=j2d, rgb=rgb, bbox_centres=c, bbox_heights=h,
                                      bbox_widths=w, orig_scale_factor=scale)
        assert rel_err(o["joints2D"], g[f"joints2D_{tag}"]) < 1e-6
        assert rel_err(o["rgb"][:, :, ::31, :], g[f"rgb_rows_{tag}"]) < 1e-5


def test_hrnet_keypoints_match_reference_golden(built_lib):
    from hierarchicalprobabilistic3dhuman_b200.crop import get_kp_locations_confs_from_heatmaps
    g = load_golden("crop_b3")
    rs = np.random.RandomState(4)
    hm = torch.from_numpy(rs.normal(size=(2, 17, 96, 72)).astype(np.float32))
    hm[0, 3] = -1.0
    kps, confs = get_kp_locations_confs_from_heatmaps(hm.cuda())
    assert np.array_equal(kps.cpu().numpy(), g["hrnet_kps"]) and np.array_equal(confs.cpu().numpy(), g["hrnet_confs"])
