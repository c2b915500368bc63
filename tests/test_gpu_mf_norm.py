"""GPU parity of the matrix-Fisher normalising-constant kernel against the reference golden
(first run on a B200 in round 2: `profiles/r02a_unverified.log`); the arithmetic is additionally verified on the host
by tests/test_mf_norm_host.py."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_log_norm_constant_matches_reference_golden(built_lib):
    from hierarchicalprobabilistic3dhuman_b200.mf_loss import LogMFNormConstant
    g = load_golden("mf_norm")
    S = torch.from_numpy(g["S"]).cuda().requires_grad_(True)
    lc = LogMFNormConstant.apply(S)
    lc.sum().backward()
    assert (lc.detach().cpu().numpy() - g["log_c"]).__abs__().max() / np.abs(g["log_c"]).max() < 1e-6
    assert np.abs(S.grad.cpu().numpy() - g["dlogc_ds"]).max() < 5e-6


def test_expected_rotation_matches_sampler_mean(built_lib):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from hierarchicalprobabilistic3dhuman_b200.mf_loss import matrix_fisher_expected_rotation
    S = torch.tensor([[5.0, 3.0, 1.0]] * 8 + [[0.3, 0.2, 0.1]] * 8 + [[80.0, 60.0, 50.0]] * 7)[None].cuda()
    U = torch.eye(3).expand(1, 23, 3, 3).contiguous().cuda()
    N = 20000
    R = hp.pose_matrix_fisher_sampling_torch(U, S, U.clone(), N)
    ER = matrix_fisher_expected_rotation(U, S, U.clone()).view(1, 23, 3, 3)
    assert (R.mean(1) - ER).abs().max() < 4.0 / np.sqrt(N)
