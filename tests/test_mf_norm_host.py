"""The matrix-Fisher normalising constant arithmetic of csrc/mf_norm_math.h (shared by the CUDA kernel), compiled for the
host, against the reference's LogMFNormConstant (golden outputs, losses/matrix_fisher_loss.py:134-192) and its oracle."""
import ctypes

import numpy as np
import torch

from conftest import load_golden


def host_log_norm(built_lib, S):
    lib = ctypes.CDLL(built_lib.HOST_SHIM_PATH)
    S = np.ascontiguousarray(S, np.float32)
    n = S.shape[0]
    out, g = np.empty(n, np.float32), np.empty((n, 3), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.hp3d_host_mf_log_norm(p(S), ctypes.c_long(n), p(out), p(g))
    return out, g


def test_oracle_is_bit_identical_to_reference_golden():
    from oracle import mf_loss_oracle
    g = load_golden("mf_norm")
    lo, go = mf_loss_oracle.log_mf_norm_constant(torch.from_numpy(g["S"]))
    assert np.array_equal(lo.numpy(), g["log_c"]) and np.array_equal(go.numpy(), g["dlogc_ds"])


def test_device_math_matches_reference_golden(built_lib):
    g = load_golden("mf_norm")
    out, grad = host_log_norm(built_lib, g["S"])
    # the kernel sums the 512 nodes in a different order than torch.sum and uses libm's exp / sqrt
    assert np.abs(out - g["log_c"]).max() / np.abs(g["log_c"]).max() < 1e-6
    assert np.abs(grad - g["dlogc_ds"]).max() < 5e-6


def test_known_answers(built_lib):
    # S -> 0: uniform distribution on SO(3), c = 1, zero gradient
    out, grad = host_log_norm(built_lib, np.array([[1e-6, 1e-6, 1e-6]], np.float32))
    assert abs(out[0]) < 1e-4 and np.abs(grad).max() < 1e-4
    # the gradient is d log c / d s: check against central differences of log c itself
    S = np.array([[5.0, 3.0, 1.0], [0.3, 0.2, -0.1], [80.0, 60.0, 50.0]], np.float64)
    _, grad = host_log_norm(built_lib, S)
    h = 1e-2
    for k in range(3):
        e = np.zeros(3); e[k] = h
        fd = (host_log_norm(built_lib, S + e)[0].astype(np.float64) - host_log_norm(built_lib, S - e)[0]) / (2 * h)
        # rows 0-1 only: at S ~ 80, log c ~ 190 resolves 1.5e-5 in fp32, too coarse for a finite difference
        assert np.abs(fd - grad[:, k])[:2].max() < 2e-3, (k, fd, grad[:, k])
    # concentrated distributions: E[R] -> identity, i.e. every d log c / d s_k -> 1
    assert (grad[2] > 0.98).all() and (grad[2] < 1.0).all()


def test_sampler_oracle_mean_matches_normaliser_gradient(built_lib):
    """Cross-check of two independently restated pieces of the reference: for a matrix-Fisher distribution with
    F = U diag(S) V^T, E[R] = U_p diag(d log c / d s) V_p^T. The Monte-Carlo mean of the (reference-pinned) rejection
    sampler must agree with the (reference-pinned) normalising-constant gradient, including an improper factor pair."""
    from oracle import sampler_oracle
    S = torch.tensor([[[5.0, 3.0, 1.0], [0.3, 0.2, 0.1], [20.0, 15.0, 10.0], [4.0, 2.0, 0.5]]])
    U = torch.eye(3).expand(1, 4, 3, 3).clone()
    V = U.clone()
    U[0, 3, :, 2] *= -1.0                                    # det U = -1: proper s3 = -0.5
    N = 20000
    g = torch.Generator().manual_seed(3)
    R = sampler_oracle.sample(U, S, V, N, generator=g)       # (1, N, 4, 3, 3)
    Up, Sp, Vp = sampler_oracle.proper_usv(U, S, V)
    _, grad = host_log_norm(built_lib, Sp[0].numpy())
    ER = torch.einsum("jab,jb,jcb->jac", Up[0], torch.from_numpy(grad), Vp[0])
    err = (R[0].mean(0) - ER).abs().max().item()
    assert err < 4.0 / np.sqrt(N), err
