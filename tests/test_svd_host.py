"""The in-kernel LAPACK-convention 3x3 SVD (csrc/svd3.h), compiled for the host, against
torch.svd (what the reference calls on the CPU, models/poseMF_shapeGaussian_net.py:137)."""
import ctypes

import numpy as np
import pytest
import torch


def host_svd(built_lib, A):
    lib = ctypes.CDLL(built_lib.HOST_SHIM_PATH)
    A = np.ascontiguousarray(A, np.float32)
    n = A.shape[0]
    U, V, S = np.empty_like(A), np.empty_like(A), np.empty((n, 3), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.hp3d_host_svd3(p(A), ctypes.c_long(n), p(U), p(S), p(V))
    return U, S, V


@pytest.mark.parametrize("sigma", [0.05, 0.5, 3.0])
def test_sign_convention_matches_lapack(built_lib, sigma):
    rs = np.random.RandomState(int(sigma * 100))
    A = (np.eye(3)[None] + sigma * rs.normal(size=(20000, 3, 3))).astype(np.float32)
    U, S, V = host_svd(built_lib, A)
    Ut, St, Vt = (t.numpy() for t in torch.svd(torch.from_numpy(A)))
    assert np.abs(np.einsum("nij,nj,nkj->nik", U, S, V) - A).max() < 5e-5
    assert np.abs(S - St).max() < 2e-5
    assert (np.diff(S, axis=1) <= 0).all() and (S >= 0).all()
    agree = ((U * Ut).sum(1) > 0) & ((V * Vt).sum(1) > 0)       # per-column sign agreement
    assert agree.mean() > 0.999, agree.mean()                   # measured 0.9998-0.99996
    # orthogonality
    assert np.abs(np.einsum("nij,nik->njk", U, U) - np.eye(3)).max() < 1e-5
    assert np.abs(np.einsum("nij,nik->njk", V, V) - np.eye(3)).max() < 1e-5


def test_special_matrices(built_lib):
    A = np.stack([np.eye(3), np.diag([3.0, 2.0, 1.0]), np.diag([1.0, 2.0, 3.0]), -np.eye(3),
                  np.zeros((3, 3)), np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0.]]),
                  np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9.]])]).astype(np.float32)
    U, S, V = host_svd(built_lib, A)
    assert np.isfinite(U).all() and np.isfinite(S).all() and np.isfinite(V).all()
    assert np.abs(np.einsum("nij,nj,nkj->nik", U, S, V) - A).max() < 1e-5
    St = torch.svd(torch.from_numpy(A))[1].numpy()
    assert np.abs(S - St).max() < 1e-5
