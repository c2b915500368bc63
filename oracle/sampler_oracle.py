"""Matrix-Fisher sampler (Bingham rejection from an ACG envelope -> quaternion -> rotation),
restated from reference utils/sampling_utils.py:10-71 (bingham_sampling_for_matrix_fisher_torch)
and :74-143 (pose_matrix_fisher_sampling_torch). TEST INFRASTRUCTURE -- see oracle/__init__.py.
PINNED against the unmodified reference by oracle/make_golden.py.

Noise is explicit: `draw_noise` consumes torch's CPU generator in the reference's draw order --
per (image, joint): randn(8N,4) then rand(8N) (reference :51,:60) -- so seeding + this function
reproduces exactly what the reference consumes when no retry happens.
"""
import math
import torch
from .net_oracle import quat_to_rotmat


def m_star(b):
    return math.exp(-(4 - b) / 2) * ((4 / b) ** 2)         # reference :46-47


def draw_noise(batch, joints, num_samples, oversampling_ratio=8, generator=None):
    n = num_samples * oversampling_ratio
    eps = torch.empty(batch, joints, n, 4)
    w = torch.empty(batch, joints, n)
    for i in range(batch):
        for j in range(joints):
            eps[i, j] = torch.randn(n, 4, generator=generator)
            w[i, j] = torch.rand(n, generator=generator)
    return eps, w


def proper_usv(U, S, V):
    """reference :104-111 -- U_p = U diag(1,1,det U), V_p likewise, s3 *= det U det V."""
    dU, dV = torch.det(U), torch.det(V)
    Up, Sp, Vp = U.clone(), S.clone(), V.clone()
    Sp[..., 2] *= dU * dV
    Up[..., :, 2] *= dU[..., None]
    Vp[..., :, 2] *= dV[..., None]
    return Up, Sp, Vp


def sample_with_noise(U, S, V, num_samples, eps, w, b=1.5):
    """U,V (B,J,3,3), S (B,J,3), eps (B,J,8N,4), w (B,J,8N) -> R (B,N,J,3,3), accepted (B,J) counts.
    Raises if any (image, joint) accepts fewer than N (the reference would redraw, :68-69)."""
    B, J = U.shape[:2]
    Up, Sp, Vp = proper_usv(U, S, V)
    A = torch.zeros(B, J, 4, dtype=S.dtype)
    A[..., 1] = 2 * (Sp[..., 1] + Sp[..., 2])
    A[..., 2] = 2 * (Sp[..., 0] + Sp[..., 2])
    A[..., 3] = 2 * (Sp[..., 0] + Sp[..., 1])
    Omega = torch.ones(B, J, 4, dtype=S.dtype) + 2 * A / b
    sigma = Omega ** (-0.5)
    Ms = m_star(b)
    quats = torch.zeros(B, num_samples, J, 4, dtype=S.dtype)
    acc = torch.zeros(B, J, dtype=torch.long)
    for i in range(B):
        for j in range(J):
            y = sigma[i, j] * eps[i, j]
            x = y / torch.norm(y, dim=1, keepdim=True)
            p_b = torch.exp(-torch.einsum("bn,n,bn->b", x, A[i, j], x))
            p_a = torch.einsum("bn,n,bn->b", x, Omega[i, j], x) ** (-2)
            ok = w[i, j] < p_b / (Ms * p_a)
            acc[i, j] = int(ok.sum())
            if acc[i, j] < num_samples:
                raise RuntimeError(f"fewer than N accepted at ({i},{j}): {int(acc[i, j])}")
            quats[i, :, j] = x[ok][:num_samples]
    R = quat_to_rotmat(quats.view(-1, 4)).view(B, num_samples, J, 3, 3)
    R = torch.matmul(Up[:, None], torch.matmul(R, Vp[:, None].transpose(-1, -2)))
    return R, acc


def sample(U, S, V, num_samples, b=1.5, oversampling_ratio=8, generator=None):
    eps, w = draw_noise(U.shape[0], U.shape[1], num_samples, oversampling_ratio, generator)
    return sample_with_noise(U, S, V, num_samples, eps, w, b)[0]
