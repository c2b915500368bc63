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


# ------------------------------------------------------------------------------------------------
# Sample ranking by 2D-joint consistency, restated from reference
# utils/sampling_utils.py:195-233 (`joints2D_error_sorted_verts_sampling`) and its helpers
# utils/cam_utils.py:9-16 (orthographic_project_torch), utils/joints2d_utils.py:5-10
# (undo_keypoint_normalisation), utils/label_conversions.py:17 (ALL_JOINTS_TO_COCO_MAP), :127-155
# (convert_heatmaps_to_2Djoints_coordinates_torch). The helpers are PINNED against the imported reference
# (tests/golden/rank_helpers.npz); the 180-degree flip about x goes through pytorch3d in the reference
# (utils/rigid_transform_utils.py:65-77), which is not installed: it is restated as diag(1,-1,-1) -- UNPINNED.
ALL_JOINTS_TO_COCO_MAP = [24, 26, 25, 28, 27, 16, 17, 18, 19, 20, 21, 1, 2, 4, 5, 7, 8]


def heatmaps_to_joints2d(heatmaps, eps=1e-6):
    """(B,17,H,W) -> joints2D (B,17,2) as (x, y) of the arg-max, -1 where invisible; vis (B,17) = max > eps."""
    B, K, H, W = heatmaps.shape
    mx, idx = torch.max(heatmaps.reshape(B, K, -1), dim=-1)
    j2d = torch.zeros(B, K, 2, dtype=heatmaps.dtype)
    j2d[:, :, 0] = (idx % W).to(heatmaps.dtype)
    j2d[:, :, 1] = torch.floor(idx / float(W)).to(heatmaps.dtype)
    vis = mx > eps
    j2d[~vis] = -1
    return j2d, vis


def project_joints_to_pixels(joints_coco, cam_wp, img_wh):
    """joints_coco (M,17,3), cam_wp (M,3) or (1,3) -> pixel coordinates (M,17,2): flip about x by 180 degrees,
    weak-perspective projection s*(X + t), then (p + 1) * img_wh / 2."""
    flipped = joints_coco * torch.tensor([1.0, -1.0, -1.0], dtype=joints_coco.dtype)
    proj = cam_wp[:, None, [0]] * (flipped[:, :, :2] + cam_wp[:, None, 1:])
    return (proj + 1) * (img_wh / 2.0)


def rank_samples(joints_samples, heatmaps, cam_wp):
    """joints_samples (B,N,90,3), heatmaps (B,17,H,W), cam_wp (B,3) -> (order (B,N) ascending by error, err (B,N))
    with err = max over visible COCO joints of the pixel distance to the heat-map arg-max (reference :222-227)."""
    B, N = joints_samples.shape[:2]
    W = heatmaps.shape[-1]
    j2d_in, vis = heatmaps_to_joints2d(heatmaps)
    err = torch.zeros(B, N, dtype=joints_samples.dtype)
    for b in range(B):
        px = project_joints_to_pixels(joints_samples[b][:, ALL_JOINTS_TO_COCO_MAP, :], cam_wp[b:b + 1], W)   # (N,17,2)
        d = torch.norm(px[:, vis[b], :] - j2d_in[b:b + 1, vis[b], :], dim=-1)
        err[b] = d.max(dim=-1).values if d.shape[1] > 0 else 0
    return torch.argsort(err, dim=1, stable=True), err
