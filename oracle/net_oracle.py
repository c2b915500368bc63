"""ResNet-18 proxy-rep encoder + hierarchical matrix-Fisher head + rotation utils, restated as
plain functions of a reference-format state_dict (CPU torch). TEST INFRASTRUCTURE -- see
oracle/__init__.py. PINNED against the unmodified reference by oracle/make_golden.py.

Follows: reference models/resnet.py:62-78 (BasicBlock), :146-158,202-217 (stem, stages, pooling);
models/poseMF_shapeGaussian_net.py:14-21 (ancestors), :85-162 (forward);
utils/rigid_transform_utils.py:80-94 (rot6d_to_rotmat), :113-133 (quat_to_rotmat).
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by reference models/resnet.py:148


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], training=False, eps=BN_EPS)


def encoder_forward(sd, x, prefix="image_encoder.", taps=None):
    """(B,18,H,W) -> (B,512). `taps` (optional dict) receives every post-activation tensor."""
    p = prefix
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], stride=2, padding=3), sd, p + "bn1"))
    if taps is not None: taps["stem"] = x
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None: taps["pool"] = x
    for li in range(1, 5):
        for bi in range(2):
            q = f"{p}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            out = F.relu(_bn(F.conv2d(x, sd[q + "conv1.weight"], stride=stride, padding=1), sd, q + "bn1"))
            out = _bn(F.conv2d(out, sd[q + "conv2.weight"], stride=1, padding=1), sd, q + "bn2")
            if (q + "downsample.0.weight") in sd:
                x = _bn(F.conv2d(x, sd[q + "downsample.0.weight"], stride=stride), sd, q + "downsample.1")
            x = F.relu(out + x)
            if taps is not None: taps[f"layer{li}.{bi}"] = x
    return torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)


def ancestors(parents):
    """reference models/poseMF_shapeGaussian_net.py:14-21 (nearest ancestor first)."""
    anc = {}
    for i in range(1, len(parents)):
        ip = int(parents[i]) - 1
        anc[i - 1] = ([ip] + anc[ip]) if ip >= 0 else []
    return anc


def head_forward(sd, feats, parents, delta_i_weight=1.0, svd=None, teacher=None):
    """feats (B,512) -> dict(F,U,S,V,mode (B,23,..), shape_params (B,20), glob (B,6), cam (B,3),
    U_proper, S_proper). `svd` defaults to torch.svd on CPU (LAPACK), as the reference does (:137).
    `teacher` (optional dict with U_proper/S_proper/mode) teacher-forces the ancestors' inputs."""
    svd = svd or torch.svd
    B = feats.shape[0]
    lin = lambda n, v: F.linear(v, sd[n + ".weight"], sd[n + ".bias"])
    x = F.elu(lin("fc1", feats))
    shape_params = lin("fc_shape", x)
    glob = lin("fc_glob", x) + sd["init_glob"]
    cam = lin("fc_cam", x) + sd["init_cam"]
    embed = F.elu(lin("fc_embed", torch.cat([feats, shape_params, glob, cam], dim=1)))
    anc = ancestors(parents)
    nj = len(anc)
    z = lambda *s: torch.zeros(B, nj, *s, dtype=feats.dtype)
    out = dict(F=z(3, 3), U=z(3, 3), S=z(3), V=z(3, 3), mode=z(3, 3), U_proper=z(3, 3), S_proper=z(3))
    src = teacher if teacher is not None else out
    for j in range(nj):
        a = anc[j]
        if a:
            inp = torch.cat([embed, src["U_proper"][:, a].reshape(B, -1), src["S_proper"][:, a].reshape(B, -1),
                             src["mode"][:, a].reshape(B, -1)], dim=1)
        else:
            inp = embed
        Fj = lin(f"fc_pose.{j}.2", F.elu(lin(f"fc_pose.{j}.0", inp))).view(B, 3, 3)
        Fj = Fj + delta_i_weight * torch.eye(3, dtype=feats.dtype)[None]
        U, S, V = svd(Fj)
        dU, dV = torch.det(U), torch.det(V)
        Up, Sp, Vp = U.clone(), S.clone(), V.clone()
        Up[:, :, 2] *= dU[:, None]
        Sp[:, 2] *= dU * dV
        Vp[:, :, 2] *= dV[:, None]
        out["F"][:, j], out["U"][:, j], out["S"][:, j], out["V"][:, j] = Fj, U, S, V
        out["U_proper"][:, j], out["S_proper"][:, j] = Up, Sp
        out["mode"][:, j] = torch.matmul(Up, Vp.transpose(-1, -2))
    out.update(shape_params=shape_params, glob=glob, cam=cam, embed=embed)
    return out


def rot6d_to_rotmat(x):
    """reference utils/rigid_transform_utils.py:80-94, with the cross product pinned to dim=1
    (the reference's dim-less torch.cross misbehaves at B==3, SURVEY.md §7.6 -- not reproduced)."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def quat_to_rotmat(q):
    """reference utils/rigid_transform_utils.py:113-133, (w,x,y,z), re-normalised."""
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)
