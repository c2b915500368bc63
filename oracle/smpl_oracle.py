"""SMPL forward restated op-for-op after smplx 0.1.26 (`smplx/lbs.py: lbs, blend_shapes,
vertices2joints, batch_rodrigues, batch_rigid_transform`; `smplx/body_models.py: SMPL.forward`;
`smplx/vertex_joint_selector.py`) plus the reference's joint extension
(reference models/smpl_official.py:27-41). TEST INFRASTRUCTURE -- see oracle/__init__.py.

PARITY UNPINNED: smplx is a third-party dependency (reference requirements.txt:10, pinned 0.1.26)
that is neither under /root/reference nor installable here; the arithmetic below follows
SURVEY.md §8c steps 1-9. Works in float64 (checker) or float32 (CPU baseline, mirrors smplx's
materialised intermediates).
"""
import torch


class SMPLOracle:
    def __init__(self, model, dtype=torch.float64):
        t = lambda a: torch.as_tensor(a, dtype=dtype)
        self.dtype = dtype
        self.v_template = t(model["v_template"])                 # (6890,3)
        self.shapedirs = t(model["shapedirs"])                   # (6890,3,10)
        self.posedirs = t(model["posedirs"])                     # (207,20670)
        self.J_regressor = t(model["J_regressor"])               # (24,6890)
        self.lbs_weights = t(model["lbs_weights"])               # (6890,24)
        self.parents = [int(p) for p in model["parents"]]
        self.extra_vertex_ids = torch.as_tensor(model["extra_vertex_ids"], dtype=torch.long)
        self.joint_regressors_extra = t(model["joint_regressors_extra"])  # (45,6890)

    @staticmethod
    def batch_rodrigues(rot_vecs):
        """smplx lbs.batch_rodrigues: angle = ||r + 1e-8||, R = I + sin K + (1-cos) K^2."""
        n = rot_vecs.shape[0]
        angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
        rot_dir = rot_vecs / angle
        cos = torch.cos(angle)[:, None]
        sin = torch.sin(angle)[:, None]
        rx, ry, rz = torch.split(rot_dir, 1, dim=1)
        zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype)
        K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
        ident = torch.eye(3, dtype=rot_vecs.dtype)[None]
        return ident + sin * K + (1 - cos) * torch.bmm(K, K)

    def forward(self, betas, body_pose, global_orient, pose2rot=False):
        """betas (Mb,10); body_pose (M,23,3,3) / global_orient (M,1,3,3) when pose2rot=False, else
        axis-angle (M,69) / (M,3). Returns dict(vertices (M,6890,3), joints (M,90,3), v_posed, J, A)."""
        dt = self.dtype
        betas = betas.to(dt)
        if pose2rot:
            full = torch.cat([global_orient.reshape(-1, 3), body_pose.reshape(-1, 69)], dim=1).to(dt)
            M = full.shape[0]
            rot_mats = self.batch_rodrigues(full.reshape(-1, 3)).view(M, 24, 3, 3)
        else:
            rot_mats = torch.cat([global_orient.reshape(-1, 1, 3, 3), body_pose.reshape(-1, 23, 3, 3)], dim=1).to(dt)
            M = rot_mats.shape[0]
        if betas.shape[0] != M:   # smplx expands betas to the pose batch
            betas = betas.repeat_interleave(M // betas.shape[0], dim=0) if betas.shape[0] > 1 else betas.expand(M, -1)
        # 2. shape blend
        v_shaped = self.v_template[None] + torch.einsum("bl,mkl->bmk", betas, self.shapedirs)
        # 3. joints from the shaped, unposed mesh
        J = torch.einsum("bik,ji->bjk", v_shaped, self.J_regressor)
        # 5. pose-corrective blend
        ident = torch.eye(3, dtype=dt)
        pose_feature = (rot_mats[:, 1:] - ident).reshape(M, 207)
        v_posed = v_shaped + torch.matmul(pose_feature, self.posedirs).view(M, -1, 3)
        # 6. forward kinematics
        rel_J = J.clone()
        par = torch.as_tensor(self.parents[1:], dtype=torch.long)
        rel_J[:, 1:] = J[:, 1:] - J[:, par]
        L = torch.zeros(M, 24, 4, 4, dtype=dt)
        L[:, :, :3, :3] = rot_mats
        L[:, :, :3, 3] = rel_J
        L[:, :, 3, 3] = 1
        G = [L[:, 0]]
        for i in range(1, 24):
            G.append(torch.matmul(G[self.parents[i]], L[:, i]))
        G = torch.stack(G, dim=1)
        posed_joints = G[:, :, :3, 3]
        Jh = torch.cat([J, torch.zeros(M, 24, 1, dtype=dt)], dim=2)[..., None]      # (M,24,4,1)
        corr = torch.matmul(G, Jh)                                                    # (M,24,4,1)
        A = G.clone()
        A[:, :, :, 3] = A[:, :, :, 3] - corr[..., 0]
        # 7. skinning
        T = torch.matmul(self.lbs_weights[None].expand(M, -1, -1), A.view(M, 24, 16)).view(M, -1, 4, 4)
        vh = torch.cat([v_posed, torch.ones(M, v_posed.shape[1], 1, dtype=dt)], dim=2)
        verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]
        # 8. 24 posed joints + 21 picked vertices; 9. + 45 regressed joints
        joints45 = torch.cat([posed_joints, verts[:, self.extra_vertex_ids]], dim=1)
        extra = torch.einsum("bik,ji->bjk", verts, self.joint_regressors_extra)
        joints = torch.cat([joints45, extra], dim=1)
        return dict(vertices=verts, joints=joints, v_posed=v_posed, J=J, A=A[:, :, :3, :], v_shaped=v_shaped)
