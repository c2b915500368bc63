"""Value tests of the reference-signature wrappers and call patterns around the sampler / SMPL kernels:
`compute_vertex_uncertainties_by_poseMF_shapeGaussian_sampling` with mean AND sampled shape (reference
utils/sampling_utils.py:146-192), the evaluation driver's call pattern (evaluate/evaluate_poseMF_shapeGaussian_net.py:
159-178), `transl`, argument validation, the SMPL model-file loader, and the sampler's shortfall report."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import rel_err
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
from oracle.smpl_oracle import SMPLOracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _usv(B, seed):
    """improper-capable random SVD factors like the head returns them"""
    rs = np.random.RandomState(seed)
    U = np.linalg.qr(rs.normal(size=(B, 23, 3, 3)))[0]
    V = np.linalg.qr(rs.normal(size=(B, 23, 3, 3)))[0]
    S = np.sort(np.exp(rs.uniform(np.log(0.05), np.log(40.0), size=(B, 23, 3))), axis=-1)[..., ::-1].copy()
    t = lambda a: torch.from_numpy(a.astype(np.float32)).cuda()
    return t(U), t(S), t(V)


def _oracle_uncertainty(vertices):
    """utils/sampling_utils.py:189-190"""
    mean = vertices.mean(dim=0)
    return torch.norm(vertices - mean, dim=-1).mean(dim=0)


@pytest.mark.parametrize("use_mean_shape", [True, False])
def test_compute_vertex_uncertainties_matches_oracle_composition(built_lib, use_mean_shape):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    smpl = hp.SMPL(model=model).cuda()
    N = 12
    U, S, V = _usv(1, 11)
    dist = torch.distributions.Normal(torch.randn(1, 10, device="cuda") * 1.25, torch.rand(1, 10, device="cuda") * 0.3 + 0.05)
    glob_R = hp.rot6d_to_rotmat(torch.randn(1, 6, device="cuda"))
    # replay the wrapper's generator use: sampler first, then (sampled shape only) Normal.sample
    torch.manual_seed(7)
    R = hp.pose_matrix_fisher_sampling_torch(U, S, V, N)
    betas = dist.loc.expand(N, -1) if use_mean_shape else dist.sample([N])[:, 0, :]
    torch.manual_seed(7)
    unc, verts, joints = hp.compute_vertex_uncertainties_by_poseMF_shapeGaussian_sampling(U, S, V, dist, glob_R, N, smpl,
                                                                                         use_mean_shape=use_mean_shape)
    ref = SMPLOracle(model, torch.float64).forward(betas.cpu(), R[0].cpu(), glob_R.cpu().expand(N, -1, -1)[:, None])
    assert verts.shape == (N, 6890, 3) and joints.shape == (N, 90, 3) and unc.shape == (6890,)
    assert rel_err(verts, ref["vertices"]) < TOL and rel_err(joints, ref["joints"]) < TOL
    assert rel_err(unc, _oracle_uncertainty(ref["vertices"])) < TOL
    if not use_mean_shape:      # the sampled betas really differ per sample
        assert (betas[0] - betas[1]).abs().max() > 1e-3


def test_evaluate_driver_call_pattern(built_lib):
    """evaluate/...:159-178: sampler with sample_on_cpu=True, rsample([N])[:,0,:] betas, SMPL on the N samples with the
    global orientation expanded, and the T-pose call with axis-angle zeros (default pose2rot=True)."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    smpl = hp.SMPL(model=model).cuda()
    N = 10
    U, S, V = _usv(1, 12)
    dist = torch.distributions.Normal(torch.randn(1, 10, device="cuda"), torch.rand(1, 10, device="cuda") * 0.2 + 0.05)
    glob_R = hp.rot6d_to_rotmat(torch.randn(1, 6, device="cuda"))
    R = hp.pose_matrix_fisher_sampling_torch(U, S, V, num_samples=N, b=1.5, oversampling_ratio=8, sample_on_cpu=True)
    shape_samples = dist.rsample([N])[:, 0, :]
    out = smpl(body_pose=R[0, :, :, :, :], global_orient=glob_R.unsqueeze(1).expand(N, -1, -1, -1), betas=shape_samples, pose2rot=False)
    orc = SMPLOracle(model, torch.float64)
    ref = orc.forward(shape_samples.cpu(), R[0].cpu(), glob_R.cpu().expand(N, -1, -1)[:, None])
    assert rel_err(out.vertices, ref["vertices"]) < TOL and rel_err(out.joints, ref["joints"]) < TOL
    tpose = smpl(body_pose=torch.zeros(N, 69, device="cuda"), global_orient=torch.zeros(N, 3, device="cuda"), betas=shape_samples)
    ref_t = orc.forward(shape_samples.cpu(), torch.zeros(N, 69), torch.zeros(N, 3), pose2rot=True)
    assert rel_err(tpose.vertices, ref_t["vertices"]) < TOL and rel_err(tpose.joints, ref_t["joints"]) < TOL


def test_transl_follows_smplx_then_reference_regressors(built_lib):
    """smplx adds transl to the vertices and its 45 joints; the reference's extra regressors (models/smpl_official.py:30-32)
    then see the TRANSLATED vertices, so joints 45..89 move by (regressor row sum) x transl."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    smpl = hp.SMPL(model=model).cuda()
    M = 5
    R = hp.rot6d_to_rotmat(torch.randn(M * 23, 6, device="cuda")).view(M, 23, 3, 3)
    gR = hp.rot6d_to_rotmat(torch.randn(M, 6, device="cuda"))
    betas = torch.randn(M, 10, device="cuda")
    t = torch.randn(M, 3, device="cuda") * 3
    out = smpl(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, transl=t, pose2rot=False)
    orc = SMPLOracle(model, torch.float64)
    ref = orc.forward(betas.cpu(), R.cpu(), gR.cpu()[:, None])
    v_ref = ref["vertices"] + t.cpu().double()[:, None]
    j45 = ref["joints"][:, :45] + t.cpu().double()[:, None]
    jx = torch.einsum("bik,ji->bjk", v_ref, orc.joint_regressors_extra)
    assert rel_err(out.vertices, v_ref) < TOL and rel_err(out.joints, torch.cat([j45, jx], 1)) < TOL


def test_pose2rot_false_requires_rotation_matrices(built_lib):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    smpl = hp.SMPL(model=syn.synthetic_smpl_model()).cuda()
    with pytest.raises(ValueError):
        smpl(betas=torch.zeros(1, 10, device="cuda"), pose2rot=False)


def _write_smpl_files(model, d):
    """the synthetic model in the SMPL file layout (what `SMPL_NEUTRAL.pkl` / .npz hold): posedirs (6890,3,207),
    shapedirs (6890,3,>=10), weights, J_regressor, kintree_table (2,24), f"""
    import scipy.sparse as sp
    kt = np.stack([np.where(model["parents"] < 0, 4294967295, model["parents"]).astype(np.int64), np.arange(24)])
    blob = dict(v_template=model["v_template"], shapedirs=np.concatenate([model["shapedirs"], np.zeros((6890, 3, 290))], -1),
                posedirs=model["posedirs"].T.reshape(6890, 3, 207).copy(), J_regressor=sp.csc_matrix(model["J_regressor"]),
                weights=model["lbs_weights"], kintree_table=kt, f=model["faces"])
    with open(os.path.join(d, "SMPL_NEUTRAL.pkl"), "wb") as f:
        pickle.dump(blob, f)
    dense = dict(blob, J_regressor=model["J_regressor"])
    np.savez(os.path.join(d, "SMPL_MALE.npz"), **dense)


@pytest.mark.parametrize("gender", ["neutral", "male"])
def test_smpl_model_file_round_trip(built_lib, tmp_path, gender):
    """A user with the licence-gated file gets a checked path: a model written in the SMPL .pkl (sparse J_regressor, 300
    shape components, uint32-max root parent) / .npz layout loads to the same constants and the same forward results."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    _write_smpl_files(model, str(tmp_path))
    a = hp.SMPL(str(tmp_path), batch_size=1, gender=gender).cuda()
    b = hp.SMPL(model=model).cuda()
    assert not a.is_synthetic and a.parents.tolist() == b.parents.tolist() and np.array_equal(a.faces, b.faces)
    M = 6
    R = hp.rot6d_to_rotmat(torch.randn(M * 23, 6, device="cuda")).view(M, 23, 3, 3)
    gR = hp.rot6d_to_rotmat(torch.randn(M, 6, device="cuda"))
    betas = torch.randn(M, 10, device="cuda")
    oa = a(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
    ob = b(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
    assert torch.equal(oa.vertices, ob.vertices) and torch.equal(oa.joints, ob.joints)


def test_sampler_shortfall_is_reported_not_silent(built_lib):
    """The reference redraws when fewer than N proposals are accepted (utils/sampling_utils.py:50,68-69); the kernel cannot
    draw more injected noise than it was given, so it must say so instead of returning mode-filled samples."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    U, S, V = _usv(2, 13)
    N = 8
    eps = torch.randn(2, 23, 8 * N, 4, device="cuda")
    w = torch.full((2, 23, 8 * N), 2.0, device="cuda")     # never < p_B / (M* p_ACG) <= 1 (+ rounding): nothing is accepted
    with pytest.raises(hp.SamplerShortfall):
        hp.pose_matrix_fisher_sampling_torch(U, S, V, N, noise=(eps, w))
    R, stats = hp.pose_matrix_fisher_sampling_torch(U, S, V, N, noise=(eps, w), return_stats=True)
    assert int(stats[2]) == 2 * 23 and int(stats[1]) == 0
    # Philox mode: the asynchronous status ring reports nothing for a healthy launch
    hp.pose_matrix_fisher_sampling_torch(U, S, V, N)
    hp.check_sampler_status()


def test_batched_composition_with_sampled_betas(built_lib):
    """train/train_poseMF_shapeGaussian_net.py:293-308: B images x N samples in one SMPL call with per-sample betas
    (`Normal.sample([N])` transposed to image-major) and the global orientation expanded per image."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    smpl = hp.SMPL(model=model).cuda()
    B, N = 3, 6
    U, S, V = _usv(B, 14)
    dist = torch.distributions.Normal(torch.randn(B, 10, device="cuda"), torch.rand(B, 10, device="cuda") * 0.3 + 0.05)
    glob_R = hp.rot6d_to_rotmat(torch.randn(B, 6, device="cuda"))
    torch.manual_seed(21)
    res = hp.sample_meshes_batched(U, S, V, dist, glob_R, N, smpl, use_mean_shape=False)
    betas = res["betas"]                                             # (B*N, 10), image-major
    assert betas.shape == (B * N, 10) and (betas[0] - betas[1]).abs().max() > 1e-3
    ref = SMPLOracle(model, torch.float64).forward(betas.cpu(), res["rotmats"].cpu().view(B * N, 23, 3, 3),
                                                   glob_R.cpu().repeat_interleave(N, 0)[:, None])
    v_ref = ref["vertices"].view(B, N, 6890, 3)
    assert rel_err(res["vertices"], v_ref) < TOL and rel_err(res["joints"], ref["joints"].view(B, N, 90, 3)) < TOL
    u_ref = (v_ref - v_ref.mean(1, keepdim=True)).norm(dim=-1).mean(1)
    assert rel_err(res["per_vertex_uncertainty"], u_ref) < TOL and rel_err(res["mean_vertices"], v_ref.mean(1)) < TOL
    # the sampled rotations are proper rotations
    R = res["rotmats"]
    assert (torch.linalg.det(R) - 1).abs().max() < 1e-5
