"""Known-answer algebra for the SMPL restatement (parity unpinned by the reference: smplx absent)."""
import numpy as np
import torch

from oracle.smpl_oracle import SMPLOracle
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn

MODEL = syn.synthetic_smpl_model()
ORACLE = SMPLOracle(MODEL, torch.float64)
EYE = torch.eye(3, dtype=torch.float64)


def test_model_shapes_and_partition_of_unity():
    assert MODEL["posedirs"].shape == (207, 20670) and MODEL["shapedirs"].shape == (6890, 3, 10)
    np.testing.assert_allclose(MODEL["lbs_weights"].sum(1), 1.0, atol=1e-12)
    np.testing.assert_allclose(MODEL["J_regressor"].sum(1), 1.0, atol=1e-12)
    assert ((MODEL["lbs_weights"] != 0).sum(1) <= 4).all()
    assert MODEL["joint_regressors_extra"].shape == (45, 6890) and (MODEL["joint_regressors_extra"] != 0).sum() == 255


def test_zero_pose_zero_betas_gives_template():
    out = ORACLE.forward(torch.zeros(1, 10), EYE.expand(1, 23, 3, 3), EYE.expand(1, 1, 3, 3))
    vt = torch.as_tensor(MODEL["v_template"])
    assert (out["vertices"][0] - vt).abs().max() < 1e-12
    J = torch.as_tensor(MODEL["J_regressor"]) @ vt
    assert (out["joints"][0, :24] - J).abs().max() < 1e-12
    assert (out["joints"][0, 24:45] - vt[MODEL["extra_vertex_ids"]]).abs().max() < 1e-12
    assert (out["joints"][0, 45:] - torch.as_tensor(MODEL["joint_regressors_extra"]) @ vt).abs().max() < 1e-12
    assert out["joints"].shape == (1, 90, 3)


def test_global_orient_is_rigid_about_root():
    rs = np.random.RandomState(0)
    R = torch.as_tensor(syn.random_rotmats(rs, (1,)))
    betas = torch.as_tensor(rs.normal(size=(1, 10)))
    base = ORACLE.forward(betas, EYE.expand(1, 23, 3, 3), EYE.expand(1, 1, 3, 3))
    rot = ORACLE.forward(betas, EYE.expand(1, 23, 3, 3), R[:, None])
    root = base["J"][0, 0]
    expect = (base["vertices"][0] - root) @ R[0].T + root
    assert (rot["vertices"][0] - expect).abs().max() < 1e-10


def test_single_joint_rotation_moves_only_weighted_vertices():
    rs = np.random.RandomState(1)
    pose = EYE.expand(1, 23, 3, 3).clone()
    pose[0, 19] = torch.as_tensor(syn.random_rotmats(rs, ()))     # body joint 19 = SMPL joint 20, child 22
    out = ORACLE.forward(torch.zeros(1, 10), pose, EYE.expand(1, 1, 3, 3))
    # remove the pose-corrective offsets to isolate skinning
    off = out["v_posed"][0] - out["v_shaped"][0]
    w = torch.as_tensor(MODEL["lbs_weights"])
    affected = (w[:, 20] + w[:, 22]) > 0
    moved = ((out["vertices"][0] - off) - torch.as_tensor(MODEL["v_template"])).abs().max(1).values > 1e-9
    assert not moved[~affected].any()


def test_fp32_restatement_close_to_fp64():
    rs = np.random.RandomState(2)
    M = 4
    pose = torch.as_tensor(syn.random_rotmats(rs, (M, 23)))
    glob = torch.as_tensor(syn.random_rotmats(rs, (M, 1)))
    betas = torch.as_tensor(rs.normal(0, 1.25, size=(M, 10)))
    o64 = ORACLE.forward(betas, pose, glob)
    o32 = SMPLOracle(MODEL, torch.float32).forward(betas.float(), pose.float(), glob.float())
    assert ((o32["vertices"].double() - o64["vertices"]).abs().max() / o64["vertices"].abs().max()) < 1e-5


def test_rodrigues_matches_matrix_exponential():
    rs = np.random.RandomState(3)
    r = torch.as_tensor(rs.normal(size=(16, 3)))
    R = SMPLOracle.batch_rodrigues(r)
    K = torch.zeros(16, 3, 3, dtype=torch.float64)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -r[:, 2], r[:, 1], r[:, 2], -r[:, 0], -r[:, 1], r[:, 0]
    assert (R - torch.matrix_exp(K)).abs().max() < 1e-6
