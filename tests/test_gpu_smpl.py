"""CUDA SMPL forward (through the SMPL drop-in -> C ABI) vs the fp64 oracle restatement."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle.smpl_oracle import SMPLOracle
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 1e-4   # north_star: 1e-4 relative fp32 (max|d| / max|ref|)


@pytest.fixture(scope="module")
def setup(built_lib):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    return hp.SMPL(model=model).cuda(), SMPLOracle(model, torch.float64), model


def _rand(rs, M, Mb=None):
    pose = torch.as_tensor(syn.random_rotmats(rs, (M, 23)), dtype=torch.float32)
    glob = torch.as_tensor(syn.random_rotmats(rs, (M, 1)), dtype=torch.float32)
    betas = torch.as_tensor(rs.normal(0, 1.25, size=(Mb or M, 10)), dtype=torch.float32)
    return betas, pose, glob


@pytest.mark.parametrize("M", [1, 2, 5, 32, 257])
def test_vertices_and_joints_match_oracle(setup, M):
    smpl, oracle, _ = setup
    betas, pose, glob = _rand(np.random.RandomState(M), M)
    out = smpl(betas=betas.cuda(), body_pose=pose.cuda(), global_orient=glob.cuda(), pose2rot=False)
    ref = oracle.forward(betas, pose, glob)
    assert out.vertices.shape == (M, 6890, 3) and out.joints.shape == (M, 90, 3)
    assert rel_err(out.vertices, ref["vertices"]) < TOL
    assert rel_err(out.joints, ref["joints"]) < TOL


def test_stage_outputs_match_oracle(setup):
    """shape blend / pose blend / J through the stage-level C entry points."""
    import ctypes
    from hierarchicalprobabilistic3dhuman_b200 import _lib
    smpl, oracle, _ = setup
    M = 6
    betas, pose, glob = _rand(np.random.RandomState(11), M)
    ref = oracle.forward(betas, pose, glob)
    L = _lib.lib()
    h = smpl._handle(torch.device("cuda", 0))
    vs = torch.empty(M, 20672, device="cuda"); J = torch.empty(M, 24, 3, device="cuda"); vp = torch.empty(M, 20672, device="cuda")
    b, p = betas.cuda(), pose.cuda().contiguous()
    _lib.check(L.hp3d_smpl_shape_blend(h, b.data_ptr(), M, vs.data_ptr(), J.data_ptr(), None))
    wsb = torch.empty(L.hp3d_smpl_pose_blend_workspace_bytes(M), dtype=torch.uint8, device="cuda")
    _lib.check(L.hp3d_smpl_pose_blend(h, b.data_ptr(), vs.data_ptr(), M, p.data_ptr(), M, vp.data_ptr(), wsb.data_ptr(), wsb.numel(), None))
    torch.cuda.synchronize()
    assert rel_err(vs[:, :20670].reshape(M, 6890, 3), ref["v_shaped"]) < 1e-6
    assert rel_err(J, ref["J"]) < 1e-5
    assert rel_err(vp[:, :20670].reshape(M, 6890, 3), ref["v_posed"]) < 1e-5


def test_known_answers(setup):
    smpl, _, model = setup
    eye = torch.eye(3, device="cuda")
    out = smpl(betas=torch.zeros(1, 10, device="cuda"), body_pose=eye.expand(1, 23, 3, 3), global_orient=eye.expand(1, 1, 3, 3), pose2rot=False)
    vt = torch.as_tensor(model["v_template"], dtype=torch.float32)
    assert rel_err(out.vertices[0], vt) < 1e-6
    assert rel_err(out.joints[0, :24], torch.as_tensor(model["J_regressor"] @ model["v_template"])) < 1e-5
    assert rel_err(out.joints[0, 24:45], vt[model["extra_vertex_ids"]]) < 1e-6
    # default call: stored zero axis-angle parameters (reference predict/...:136)
    out2 = smpl(betas=torch.zeros(1, 10, device="cuda"))
    assert rel_err(out2.vertices, out.vertices) < 1e-5


def test_axis_angle_mode_matches_oracle(setup):
    smpl, oracle, _ = setup
    rs = np.random.RandomState(5)
    M = 3
    bp = torch.as_tensor(rs.normal(0, 0.4, size=(M, 69)), dtype=torch.float32)
    go = torch.as_tensor(rs.normal(0, 1.0, size=(M, 3)), dtype=torch.float32)
    betas = torch.as_tensor(rs.normal(size=(M, 10)), dtype=torch.float32)
    out = smpl(betas=betas.cuda(), body_pose=bp.cuda(), global_orient=go.cuda())      # pose2rot=True default
    ref = oracle.forward(betas, bp, go, pose2rot=True)
    assert rel_err(out.vertices, ref["vertices"]) < TOL and rel_err(out.joints, ref["joints"]) < TOL


def test_per_image_broadcast_of_betas_and_global_orient(setup):
    """B images x N samples with one betas / global_orient row per image (utils/sampling_utils.py:178-185)."""
    smpl, oracle, _ = setup
    rs = np.random.RandomState(6)
    B, N = 3, 4
    betas, pose, _ = _rand(rs, B * N, Mb=B)
    glob = torch.as_tensor(syn.random_rotmats(rs, (B, 1)), dtype=torch.float32)
    out = smpl(betas=betas.cuda(), body_pose=pose.cuda(), global_orient=glob.cuda(), pose2rot=False)
    ref = oracle.forward(betas.repeat_interleave(N, 0), pose, glob.repeat_interleave(N, 0))
    assert rel_err(out.vertices, ref["vertices"]) < TOL and rel_err(out.joints, ref["joints"]) < TOL


def test_output_into_offset_slice(setup):
    """Outputs may live inside a larger (all-gather) buffer: odd mesh offsets are only 8-byte aligned."""
    import ctypes
    from hierarchicalprobabilistic3dhuman_b200 import _lib
    smpl, oracle, _ = setup
    M = 3
    betas, pose, glob = _rand(np.random.RandomState(8), M)
    big = torch.zeros(M + 2, 6890, 3, device="cuda")
    L = _lib.lib(); h = smpl._handle(torch.device("cuda", 0))
    ws = torch.empty(L.hp3d_smpl_workspace_bytes(h, M, M), dtype=torch.uint8, device="cuda")
    b, p, g = betas.cuda(), pose.cuda().contiguous(), glob.cuda().contiguous()
    _lib.check(L.hp3d_smpl_forward(h, b.data_ptr(), M, g.data_ptr(), M, p.data_ptr(), M, big[1:].data_ptr(), None,
                                   ws.data_ptr(), ws.numel(), None))
    torch.cuda.synchronize()
    ref = oracle.forward(betas, pose, glob)
    assert rel_err(big[1:1 + M], ref["vertices"]) < TOL
    assert big[0].abs().max() == 0 and big[M + 1].abs().max() == 0


def test_vertex_uncertainty(setup):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    rs = np.random.RandomState(9)
    v = torch.as_tensor(rs.normal(size=(2, 7, 6890, 3)), dtype=torch.float32)
    mean, dist = hp.vertex_uncertainty(v.cuda())
    m = v.double().mean(1)
    d = (v.double() - m[:, None]).norm(dim=-1).mean(1)
    assert rel_err(mean, m) < 1e-5 and rel_err(dist, d) < 1e-5


def test_rot6d(setup):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from oracle import net_oracle
    x = torch.randn(17, 6)
    assert rel_err(hp.rot6d_to_rotmat(x.cuda()), net_oracle.rot6d_to_rotmat(x)) < 1e-5


def test_errors_are_loud(setup):
    smpl, _, _ = setup
    with pytest.raises(RuntimeError):
        smpl(betas=torch.zeros(1, 10), body_pose=torch.eye(3).expand(1, 23, 3, 3), global_orient=torch.eye(3).expand(1, 1, 3, 3), pose2rot=False)
    with pytest.raises(ValueError):
        smpl(betas=torch.zeros(2, 10).cuda(), body_pose=torch.eye(3).expand(3, 23, 3, 3).cuda(), global_orient=torch.eye(3).expand(3, 1, 3, 3).cuda(), pose2rot=False)


def _tile_nq_max(model):
    W = model["lbs_weights"]
    return max(int((W[t * 64:(t + 1) * 64] != 0).any(0).sum()) for t in range(108))


@pytest.mark.parametrize("variant", ["tile8", "tile12", "generic"])
def test_all_skinning_kernel_variants(built_lib, variant):
    """The tile-local LBS kernel (<=8 / <=12 joints per 64-vertex tile) and the generic per-vertex gather
    kernel (models without spatial coherence) must agree with the oracle."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    rs = np.random.RandomState(3)
    W = model["lbs_weights"].copy()
    if variant == "tile12":          # shuffle rows inside windows of 420 vertices: tiles see more joints
        for s in range(0, 6890, 420):
            idx = np.arange(s, min(s + 420, 6890)); W[idx] = W[rs.permutation(idx)]
    elif variant == "generic":       # no coherence at all
        W = W[rs.permutation(6890)]
    model["lbs_weights"] = W
    nq = _tile_nq_max(model)
    assert {"tile8": nq <= 8, "tile12": 8 < nq <= 12, "generic": nq > 12}[variant], nq
    smpl = hp.SMPL(model=model).cuda()
    oracle = SMPLOracle(model, torch.float64)
    for M in (3, 17):
        betas, pose, glob = _rand(np.random.RandomState(40 + M), M)
        out = smpl(betas=betas.cuda(), body_pose=pose.cuda(), global_orient=glob.cuda(), pose2rot=False)
        ref = oracle.forward(betas, pose, glob)
        assert rel_err(out.vertices, ref["vertices"]) < TOL and rel_err(out.joints, ref["joints"]) < TOL


def test_full_size_smpl_properties(built_lib):
    """BASELINE configs[1] size (64 x 100 meshes): linearity in the global rotation and finiteness."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    smpl = hp.SMPL(model=model).cuda()
    rs = np.random.RandomState(12)
    B, N = 64, 100
    pose = torch.as_tensor(syn.random_rotmats(rs, (B * N, 23)), dtype=torch.float32).cuda()
    glob = torch.as_tensor(syn.random_rotmats(rs, (B, 1)), dtype=torch.float32).cuda()
    betas = torch.as_tensor(rs.normal(0, 1.25, size=(B, 10)), dtype=torch.float32).cuda()
    out = smpl(betas=betas, body_pose=pose, global_orient=glob, pose2rot=False)
    eye = torch.eye(3, device="cuda").expand(B, 1, 3, 3).contiguous()
    base = smpl(betas=betas, body_pose=pose, global_orient=eye, pose2rot=False)
    assert torch.isfinite(out.vertices).all()
    # global orientation acts rigidly about the (shape-dependent) root joint
    root = base.joints[:, 0:1]                                         # (M,1,3) posed root == J_0
    Rg = glob[:, 0].repeat_interleave(N, 0)
    expect = torch.einsum("mij,mvj->mvi", Rg, base.vertices - root) + root
    assert rel_err(out.vertices, expect) < 1e-5
    # spot-check 3 meshes against the oracle
    oracle = SMPLOracle(model, torch.float64)
    idx = [0, 3217, B * N - 1]
    ref = oracle.forward(betas.cpu()[[i // N for i in idx]], pose.cpu()[idx], glob.cpu()[[i // N for i in idx]])
    assert rel_err(out.vertices[idx], ref["vertices"]) < TOL and rel_err(out.joints[idx], ref["joints"]) < TOL
