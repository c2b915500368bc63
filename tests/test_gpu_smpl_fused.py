"""The fused SMPL kernel (csrc/smpl_fused.cu: transposed blend GEMM -> skinning -> per-vertex statistics in ONE tensor-core
kernel, v_posed never in HBM; the default path) against the fp64 SMPL oracle and the staged three-kernel path
(HP3D_SMPL=staged), incl. models whose vertex order is shuffled and the forced create-time re-ordering (scattered stores)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
from oracle.smpl_oracle import SMPLOracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture
def fused_env():
    old = {k: os.environ.get(k) for k in ("HP3D_SMPL", "HP3D_SMPL_ORDER", "HP3D_SMPL_CLUSTER")}
    os.environ["HP3D_SMPL"] = "fused"
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _inputs(M, Mb, Mg, seed):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    torch.manual_seed(seed)
    R = hp.rot6d_to_rotmat(torch.randn(M * 23, 6, device="cuda")).view(M, 23, 3, 3)
    gR = hp.rot6d_to_rotmat(torch.randn(Mg, 6, device="cuda"))
    betas = torch.randn(Mb, 10, device="cuda") * 1.25
    return R, gR, betas


def _layout(smpl):
    import ctypes
    from hierarchicalprobabilistic3dhuman_b200 import _lib
    v = [ctypes.c_int() for _ in range(4)]
    _lib.check(_lib.lib().hp3d_smpl_layout_info(smpl._handle(torch.device("cuda", 0)), *[ctypes.byref(x) for x in v]))
    return dict(fused=v[0].value, permuted=v[1].value, nq_sum=v[2].value, nq_max=v[3].value)


@pytest.mark.parametrize("M,Mb,Mg", [(21, 3, 3), (500, 5, 5), (256, 256, 256), (113, 1, 113)])
def test_fused_forward_matches_oracle(built_lib, fused_env, M, Mb, Mg):
    import hierarchicalprobabilistic3dhuman_b200 as hp
    model = syn.synthetic_smpl_model()
    smpl = hp.SMPL(model=model).cuda()
    assert _layout(smpl)["fused"] == 1
    R, gR, betas = _inputs(M, Mb, Mg, M)
    out = smpl(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
    ref = SMPLOracle(model, torch.float64).forward(betas.cpu().repeat_interleave(M // Mb, 0), R.cpu(),
                                                   gR.cpu().repeat_interleave(M // Mg, 0)[:, None])
    assert rel_err(out.vertices, ref["vertices"]) < TOL and rel_err(out.joints, ref["joints"]) < TOL
    # and against the staged path (same arithmetic up to the association of the blend sum)
    os.environ["HP3D_SMPL"] = "staged"
    st = smpl(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
    assert rel_err(out.vertices, st.vertices) < 1e-5 and rel_err(out.joints, st.joints) < 1e-5


@pytest.mark.parametrize("B,N,cluster", [(3, 100, "0"), (5, 8, "0"), (2, 112, "0"), (2, 17, "0"), (3, 5, "0"), (3, 100, "1"), (4, 33, "1")])
def test_fused_statistics_match_oracle(built_lib, fused_env, B, N, cluster):
    """per-vertex mean distance to the mean mesh (utils/sampling_utils.py:189-190) out of the SMPL kernel itself"""
    import ctypes
    import hierarchicalprobabilistic3dhuman_b200 as hp
    from hierarchicalprobabilistic3dhuman_b200 import _lib
    os.environ["HP3D_SMPL_CLUSTER"] = cluster          # "1": 2-CTA clusters, posedirs tiles multicast (odd chunk count: B = 3)
    model = syn.synthetic_smpl_model()
    smpl = hp.SMPL(model=model).cuda()
    M = B * N
    R, gR, betas = _inputs(M, B, B, 100 + M)
    dev = torch.device("cuda", 0)
    L, h = _lib.lib(), smpl._handle(dev)
    verts = torch.empty(M, 6890, 3, device=dev); joints = torch.empty(M, 90, 3, device=dev)
    unc = torch.empty(B, 6890, device=dev); mean = torch.empty(B, 6890, 3, device=dev)
    ws = torch.empty(L.hp3d_smpl_workspace_bytes(h, M, B), dtype=torch.uint8, device=dev)
    _lib.check(L.hp3d_smpl_forward_stats(h, betas.data_ptr(), B, gR.data_ptr(), B, R.data_ptr(), M, N, verts.data_ptr(),
                                         joints.data_ptr(), unc.data_ptr(), mean.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _lib.stream_ptr()), "hp3d_smpl_forward_stats")
    ref = SMPLOracle(model, torch.float64).forward(betas.cpu().repeat_interleave(N, 0), R.cpu(), gR.cpu().repeat_interleave(N, 0)[:, None])
    v = ref["vertices"].view(B, N, 6890, 3)
    m_ref = v.mean(1)
    u_ref = (v - m_ref[:, None]).norm(dim=-1).mean(1)
    assert rel_err(verts, ref["vertices"]) < TOL and rel_err(joints, ref["joints"]) < TOL
    assert rel_err(mean, m_ref) < TOL and rel_err(unc, u_ref) < TOL


@pytest.mark.parametrize("block,order", [(1, None), (16, None), (1, "sorted"), (64, "sorted")])
def test_fused_on_shuffled_vertex_orders(built_lib, fused_env, block, order):
    """Real SMPL is not ordered by body part. Skinning runs as a dense K = 24 tensor-core contraction, so a shuffled vertex
    order (tiles touching up to 22 joints) takes the same code path as the part-ordered model; HP3D_SMPL_ORDER=sorted forces
    the create-time re-ordering and with it the scattered-store path."""
    import hierarchicalprobabilistic3dhuman_b200 as hp
    if order:
        os.environ["HP3D_SMPL_ORDER"] = order
    model = syn.shuffle_smpl_vertices(syn.synthetic_smpl_model(), seed=3, block=block)
    smpl = hp.SMPL(model=model).cuda()
    info = _layout(smpl)
    assert info["permuted"] == (1 if order == "sorted" else 0), info
    if order is None and block == 1:
        assert info["nq_max"] > 12, info
    M = 37
    R, gR, betas = _inputs(M, M, M, 9)
    out = smpl(body_pose=R, global_orient=gR.unsqueeze(1), betas=betas, pose2rot=False)
    ref = SMPLOracle(model, torch.float64).forward(betas.cpu(), R.cpu(), gR.cpu()[:, None])
    assert rel_err(out.vertices, ref["vertices"]) < TOL and rel_err(out.joints, ref["joints"]) < TOL
