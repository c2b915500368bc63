"""CPU checks of host-side helpers added in round 2: the vertex-shuffled SMPL model used for the vertex-order robustness
measurements (must be the SAME model up to a permutation of its vertices), and bench.py's lookup of a kernel's measured
DRAM traffic in the committed ncu summaries."""
import numpy as np
import torch

from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
from oracle.smpl_oracle import SMPLOracle


def test_shuffled_model_is_a_vertex_permutation_of_the_original():
    model = syn.synthetic_smpl_model()
    for block in (1, 16):
        shuf = syn.shuffle_smpl_vertices(model, seed=3, block=block)
        # recover the permutation from the template (rows are distinct)
        key = {tuple(np.round(v, 12)): i for i, v in enumerate(model["v_template"])}
        order = np.array([key[tuple(np.round(v, 12))] for v in shuf["v_template"]])
        assert sorted(order.tolist()) == list(range(6890))
        if block == 16:      # blocks of 16 consecutive vertices stay together
            assert (np.diff(order).reshape(-1)[:15] == 1).all()
        rs = np.random.RandomState(0)
        M = 3
        betas = torch.from_numpy(rs.normal(size=(M, 10)))
        R = torch.from_numpy(syn.random_rotmats(rs, (M, 23)))
        gR = torch.from_numpy(syn.random_rotmats(rs, (M, 1)))
        a = SMPLOracle(model, torch.float64).forward(betas, R, gR)
        b = SMPLOracle(shuf, torch.float64).forward(betas, R, gR)
        assert torch.allclose(b["vertices"], a["vertices"][:, order], atol=1e-12)
        assert torch.allclose(b["joints"][:, :24], a["joints"][:, :24], atol=1e-12)      # skeleton joints do not depend on the order
        assert np.array_equal(shuf["faces"], np.argsort(order)[model["faces"]])


def test_bench_reads_kernel_traffic_from_the_committed_ncu_summaries():
    import bench
    fused, src_f = bench.traffic_from_profiles("smpl_fused_kernel", 148)
    lbs, src_l = bench.traffic_from_profiles("lbs_tile_kernel", 3200)
    # one launch on 25,600 meshes: fused ~2.2 GB algorithmic (+ operand misses), staged LBS 4.29 GB = its algorithmic bytes
    assert src_f and 2.1e9 < fused < 3.0e9, (fused, src_f)
    assert src_l and abs(lbs - 167592 * 25600) / (167592 * 25600) < 0.01, (lbs, src_l)
    assert bench.traffic_from_profiles("no_such_kernel", 1) == (None, None)
