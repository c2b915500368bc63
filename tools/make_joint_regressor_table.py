"""Pack the reference's three extra joint regressors into one sparse table.

The reference registers three dense (J, 6890) float64 matrices as fp32 buffers
(reference models/smpl_official.py:17-25, files named in configs/paths.py:3-5).
They hold 255 non-zeros in total, so the build ships them as COO triplets
(row, vertex, weight) in joint order extra(9) | cocoplus(19) | h36m(17) = 45 rows.
This is model *data*, not source; /root/reference does not exist on the GPU box.

Run here (needs /root/reference):  python tools/make_joint_regressor_table.py
"""
import numpy as np, os, sys
REF = os.environ.get("HP3D_REFERENCE", "/root/reference")
names = ["J_regressor_extra.npy", "cocoplus_regressor.npy", "J_regressor_h36m.npy"]
rows, cols, vals, counts = [], [], [], []
base = 0
for n in names:
    a = np.load(os.path.join(REF, "model_files", n))
    assert a.shape[1] == 6890
    r, c = np.nonzero(a)
    rows.append(r + base); cols.append(c); vals.append(a[r, c])
    counts.append(a.shape[0]); base += a.shape[0]
out = os.path.join(os.path.dirname(__file__), "..", "hierarchicalprobabilistic3dhuman_b200", "data", "joint_regressors.npz")
np.savez_compressed(out, rows=np.concatenate(rows).astype(np.int32), cols=np.concatenate(cols).astype(np.int32),
                    vals=np.concatenate(vals).astype(np.float64), counts=np.array(counts, np.int32))
print("wrote", out, "nnz", sum(len(v) for v in vals), "rows", base)
