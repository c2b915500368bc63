"""Seeded synthetic stand-ins for the assets the reference needs but cannot ship.

* SMPL model file: licence-gated (reference README.md:46) -> `synthetic_smpl_model`
  builds an SMPL-*shaped* model (6890 vertices, 24 joints, 10 betas, 207 pose
  features, true kinematic tree, <=4 skinning weights per vertex).
* Trained checkpoint: not available -> `randomise_bn_stats` perturbs BatchNorm
  running statistics of a default-initialised network so BN folding is exercised.
* Proxy representation input (edge map + 17 joint heatmaps, reference
  predict/predict_poseMF_shapeGaussian_net.py:91-100) -> `synthetic_proxy_rep`.

Everything here is deterministic numpy (`RandomState`) so the GPU box regenerates
bit-identical inputs from seeds; only small reference *outputs* are committed as
golden fixtures (tests/golden/).
"""
import os
import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_FEATS = 207
NUM_FACES = 13776

# smplx 0.1.26 SMPL kinematic tree
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14,
                         16, 17, 18, 19, 20, 21], dtype=np.int64)

# smplx vertex_ids['smplh'] order used by VertexJointSelector for SMPL (SURVEY.md §8c step 8;
# recalled from smplx 0.1.26 -- data, not logic; replaceable through the model dict).
SMPL_EXTRA_VERTEX_IDS = np.array(
    [332, 6260, 2800, 4071, 583,            # nose, reye, leye, rear, lear
     3216, 3226, 3387, 6617, 6624, 6787,    # LBigToe LSmallToe LHeel RBigToe RSmallToe RHeel
     2746, 2319, 2445, 2556, 2673,          # l thumb index middle ring pinky
     6191, 5782, 5905, 6016, 6133],         # r thumb index middle ring pinky
    dtype=np.int32)


def load_joint_regressor_table():
    """(45, 6890) float64 dense matrix: extra(9) | cocoplus(19) | h36m(17) rows
    (reference models/smpl_official.py:17-25,30-34); packed by tools/make_joint_regressor_table.py."""
    p = os.path.join(os.path.dirname(__file__), "data", "joint_regressors.npz")
    z = np.load(p)
    dense = np.zeros((int(z["counts"].sum()), NUM_VERTS), dtype=np.float64)
    dense[z["rows"], z["cols"]] = z["vals"]
    return dense


def shuffle_smpl_vertices(model, seed=7, block=1):
    """The same model with its vertex order permuted (blocks of `block` consecutive vertices stay together): the real SMPL
    template is NOT ordered by body part the way `synthetic_smpl_model` is, so kernels whose speed depends on the vertex
    order are also measured / tested on a shuffled copy (block=1: worst case, every vertex on its own)."""
    rs = np.random.RandomState(seed)
    nb = -(-NUM_VERTS // block)
    order = np.concatenate([np.arange(b * block, min(NUM_VERTS, (b + 1) * block)) for b in rs.permutation(nb)])   # new -> old
    inv = np.empty(NUM_VERTS, dtype=np.int64)
    inv[order] = np.arange(NUM_VERTS)                                                                          # old -> new
    m = dict(model)
    m["v_template"] = model["v_template"][order]
    m["shapedirs"] = model["shapedirs"][order]
    m["posedirs"] = model["posedirs"].reshape(NUM_POSE_FEATS, NUM_VERTS, 3)[:, order].reshape(NUM_POSE_FEATS, NUM_VERTS * 3)
    m["J_regressor"] = model["J_regressor"][:, order]
    m["lbs_weights"] = model["lbs_weights"][order]
    m["faces"] = inv[model["faces"]]
    # extra_vertex_ids / joint_regressors_extra keep addressing vertex INDICES (they are data about the real template)
    return m


def synthetic_smpl_model(seed=2, max_skin_nnz=4):
    """SMPL-shaped model constants (float64 numpy) with the layouts smplx 0.1.26 uses
    (SURVEY.md §8a a11): v_template (6890,3), shapedirs (6890,3,10), posedirs (207,20670),
    J_regressor (24,6890), lbs_weights (6890,24), parents (24,), faces (13776,3)."""
    rs = np.random.RandomState(seed)
    v_template = rs.normal(0.0, 0.3, size=(NUM_VERTS, 3))
    shapedirs = rs.normal(0.0, 0.01, size=(NUM_VERTS, 3, NUM_BETAS))
    posedirs = rs.normal(0.0, 0.001, size=(NUM_POSE_FEATS, NUM_VERTS * 3))
    J_regressor = np.zeros((NUM_JOINTS, NUM_VERTS))
    for j in range(NUM_JOINTS):
        idx = rs.choice(NUM_VERTS, size=30, replace=False)
        w = rs.uniform(0.1, 1.0, size=30)
        J_regressor[j, idx] = w / w.sum()
    # Skinning weights: like real SMPL, vertex index ranges belong to one body part and
    # blend with the part's neighbours in the tree (parent / children) plus one random joint.
    children = [[c for c in range(NUM_JOINTS) if SMPL_PARENTS[c] == j] for j in range(NUM_JOINTS)]
    lbs_weights = np.zeros((NUM_VERTS, NUM_JOINTS))
    primary = (np.arange(NUM_VERTS) * NUM_JOINTS) // NUM_VERTS
    for v in range(NUM_VERTS):
        j = int(primary[v])
        cand = [j]
        if SMPL_PARENTS[j] >= 0:
            cand.append(int(SMPL_PARENTS[j]))
        cand += children[j]
        # fill up with the next-nearest joints in the tree (grandparent, siblings, grandchildren), like the
        # smooth, spatially coherent weights of the real model -- never with an unrelated random joint
        ring = []
        if SMPL_PARENTS[j] >= 0:
            pj = int(SMPL_PARENTS[j])
            if SMPL_PARENTS[pj] >= 0:
                ring.append(int(SMPL_PARENTS[pj]))
            ring += [c for c in children[pj] if c != j]
        for c in children[j]:
            ring += children[c]
        ring += [(j + d) % NUM_JOINTS for d in (1, 2, 3)]
        for r in ring:
            if len(cand) >= max_skin_nnz:
                break
            if r not in cand:
                cand.append(r)
        cand = cand[:max_skin_nnz]
        w = rs.uniform(0.05, 1.0, size=len(cand))
        w[0] += 1.0
        lbs_weights[v, cand] = w / w.sum()
    faces = rs.randint(0, NUM_VERTS, size=(NUM_FACES, 3)).astype(np.int64)
    return dict(v_template=v_template, shapedirs=shapedirs, posedirs=posedirs,
                J_regressor=J_regressor, lbs_weights=lbs_weights,
                parents=SMPL_PARENTS.copy(), faces=faces,
                extra_vertex_ids=SMPL_EXTRA_VERTEX_IDS.copy(),
                joint_regressors_extra=load_joint_regressor_table())


def random_rotmats(rs, shape):
    """Random proper rotations via QR of Gaussian matrices (float64)."""
    a = rs.normal(size=tuple(shape) + (3, 3))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diagonal(r, axis1=-2, axis2=-1))[..., None, :]
    det = np.linalg.det(q)
    q[..., :, 2] *= det[..., None]
    return q


def synthetic_usv(batch, joints=23, seed=1, s_lo=1e-2, s_hi=5e2):
    """Direct sampler inputs: U, V random O(3) (improper factors occur),
    S descending, log-uniform in [s_lo, s_hi]. float32 numpy."""
    rs = np.random.RandomState(seed)
    def orth(n):
        a = rs.normal(size=(n, 3, 3))
        q, _ = np.linalg.qr(a)
        return q
    U = orth(batch * joints).reshape(batch, joints, 3, 3)
    V = orth(batch * joints).reshape(batch, joints, 3, 3)
    S = np.exp(rs.uniform(np.log(s_lo), np.log(s_hi), size=(batch, joints, 3)))
    S = -np.sort(-S, axis=-1)
    return U.astype(np.float32), S.astype(np.float32), V.astype(np.float32)


def synthetic_proxy_rep(batch, seed=0, size=256, std=4.0):
    """(B,18,size,size) float32: channel 0 a sparse non-negative thin-edge map, channels 1-17
    Gaussian joint heatmaps with sigma=std (reference utils/label_conversions.py:105-124 indexing:
    `exp(-((xx - v)/std)^2/2 - ((yy - u)/std)^2/2)` with xx the ROW index under torch.meshgrid 'ij'),
    joints {7,8,9,10,13,14,15,16} zeroed w.p. 0.1 (reference predict/...:97-99)."""
    rs = np.random.RandomState(seed)
    out = np.zeros((batch, 18, size, size), dtype=np.float32)
    rows = np.arange(size, dtype=np.float32)[:, None]
    cols = np.arange(size, dtype=np.float32)[None, :]
    for b in range(batch):
        # edge channel: zero level-set band of a smooth random field, gradient-magnitude valued
        k = 6
        coarse = rs.normal(size=(size // 16 + k, size // 16 + k)).astype(np.float32)
        fy = np.linspace(0, coarse.shape[0] - 1.001, size)
        fx = np.linspace(0, coarse.shape[1] - 1.001, size)
        y0 = fy.astype(int); x0 = fx.astype(int)
        wy = (fy - y0)[:, None].astype(np.float32); wx = (fx - x0)[None, :].astype(np.float32)
        f = ((1 - wy) * (1 - wx) * coarse[np.ix_(y0, x0)] + (1 - wy) * wx * coarse[np.ix_(y0, x0 + 1)]
             + wy * (1 - wx) * coarse[np.ix_(y0 + 1, x0)] + wy * wx * coarse[np.ix_(y0 + 1, x0 + 1)])
        gy, gx = np.gradient(f)
        mag = np.sqrt(gx * gx + gy * gy)
        band = np.abs(f) < 0.6 * mag
        out[b, 0] = np.where(band, mag * 4.0, 0.0)
        j2d = rs.uniform(32.0, size - 32.0, size=(17, 2)).astype(np.float32)
        vis = np.ones(17, dtype=bool)
        for j in (7, 8, 9, 10, 13, 14, 15, 16):
            vis[j] = rs.uniform() >= 0.1
        for j in range(17):
            if vis[j]:
                u, v = j2d[j]
                out[b, 1 + j] = np.exp(-(((rows - v) / std) ** 2) / 2 - (((cols - u) / std) ** 2) / 2)
    return out


def synthetic_images(batch, seed=0, size=256):
    """Image-space inputs of the proxy-representation generator (reference predict/...:84-99): a (B,3,size,size)
    float32 RGB crop in [0,1] (smooth random colour field + mild pixel noise, so Canny finds real edge structure),
    (B,17,2) float32 2D joints as (u, v) -- a few of them outside the image --, and the (B,17) bool visibility mask
    (joints {7,8,9,10,13,14,15,16} dropped w.p. 0.1, predict/...:97-98)."""
    rs = np.random.RandomState(seed)
    rgb = np.zeros((batch, 3, size, size), dtype=np.float32)
    fy = np.linspace(0, 18.999, size)
    y0 = fy.astype(int)
    wy = (fy - y0).astype(np.float32)
    for b in range(batch):
        coarse = rs.uniform(0, 1, size=(3, 21, 21)).astype(np.float32)
        f = ((1 - wy)[None, :, None] * (1 - wy)[None, None, :] * coarse[:, y0][:, :, y0]
             + (1 - wy)[None, :, None] * wy[None, None, :] * coarse[:, y0][:, :, y0 + 1]
             + wy[None, :, None] * (1 - wy)[None, None, :] * coarse[:, y0 + 1][:, :, y0]
             + wy[None, :, None] * wy[None, None, :] * coarse[:, y0 + 1][:, :, y0 + 1])
        f = f + rs.normal(0, 0.02, size=f.shape).astype(np.float32)
        rgb[b] = np.clip(f, 0.0, 1.0)
    j2d = rs.uniform(-8.0, size + 8.0, size=(batch, 17, 2)).astype(np.float32)
    vis = np.ones((batch, 17), dtype=bool)
    for b in range(batch):
        for j in (7, 8, 9, 10, 13, 14, 15, 16):
            vis[b, j] = rs.uniform() >= 0.1
    return rgb, j2d, vis


def synthetic_crop_inputs(batch, seed=0, height=384, width=288):
    """Inputs of the crop / resample step in front of the proxy representation (reference predict/...:73-93,
    SURVEY.md §8f rank 3): an HRNet-sized RGB image (B,3,height,width) in [0,1], (B,17,2) joints in it, bounding-box
    centres (B,2) as (vertical, horizontal), heights (B,), widths (B,) -- the first box is the whole image like the
    predict path's, the others are arbitrary (also partly outside the image)."""
    rs = np.random.RandomState(seed)
    rgb = rs.uniform(0, 1, size=(batch, 3, height // 8, width // 8)).astype(np.float32)
    rgb = np.repeat(np.repeat(rgb, 8, axis=2), 8, axis=3)
    rgb = np.clip(rgb + rs.normal(0, 0.05, size=rgb.shape).astype(np.float32), 0.0, 1.0)
    j2d = (rs.uniform(0, 1, size=(batch, 17, 2)) * np.array([width, height])).astype(np.float32)
    centres = np.stack([rs.uniform(0.3 * height, 0.7 * height, batch), rs.uniform(0.3 * width, 0.7 * width, batch)], axis=1)
    heights = rs.uniform(0.3 * height, 1.1 * height, batch)
    widths = rs.uniform(0.3 * width, 1.1 * width, batch)
    centres[0] = (height * 0.5, width * 0.5)
    heights[0] = height
    widths[0] = height
    return rgb, j2d, centres.astype(np.float32), heights.astype(np.float32), widths.astype(np.float32)


def randomise_bn_stats(module, seed=0):
    """running_mean~N(0,0.1), running_var~U(0.5,1.5), gamma~U(0.5,1.5), beta~N(0,0.1) on every
    BatchNorm2d of a torch module, from a numpy stream (torch-version independent)."""
    import torch
    rs = np.random.RandomState(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            c = m.num_features
            with torch.no_grad():
                m.running_mean.copy_(torch.from_numpy(rs.normal(0, 0.1, c).astype(np.float32)))
                m.running_var.copy_(torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32)))
                m.weight.copy_(torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32)))
                m.bias.copy_(torch.from_numpy(rs.normal(0, 0.1, c).astype(np.float32)))


def synthetic_state_dict(seed=0, parents=None):
    """A full `PoseMFShapeGaussianNet` state_dict with the reference's parameter names and shapes
    (SURVEY.md §5 'checkpoint/resume' row; reference models/poseMF_shapeGaussian_net.py:25-83,
    models/resnet.py:146-157), filled from a numpy stream: conv ~ kaiming-normal(fan_out),
    linear ~ U(+-1/sqrt(fan_in)), BN as in `randomise_bn_stats`. float32 torch tensors."""
    import torch
    rs = np.random.RandomState(seed)
    parents = SMPL_PARENTS if parents is None else np.asarray(parents)
    sd = {}

    def conv(name, cout, cin, k):
        std = np.sqrt(2.0 / (cout * k * k))
        sd[name + ".weight"] = rs.normal(0, std, size=(cout, cin, k, k)).astype(np.float32)

    def bn(name, c):
        sd[name + ".weight"] = rs.uniform(0.5, 1.5, c).astype(np.float32)
        sd[name + ".bias"] = rs.normal(0, 0.1, c).astype(np.float32)
        sd[name + ".running_mean"] = rs.normal(0, 0.1, c).astype(np.float32)
        sd[name + ".running_var"] = rs.uniform(0.5, 1.5, c).astype(np.float32)
        sd[name + ".num_batches_tracked"] = np.array(0, dtype=np.int64)

    def lin(name, cout, cin):
        b = 1.0 / np.sqrt(cin)
        sd[name + ".weight"] = rs.uniform(-b, b, size=(cout, cin)).astype(np.float32)
        sd[name + ".bias"] = rs.uniform(-b, b, size=(cout,)).astype(np.float32)

    e = "image_encoder."
    conv(e + "conv1", 64, 18, 7); bn(e + "bn1", 64)
    inpl = 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            p = f"{e}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            conv(p + "conv1", planes, inpl, 3); bn(p + "bn1", planes)
            conv(p + "conv2", planes, planes, 3); bn(p + "bn2", planes)
            if stride != 1 or inpl != planes:
                conv(p + "downsample.0", planes, inpl, 1); bn(p + "downsample.1", planes)
            inpl = planes
    sd["init_glob"] = np.array([[1, 0, 0, 1, 0, 0]], dtype=np.float32)
    sd["init_cam"] = np.array([0.9, 0.0, 0.0], dtype=np.float32)
    lin("fc1", 512, 512); lin("fc_shape", 20, 512); lin("fc_glob", 6, 512); lin("fc_cam", 3, 512)
    lin("fc_embed", 256, 512 + 20 + 6 + 3)
    anc = ancestors_from_parents(parents)
    for j in range(23):
        lin(f"fc_pose.{j}.0", 128, 256 + 21 * len(anc[j]))
        lin(f"fc_pose.{j}.2", 9, 128)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def ancestors_from_parents(parents):
    """Ancestor lists per body joint (0-based body-joint ids, nearest ancestor first), i.e. the
    reference's `immediate_parents_to_all_parents` (models/poseMF_shapeGaussian_net.py:14-21)."""
    parents = [int(p) for p in parents]
    anc = {}
    for i in range(1, len(parents)):
        j = i - 1
        ip = parents[i] - 1
        anc[j] = ([ip] + anc[ip]) if ip >= 0 else []
    return anc
