"""Import the UNMODIFIED reference from /root/reference (build container only; absent on the GPU
box). TEST INFRASTRUCTURE -- see oracle/__init__.py. Nothing under `-m gpu`, smoke() or bench.py
may depend on this succeeding."""
import os
import sys
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("HP3D_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def reference_config():
    # the 6 keys the hot path reads (reference configs/poseMF_shapeGaussian_net_config.py:8-13)
    return SimpleNamespace(MODEL=SimpleNamespace(NUM_IN_CHANNELS=18, NUM_RESNET_LAYERS=18, EMBED_DIM=256,
                                                 DELTA_I=True, DELTA_I_WEIGHT=1.0, NUM_SMPL_BETAS=10))


def import_reference():
    """Returns a namespace with the reference's PoseMFShapeGaussianNet, sampler and rotation utils."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # read-only mount
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        from models.poseMF_shapeGaussian_net import PoseMFShapeGaussianNet
        from utils.sampling_utils import pose_matrix_fisher_sampling_torch, bingham_sampling_for_matrix_fisher_torch
        from utils.rigid_transform_utils import rot6d_to_rotmat, quat_to_rotmat
    return SimpleNamespace(PoseMFShapeGaussianNet=PoseMFShapeGaussianNet,
                           pose_matrix_fisher_sampling_torch=pose_matrix_fisher_sampling_torch,
                           bingham_sampling_for_matrix_fisher_torch=bingham_sampling_for_matrix_fisher_torch,
                           rot6d_to_rotmat=rot6d_to_rotmat, quat_to_rotmat=quat_to_rotmat,
                           config=reference_config())
