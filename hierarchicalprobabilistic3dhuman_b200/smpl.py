"""Drop-in for the reference's `models.smpl_official.SMPL` (reference models/smpl_official.py:13-41;
constructed at run_predict.py:61-64) running on libhp3d's fused sm_100a kernels.

Same constructor and forward() call conventions (betas / body_pose / global_orient, pose2rot),
same outputs (.vertices (M,6890,3), .joints (M,90,3), ...), `.parents`, `.faces`.
The SMPL model file is licence-gated: if `model_path` does not hold SMPL_{GENDER}.pkl/.npz, or a
`model=` dict is given, a seeded synthetic SMPL-shaped model is used (see synthetic.py) -- reported
through `.is_synthetic`.
"""
import ctypes
import os
import pickle
import warnings
from collections import namedtuple

import numpy as np
import torch
from torch import nn

from . import _lib
from .synthetic import synthetic_smpl_model, load_joint_regressor_table, SMPL_EXTRA_VERTEX_IDS

SMPLOutput = namedtuple("SMPLOutput", ["vertices", "joints", "full_pose", "betas", "global_orient", "body_pose"])
SMPLOutput.__new__.__defaults__ = (None,) * 6


def _load_model_file(model_path, gender):
    """Real SMPL file -> dict in the layouts smplx uses. Returns None if absent."""
    if model_path is None:
        return None
    cands = []
    if os.path.isdir(model_path):
        g = gender.upper()
        cands = [os.path.join(model_path, f"SMPL_{g}.npz"), os.path.join(model_path, f"SMPL_{g}.pkl")]
    elif os.path.isfile(model_path):
        cands = [model_path]
    for p in cands:
        if not os.path.isfile(p):
            continue
        if p.endswith(".npz"):
            d = dict(np.load(p, allow_pickle=True))
        else:
            with open(p, "rb") as f:
                d = pickle.load(f, encoding="latin1")
        arr = lambda k: np.asarray(d[k].todense() if hasattr(d[k], "todense") else d[k], dtype=np.float64)
        posedirs = arr("posedirs")                      # (6890,3,207) in the file
        nv = posedirs.shape[0]
        model = dict(v_template=arr("v_template"), shapedirs=arr("shapedirs")[:, :, :10],
                     posedirs=posedirs.reshape(nv * 3, -1).T.copy(),    # smplx: (207, 20670)
                     J_regressor=arr("J_regressor"), lbs_weights=arr("weights"),
                     parents=np.asarray(d["kintree_table"])[0].astype(np.int64),
                     faces=np.asarray(d["f"]).astype(np.int64))
        model["parents"][0] = -1
        model["extra_vertex_ids"] = SMPL_EXTRA_VERTEX_IDS.copy()
        model["joint_regressors_extra"] = load_joint_regressor_table()
        return model
    return None


class SMPL(nn.Module):
    NUM_BODY_JOINTS = 23

    def __init__(self, model_path=None, batch_size=1, gender="neutral", num_betas=10, model=None, **kwargs):
        super().__init__()
        if num_betas != 10:
            raise ValueError("libhp3d SMPL supports num_betas == 10 (reference default MODEL.NUM_SMPL_BETAS)")
        self.batch_size = batch_size
        self.gender = gender
        self.num_betas = num_betas
        self.is_synthetic = False
        if model is None:
            model = _load_model_file(model_path, gender)
        if model is None:
            warnings.warn("SMPL model file not found (licence-gated); using the seeded synthetic SMPL-shaped model")
            model = synthetic_smpl_model()
            self.is_synthetic = True
        self._model = {k: np.ascontiguousarray(v) for k, v in model.items()}
        self.faces = self._model["faces"]
        self.register_buffer("faces_tensor", torch.as_tensor(self.faces, dtype=torch.long))
        self.register_buffer("parents", torch.as_tensor(self._model["parents"], dtype=torch.long))
        # smplx keeps default (zero) parameters that are used when an argument is omitted
        self.betas = nn.Parameter(torch.zeros(batch_size, num_betas), requires_grad=False)
        self.global_orient = nn.Parameter(torch.zeros(batch_size, 3), requires_grad=False)
        self.body_pose = nn.Parameter(torch.zeros(batch_size, 69), requires_grad=False)
        self._handles = {}
        self._ws = _lib.Workspace()

    # ------------------------------------------------------------------ handle management
    def _handle(self, device):
        idx = torch.device(device).index
        key = idx if idx is not None else torch.cuda.current_device()    # an index-less 'cuda' device is the CURRENT one
        h = self._handles.get(key)
        if h is None:
            L = _lib.lib()
            m = self._model
            keep = dict(
                v_template=np.ascontiguousarray(m["v_template"], np.float64),
                shapedirs=np.ascontiguousarray(m["shapedirs"], np.float64),
                posedirs=np.ascontiguousarray(m["posedirs"], np.float64),
                J_regressor=np.ascontiguousarray(m["J_regressor"], np.float64),
                lbs_weights=np.ascontiguousarray(m["lbs_weights"], np.float64),
                parents=np.ascontiguousarray(m["parents"], np.int32),
                extra_vertex_ids=np.ascontiguousarray(m["extra_vertex_ids"], np.int32),
                joint_regressors_extra=np.ascontiguousarray(m["joint_regressors_extra"], np.float64))
            assert keep["v_template"].shape == (6890, 3) and keep["posedirs"].shape == (207, 20670)
            assert keep["shapedirs"].shape == (6890, 3, 10) and keep["lbs_weights"].shape == (6890, 24)
            sm = _lib.SmplModel(**{k: v.ctypes.data_as(ctypes.c_void_p) for k, v in keep.items()})
            out = ctypes.c_void_p()
            with torch.cuda.device(key):
                _lib.check(L.hp3d_smpl_create(ctypes.byref(sm), ctypes.byref(out)), "hp3d_smpl_create")
            h = out
            self._handles[key] = h
        return h

    def __del__(self):
        try:
            L = _lib.lib()
            for h in self._handles.values():
                L.hp3d_smpl_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_verts=True,
                return_full_pose=False, pose2rot=True, **kwargs):
        """Same contract as smplx.SMPL.forward as extended by the reference (models/smpl_official.py:27-41).
        pose2rot=False: body_pose (M,23,3,3), global_orient (M,1,3,3) [or (Mg,1,3,3) with M % Mg == 0].
        pose2rot=True : axis-angle body_pose (M,69), global_orient (M,3). betas (Mb,10), M % Mb == 0."""
        dev = self.parents.device
        for t in (betas, body_pose, global_orient):
            if t is not None:
                dev = t.device
                break
        if dev.type != "cuda":
            raise RuntimeError("SMPL.forward: tensors/module must be on a CUDA device (no CPU fallback)")
        L = _lib.lib()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        betas = f32(self.betas if betas is None else betas).reshape(-1, 10)
        with torch.cuda.device(dev):
            if pose2rot:
                bp = f32(self.body_pose if body_pose is None else body_pose).reshape(-1, 69)
                go = f32(self.global_orient if global_orient is None else global_orient).reshape(-1, 3)
                M = max(bp.shape[0], go.shape[0], betas.shape[0])
                if bp.shape[0] != M: bp = bp.expand(M, -1).contiguous()
                if go.shape[0] != M: go = go.expand(M, -1).contiguous()
                bp_r = torch.empty(M, 23, 3, 3, device=dev, dtype=torch.float32)
                go_r = torch.empty(M, 1, 3, 3, device=dev, dtype=torch.float32)
                _lib.check(L.hp3d_rodrigues(bp.data_ptr(), M * 23, bp_r.data_ptr(), _lib.stream_ptr()), "hp3d_rodrigues")
                _lib.check(L.hp3d_rodrigues(go.data_ptr(), M, go_r.data_ptr(), _lib.stream_ptr()), "hp3d_rodrigues")
                full_pose = torch.cat([go, bp], dim=1)
                body_pose_out, global_orient_out = bp, go
            else:
                if body_pose is None or global_orient is None:
                    raise ValueError("SMPL.forward(pose2rot=False) needs body_pose (M,23,3,3) and global_orient (M,1,3,3) rotation "
                                     "matrices (the stored default parameters are axis-angle)")
                bp_r = f32(body_pose).reshape(-1, 23, 3, 3)
                go_r = f32(global_orient).reshape(-1, 1, 3, 3)
                M = bp_r.shape[0]
                full_pose = None
                body_pose_out, global_orient_out = body_pose, global_orient
            Mb, Mg = betas.shape[0], go_r.shape[0]
            if M % Mb or M % Mg:
                raise ValueError(f"batch sizes do not broadcast: meshes {M}, betas {Mb}, global_orient {Mg}")
            verts = torch.empty(M, 6890, 3, device=dev, dtype=torch.float32)
            joints = torch.empty(M, 90, 3, device=dev, dtype=torch.float32)
            h = self._handle(dev)
            nbytes = L.hp3d_smpl_workspace_bytes(h, M, Mb)
            ws = self._ws.get(nbytes, dev)
            _lib.check(L.hp3d_smpl_forward(h, betas.data_ptr(), Mb, go_r.data_ptr(), Mg, bp_r.data_ptr(), M,
                                           verts.data_ptr(), joints.data_ptr(), ws.data_ptr(), ws.numel(),
                                           _lib.stream_ptr()), "hp3d_smpl_forward")
            if transl is not None:
                # smplx translates vertices and its 45 joints AFTER lbs(); the reference's three extra regressors
                # (models/smpl_official.py:30-32) then act on the TRANSLATED vertices, i.e. joint r moves by rowsum_r * transl
                t = f32(transl).reshape(-1, 1, 3)
                verts = verts + t
                rs = torch.ones(90, device=dev, dtype=torch.float32)
                rs[45:] = torch.as_tensor(self._model["joint_regressors_extra"].sum(axis=1), dtype=torch.float32, device=dev)
                joints = joints + rs[None, :, None] * t
            if full_pose is None:
                go_full = go_r if Mg == M else go_r.repeat_interleave(M // Mg, dim=0)
                full_pose = torch.cat([go_full, bp_r], dim=1)
        return SMPLOutput(vertices=verts, joints=joints, full_pose=full_pose, betas=betas,
                          global_orient=global_orient_out, body_pose=body_pose_out)
