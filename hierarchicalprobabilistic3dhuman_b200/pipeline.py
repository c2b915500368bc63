"""Batched hot-path driver: the reference's per-image sequence
(predict/predict_poseMF_shapeGaussian_net.py:102-165) for B images x N samples in one pass,

    proxy rep -> encoder -> MF head -> rot6d -> SMPL(mode) -> MF sampler -> SMPL(B*N) -> per-vertex uncertainty

with the two things a Python caller cannot get from the drop-in modules alone:
  * outputs written straight into caller-provided (e.g. all-gather) buffers, no pack/copy;
  * `run_host`: inputs in pinned HOST memory are streamed to the GPU in chunks on a copy stream while the
    encoder already works on the chunks that have landed, results are read back on a third stream, and
    consecutive calls overlap (double-buffered staging) -- PCIe, not the kernels, bounds this path.
"""
import torch

from . import _lib
from .rigid import rot6d_to_rotmat
from .sampling import pose_matrix_fisher_sampling_torch


class HotPathPipeline:
    def __init__(self, net, smpl, batch, num_samples, device, rotmats_out=None, betas_out=None, vertices_out=None,
                 uncertainty_out=None, chunks=4, on_vertices_chunk=None, image_offset=0):
        """`vertices_out` may be one (B,N,6890,3) tensor or a list of equally sized chunk tensors (cb,N,6890,3) --
        e.g. this rank's slices of per-chunk all-gather buffers; `on_vertices_chunk(c)` is then called right after
        chunk c's SMPL kernels are enqueued, so a collective on another stream can overlap the next chunk.
        `image_offset`: global index of this pipeline's first image when the batch is sharded over ranks (keys the sampler's
        Philox stream, so results do not depend on the world size)."""
        self.net, self.smpl, self.B, self.N, self.dev = net, smpl, batch, num_samples, torch.device(device)
        B, N, dev = batch, num_samples, self.dev
        mk = lambda t, *s: t if t is not None else torch.empty(*s, device=dev, dtype=torch.float32)
        self.rotmats = mk(rotmats_out, B, N, 23, 3, 3)
        self.betas = mk(betas_out, B, 10)
        if isinstance(vertices_out, (list, tuple)):
            self.vertex_chunks = list(vertices_out)
            self.vertices = None
            assert B % len(self.vertex_chunks) == 0
        else:
            self.vertices = mk(vertices_out, B, N, 6890, 3)
            self.vertex_chunks = [self.vertices]
        self.on_vertices_chunk = on_vertices_chunk
        self.image_offset = int(image_offset)
        self.uncertainty = mk(uncertainty_out, B, 6890)
        for t in [self.rotmats, self.betas, self.uncertainty] + self.vertex_chunks:
            assert t.is_contiguous() and t.device == dev
        self.joints = torch.empty(B * N, 90, 3, device=dev)
        self.L = _lib.lib()
        self.h_smpl = smpl._handle(dev)
        cb = B // len(self.vertex_chunks)
        self.ws = torch.empty(self.L.hp3d_smpl_workspace_bytes(self.h_smpl, cb * N, cb), dtype=torch.uint8, device=dev)
        # host-streaming state
        self.chunks = chunks if B % chunks == 0 else 1
        self._stage = None
        self._slot = 0
        # libhp3d kernels launched by one pass, counted on the ncu launch list (profiles/): encoder 24 (input cast, arg-max
        # decode, stem, max-pool, 19 convolutions, avg-pool), head 6, rot6d 1, mode SMPL 4 (feature split, FK, fused kernel,
        # extra joints; + 1 memset node), sampler 1, per vertex chunk SMPL + statistics 4, sample ranking 1
        self.launches_per_pass = 24 + 6 + 1 + 4 + 1 + 4 * len(self.vertex_chunks) + 1

    # ------------------------------------------------------------------ device-resident pass
    def _after_encoder(self, feats, proxy_rep=None, joints2d=None, joints2d_px=None):
        L, B, N = self.L, self.B, self.N
        with _lib.nvtx("hp3d.head"):
            F, U, S, V, mode, shape_params, glob, cam = self.net.head(feats)
            self.betas.copy_(shape_params[:, :10])       # one strided copy straight into the (gather) output slice
            loc = self.betas
            glob_R = rot6d_to_rotmat(glob)
        with _lib.nvtx("hp3d.smpl_mode"):
            out_mode = self.smpl(body_pose=mode, global_orient=glob_R.unsqueeze(1), betas=loc, pose2rot=False)
        with _lib.nvtx("hp3d.mf_sampler"):
            R = pose_matrix_fisher_sampling_torch(U, S, V, N, out=self.rotmats, image_offset=self.image_offset)
        C = len(self.vertex_chunks)
        cb = B // C
        for c, vch in enumerate(self.vertex_chunks):       # images [c*cb, (c+1)*cb): SMPL on cb*N meshes + statistics
            i0 = c * cb
            with _lib.nvtx("hp3d.smpl_samples+statistics"):
                # SMPL on the chunk's cb*N meshes AND the per-vertex statistics of its cb images in one call: with the default
                # fused kernel (8 <= N <= 112) v_posed never exists and the sample vertices are never re-read from HBM
                _lib.check(L.hp3d_smpl_forward_stats(self.h_smpl, loc[i0:].data_ptr(), cb, glob_R[i0:].data_ptr(), cb, R[i0:].data_ptr(),
                                                     cb * N, N, vch.data_ptr(), self.joints[i0 * N:].data_ptr(),
                                                     self.uncertainty[i0:].data_ptr(), None, self.ws.data_ptr(), self.ws.numel(),
                                                     _lib.stream_ptr()), "hp3d_smpl_forward_stats")
            if self.on_vertices_chunk is not None:
                self.on_vertices_chunk(c)
        vertices = self.vertices if self.vertices is not None else (self.vertex_chunks[0] if C == 1 else None)
        res = dict(mode_vertices=out_mode.vertices, mode_joints=out_mode.joints, joints=self.joints, rotmats=R,
                   uncertainty=self.uncertainty, vertices=vertices, betas=self.betas, pose_S=S, cam=cam)
        if proxy_rep is not None or joints2d is not None or joints2d_px is not None:
            # rank the N samples of every image by 2D-joint consistency (sampling_utils.py:195-233)
            from .sampling import rank_samples_by_joints2d
            with _lib.nvtx("hp3d.rank_samples"):
              rk = rank_samples_by_joints2d(self.joints.view(B, N, 90, 3), proxy_rep, cam, joints2d=joints2d, joints2d_px=joints2d_px,
                                          img_wh=proxy_rep.shape[-1] if proxy_rep is not None else 256)
            res["sample_order"], res["sample_error"] = rk["order"], rk["error"]
        return res

    def run_device(self, x_dev):
        """x_dev (B,18,256,256) fp32 on the GPU -> dict of device tensors (sample vertices live in `vertices`)."""
        with torch.cuda.device(self.dev):
            with _lib.nvtx("hp3d.encoder"):
                feats, j2d, vis = self.net.encode(x_dev, return_joints2d=True)     # heat-map arg-max rides on the input pass
            return self._after_encoder(feats, joints2d_px=(j2d, vis))

    def run_device_images(self, rgb, joints2D, visibility=None):
        """Image-space input: (B,3,256,256) RGB crops in [0,1], (B,17,2) joints, (B,17)
        visibility on the GPU -> the same dict as `run_device`; proxy representation generated in-kernel."""
        with torch.cuda.device(self.dev):
            with _lib.nvtx("hp3d.encoder_from_image"):
                feats = self.net.encode_image(rgb, joints2D, visibility)
            return self._after_encoder(feats, None, joints2d=(joints2D, visibility))

    # ------------------------------------------------------------------ host-streaming pass
    def _staging(self, x_host=None):
        if self._stage is None:
            B, dev = self.B, self.dev
            mk_host = lambda *s: torch.empty(*s, dtype=torch.float32).pin_memory()
            self._stage = dict(
                x={},                                                  # input dtype -> two device staging buffers
                x_free=[None, None],                                   # event: encoder finished reading slot
                out=[dict(mode_vertices=mk_host(B, 6890, 3), joints=mk_host(B * self.N, 90, 3),
                          rotmats=mk_host(B, self.N, 23, 3, 3), uncertainty=mk_host(B, 6890)) for _ in range(2)],
                out_done=[None, None],                                  # event: D2H of slot finished
                h2d=torch.cuda.Stream(device=dev), d2h=torch.cuda.Stream(device=dev))
        if x_host is not None and x_host.dtype not in self._stage["x"]:
            self._stage["x"][x_host.dtype] = [torch.empty_like(x_host, device=self.dev) for _ in range(2)]
        return self._stage

    def release_host_vertices(self):
        """free the 2 x (B,N,6890,3) pinned host buffers `run_host(..., return_vertices=True)` allocated"""
        if self._stage is not None:
            for o in self._stage["out"]:
                o.pop("vertices", None)

    def run_host(self, x_host, return_vertices=False):
        """x_host: pinned (B,18,256,256) fp32 (or, opt-in, fp16) HOST tensor. Returns (dict of pinned host result tensors, event);
        the results are valid once `event.synchronize()` returns. Calls may be issued back to back: copies of
        call i+1 overlap the kernels of call i. `return_vertices=True` also copies the (B,N,6890,3) sampled vertices back
        (8.3 MB per image: the reference's consumer keeps them on the device, predict/...:157-165)."""
        assert x_host.is_pinned() and x_host.shape[0] == self.B
        st = self._staging(x_host)
        if return_vertices:
            assert self.vertices is not None or len(self.vertex_chunks) == 1, "return_vertices needs a single (B,N,6890,3) vertices buffer"
            for o in st["out"]:
                if "vertices" not in o:
                    o["vertices"] = torch.empty(self.B, self.N, 6890, 3, dtype=torch.float32).pin_memory()
        else:
            for o in st["out"]:
                o.pop("vertices", None)
        slot = self._slot
        self._slot ^= 1
        B, C = self.B, self.chunks
        cb = B // C
        with torch.cuda.device(self.dev):
            main = torch.cuda.current_stream()
            xbuf = st["x"][x_host.dtype][slot]     # fp32 (the reference's input type) or fp16 (opt-in: half the PCIe bytes)
            evs = []
            with torch.cuda.stream(st["h2d"]):
                if st["x_free"][slot] is not None:
                    st["h2d"].wait_event(st["x_free"][slot])           # encoder of two calls ago is done with the slot
                for c in range(C):
                    xbuf[c * cb:(c + 1) * cb].copy_(x_host[c * cb:(c + 1) * cb], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(st["h2d"])
                    evs.append(e)
            feats, j2d, vis = [], [], []
            for c in range(C):
                main.wait_event(evs[c])
                f_, j_, v_ = self.net.encode(xbuf[c * cb:(c + 1) * cb], return_joints2d=True)
                feats.append(f_); j2d.append(j_); vis.append(v_)
            free = torch.cuda.Event()
            free.record(main)
            st["x_free"][slot] = free
            if st["out_done"][slot ^ 1] is not None:
                main.wait_event(st["out_done"][slot ^ 1])               # previous call's results have left the device buffers
            cat = lambda ts: torch.cat(ts) if C > 1 else ts[0]
            res = self._after_encoder(cat(feats), joints2d_px=(cat(j2d), cat(vis)))
            done = torch.cuda.Event()
            done.record(main)
            out = st["out"][slot]
            with torch.cuda.stream(st["d2h"]):
                st["d2h"].wait_event(done)
                for k in out:
                    res[k].record_stream(st["d2h"])
                    out[k].copy_(res[k], non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(st["d2h"])
            st["out_done"][slot] = fin
        return out, fin

    def run_host_images(self, rgb_host, joints2d_host, vis_host):
        """Like `run_host` from image-space inputs in pinned HOST memory: (B,3,256,256) fp32 RGB, (B,17,2) fp32 joints,
        (B,17) uint8 visibility -- 0.79 MB per image over PCIe instead of 4.72 MB of fp32 proxy representation."""
        assert rgb_host.is_pinned() and rgb_host.shape[0] == self.B
        if getattr(self, "_istage", None) is None:
            dev = self.dev
            self._istage = dict(rgb=[torch.empty_like(rgb_host, device=dev) for _ in range(2)],
                                j2d=[torch.empty_like(joints2d_host, device=dev) for _ in range(2)],
                                vis=[torch.empty_like(vis_host, device=dev) for _ in range(2)])
        st = self._staging()                         # shares the output staging / streams with run_host
        ist = self._istage
        slot = self._slot
        self._slot ^= 1
        B, C = self.B, self.chunks
        cb = B // C
        with torch.cuda.device(self.dev):
            main = torch.cuda.current_stream()
            rgb, j2d, vis = ist["rgb"][slot], ist["j2d"][slot], ist["vis"][slot]
            evs = []
            with torch.cuda.stream(st["h2d"]):
                if st["x_free"][slot] is not None:
                    st["h2d"].wait_event(st["x_free"][slot])
                j2d.copy_(joints2d_host, non_blocking=True)
                vis.copy_(vis_host, non_blocking=True)
                for c in range(C):
                    rgb[c * cb:(c + 1) * cb].copy_(rgb_host[c * cb:(c + 1) * cb], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(st["h2d"])
                    evs.append(e)
            feats = []
            for c in range(C):
                main.wait_event(evs[c])
                sl = slice(c * cb, (c + 1) * cb)
                feats.append(self.net.encode_image(rgb[sl], j2d[sl], vis[sl]))
            free = torch.cuda.Event()
            free.record(main)
            st["x_free"][slot] = free
            if st["out_done"][slot ^ 1] is not None:
                main.wait_event(st["out_done"][slot ^ 1])
            res = self._after_encoder(torch.cat(feats) if C > 1 else feats[0], None, joints2d=(j2d, vis))
            done = torch.cuda.Event()
            done.record(main)
            out = st["out"][slot]
            with torch.cuda.stream(st["d2h"]):
                st["d2h"].wait_event(done)
                for k in out:
                    res[k].record_stream(st["d2h"])
                    out[k].copy_(res[k], non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(st["d2h"])
            st["out_done"][slot] = fin
        return out, fin

    def d2h_bytes(self):
        B, N = self.B, self.N
        return 4 * (B * 6890 * 3 + B * N * 90 * 3 + B * N * 23 * 9 + B * 6890)
