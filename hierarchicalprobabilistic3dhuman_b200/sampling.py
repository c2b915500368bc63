"""Drop-ins for the reference's utils/sampling_utils.py on libhp3d kernels:
`pose_matrix_fisher_sampling_torch` (:74-143) and
`compute_vertex_uncertainties_by_poseMF_shapeGaussian_sampling` (:146-192), plus the batched
(B images x N samples) composition the reference only spells out in its training loop
(train/train_poseMF_shapeGaussian_net.py:293-308)."""
import ctypes
import torch

from . import _lib



class SamplerShortfall(RuntimeError):
    """The rejection loop of some (image, joint) pair ran out of proposals (see `check_sampler_status`)."""


class _StatusRing:
    """Asynchronous shortfall watch. The reference redraws ALL proposals of an (image, joint) pair when fewer than N were
    accepted (utils/sampling_utils.py:50,68-69); the kernel instead draws until N are accepted and gives up only after a
    bound that is never reached in practice (acceptance >= 0.43, 8x / Philox: (64 + N/2) x 32 proposals) -- in which case it
    fills the remaining samples with the mode and counts the pair in stats[2]. That must not pass silently, and checking
    must not cost the hot path a sync, an allocation or a fill kernel: every device owns ONE cumulative counter triple (the
    kernel only ever adds to it) and a small ring of pinned host snapshots; each call copies the counters behind its kernel,
    and the NEXT call on the device (or an explicit `check_sampler_status()`) raises if the shortfall count has grown."""
    SLOTS = 8

    def __init__(self):
        self.dev = {}          # device index -> dict(stats, host ring, events, next slot, last seen fail count)

    def state(self, dev):
        st = self.dev.get(dev.index)
        if st is None:
            st = dict(stats=torch.zeros(3, device=dev, dtype=torch.int64),
                      host=[torch.zeros(3, dtype=torch.int64).pin_memory() for _ in range(self.SLOTS)],
                      ev=[None] * self.SLOTS, slot=0, seen=0)
            self.dev[dev.index] = st
        return st

    def post(self, dev):
        st = self.state(dev)
        i = st["slot"]
        if st["ev"][i] is not None:
            st["ev"][i].synchronize()          # ring wrapped around (8 calls without a poll): the oldest snapshot is long done
            self._check(st, i)
        st["host"][i].copy_(st["stats"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        st["ev"][i] = ev
        st["slot"] = (i + 1) % self.SLOTS

    def _check(self, st, i):
        st["ev"][i] = None
        fails = int(st["host"][i][2])
        if fails > st["seen"]:
            new, st["seen"] = fails - st["seen"], fails
            raise SamplerShortfall(
                f"matrix-Fisher sampler: {new} (image, joint) chunk(s) ran out of proposals and were completed with the "
                f"distribution mode ({int(st['host'][i][1])} accepted of {int(st['host'][i][0])} proposals so far on this device). "
                "The reference would redraw (utils/sampling_utils.py:68-69).")

    def poll(self, dev=None, wait=False):
        for idx in ([dev.index] if dev is not None else list(self.dev)):
            st = self.dev.get(idx)
            if st is None:
                continue
            for i in range(self.SLOTS):
                ev = st["ev"][i]
                if ev is None:
                    continue
                if wait:
                    ev.synchronize()
                if ev.query():
                    self._check(st, i)


_status = _StatusRing()


def check_sampler_status(device=None):
    """Block until every sampler launch issued so far (on `device`, default all) has reported, and raise SamplerShortfall
    if one of them ran out of proposals."""
    _status.poll(torch.device(device) if device is not None else None, wait=True)


def _rng_seed_offset(device, n_rounds_bound):
    """(seed, offset) from torch's per-device Philox generator, advancing it so successive calls
    draw fresh streams (the reference consumes the global generator after torch.manual_seed)."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed = gen.initial_seed()
    off = gen.get_offset()
    gen.set_offset(off + 4 * ((2 * n_rounds_bound + 3) // 4))
    return seed & 0xFFFFFFFFFFFFFFFF, off


def pose_matrix_fisher_sampling_torch(pose_U, pose_S, pose_V, num_samples, b=1.5, oversampling_ratio=8,
                                      sample_on_cpu=False, noise=None, return_stats=False, out=None, image_offset=0):
    """(B,J,3,3),(B,J,3),(B,J,3,3) -> (B,num_samples,J,3,3) rotation samples of M(U S V^T).
    `sample_on_cpu` is accepted for signature compatibility and ignored (everything runs in one
    kernel). `noise=(eps, w)` with eps (B,J,oversampling_ratio*N,4) standard normals and
    w (B,J,oversampling_ratio*N) uniforms replays the reference's accept/compact rule exactly;
    without it an in-kernel Philox stream seeded from torch's CUDA generator is used.
    `out` may be a slice of a larger (e.g. all-gather) buffer. `image_offset`: index of pose_U[0] in the GLOBAL batch when
    the batch is sharded over ranks -- the Philox stream is keyed by the global image index, so with the same generator
    state on every rank the gathered samples do not depend on the world size.
    A shortfall of the rejection loop (reference: redraw, :68-69) raises SamplerShortfall -- immediately with injected
    noise, at the next call / `check_sampler_status()` in the asynchronous Philox mode."""
    _lib.require_cuda(pose_U, "pose_U")
    dev = pose_U.device
    B, J = pose_U.shape[0], pose_U.shape[1]
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    U, S, V = f32(pose_U), f32(pose_S), f32(pose_V)
    if out is None:
        out = torch.empty(B, num_samples, J, 3, 3, device=dev, dtype=torch.float32)
    else:
        assert out.is_contiguous() and out.shape == (B, num_samples, J, 3, 3) and out.dtype == torch.float32
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    # hot path (Philox, no stats requested): the device's cumulative counters -- no allocation, no fill kernel
    hot = noise is None and not return_stats
    stats = _status.state(dev)["stats"] if hot else torch.zeros(3, device=dev, dtype=torch.int64)
    eps_p = w_p = None
    seed = off = 0
    if noise is not None:
        eps, w = f32(noise[0]), f32(noise[1])
        assert eps.shape == (B, J, oversampling_ratio * num_samples, 4) and w.shape == (B, J, oversampling_ratio * num_samples)
        eps_p, w_p = eps.data_ptr(), w.data_ptr()
    else:
        # the kernel may split an image's samples over up to max(1, N // 32) CTAs, each with its own block of Philox counters
        seed, off = _rng_seed_offset(dev, (64 + 16 * ((num_samples + 31) // 32)) * max(1, num_samples // 32))
    with torch.cuda.device(dev):
        _status.poll(dev)                      # a shortfall reported by an EARLIER launch surfaces here (no sync)
        _lib.check(_lib.lib().hp3d_mf_sample_sharded(U.data_ptr(), S.data_ptr(), V.data_ptr(), B, J, num_samples, float(b),
                                                     seed, off, int(image_offset), eps_p, w_p, int(oversampling_ratio),
                                                     out.data_ptr(), stats.data_ptr(), _lib.stream_ptr()), "hp3d_mf_sample_sharded")
        if return_stats:
            return out, stats
        if noise is not None:                  # parity / replay mode: not a hot path, check right away
            if int(stats[2].item()) > 0:
                raise SamplerShortfall(f"injected noise exhausted for {int(stats[2].item())} (image, joint) chunk(s): the reference "
                                       "would redraw a fresh block (utils/sampling_utils.py:68-69); pass more candidates")
        else:
            _status.post(dev)
    return out


def vertex_uncertainty(vertices):
    """vertices (B,N,6890,3) -> (mean (B,6890,3), avg distance from the mean (B,6890));
    reference utils/sampling_utils.py:189-190 batched over images."""
    _lib.require_cuda(vertices, "vertices")
    v = vertices.detach().to(torch.float32).contiguous()
    B, N = v.shape[0], v.shape[1]
    mean = torch.empty(B, 6890, 3, device=v.device, dtype=torch.float32)
    dist = torch.empty(B, 6890, device=v.device, dtype=torch.float32)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().hp3d_vertex_uncertainty(v.data_ptr(), B, N, mean.data_ptr(), dist.data_ptr(),
                                                      _lib.stream_ptr()), "hp3d_vertex_uncertainty")
    return mean, dist


def compute_vertex_uncertainties_by_poseMF_shapeGaussian_sampling(pose_U, pose_S, pose_V, shape_distribution,
                                                                  glob_rotmats, num_samples, smpl_model,
                                                                  use_mean_shape=False):
    """Reference signature and returns (utils/sampling_utils.py:146-192); batch size must be 1."""
    assert pose_U.shape[0] == pose_S.shape[0] == pose_V.shape[0] == 1
    R = pose_matrix_fisher_sampling_torch(pose_U, pose_S, pose_V, num_samples, b=1.5, oversampling_ratio=8)
    if use_mean_shape:
        betas = shape_distribution.loc                                  # (1,10): broadcast inside the kernel
    else:
        betas = shape_distribution.sample([num_samples])[:, 0, :]
    out = smpl_model(body_pose=R[0], global_orient=glob_rotmats.unsqueeze(1), betas=betas, pose2rot=False)
    _, dist = vertex_uncertainty(out.vertices[None])
    return dist[0], out.vertices, out.joints


def rank_samples_by_joints2d(joints_samples, proxy_rep_or_heatmaps, cam_wp, eps=1e-6, joints2d=None, img_wh=256, std=4.0,
                             joints2d_px=None):
    """Batched core of the reference's `joints2D_error_sorted_verts_sampling` (utils/sampling_utils.py:195-233):
    joints_samples (B,N,90,3), proxy representation (B,18,H,W) or heat-maps (B,17,H,W), cam_wp (B,3) ->
    dict(order (B,N) int64 sample indices by ascending 2D-joint error, error (B,N), joints2d (B,17,2), vis (B,17)).
    Image-space path: pass `joints2d=(joints2D (B,17,2), visibility (B,17) or None)` and None for the heat-maps; their
    arg-max is then computed without materialising them (hp3d_joints2d_heatmap_argmax). `joints2d_px=(px (B,17,2),
    vis (B,17) int32)` passes an arg-max that is already known (`PoseMFShapeGaussianNet.encode(..., return_joints2d=True)`)."""
    _lib.require_cuda(joints_samples, "joints_samples")
    dev = joints_samples.device
    J = joints_samples.detach().to(torch.float32).contiguous()
    B, N = J.shape[0], J.shape[1]
    cam = cam_wp.detach().to(device=dev, dtype=torch.float32).contiguous()
    order = torch.empty(B, N, device=dev, dtype=torch.int32)
    err = torch.empty(B, N, device=dev, dtype=torch.float32)
    if joints2d_px is not None:
        j2d, vis = joints2d_px[0].contiguous(), joints2d_px[1].to(torch.int32).contiguous()
        assert j2d.shape == (B, 17, 2) and j2d.dtype == torch.float32
        hm_ptr, stride, H, W = None, 0, img_wh, img_wh
    elif joints2d is not None:
        from .proxy import joints2d_heatmap_argmax
        j2d, vis = joints2d_heatmap_argmax(joints2d[0], joints2d[1], img_wh, std, eps)
        assert j2d.shape == (B, 17, 2)
        hm_ptr, stride, H, W = None, 0, img_wh, img_wh
    else:
        hm = proxy_rep_or_heatmaps.detach().to(device=dev, dtype=torch.float32).contiguous()
        C, H, W = hm.shape[1], hm.shape[2], hm.shape[3]
        assert C in (17, 18) and hm.shape[0] == B
        hm_ptr, stride = hm[:, C - 17:].data_ptr(), C * H * W     # heat-maps start at channel C-17
        j2d = torch.empty(B, 17, 2, device=dev, dtype=torch.float32)
        vis = torch.empty(B, 17, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().hp3d_rank_samples_by_joints2d(J.data_ptr(), hm_ptr, stride, cam.data_ptr(), B, N, H, W,
                                                            float(eps), order.data_ptr(), err.data_ptr(), j2d.data_ptr(),
                                                            vis.data_ptr(), _lib.stream_ptr()), "hp3d_rank_samples_by_joints2d")
    return dict(order=order.long(), error=err, joints2d=j2d, vis=vis.bool())


def joints2D_error_sorted_verts_sampling(pred_vertices_samples, pred_joints_samples, input_joints2D_heatmaps, pred_cam_wp):
    """Reference signature (utils/sampling_utils.py:195-233): (N,6890,3), (N,90,3), (1,17,H,W), (1,3) -> vertices
    samples sorted by consistency of their projected 2D joints with the input heat-maps."""
    r = rank_samples_by_joints2d(pred_joints_samples[None], input_joints2D_heatmaps, pred_cam_wp)
    return pred_vertices_samples[r["order"][0]]


def sample_meshes_batched(pose_U, pose_S, pose_V, shape_distribution, glob_rotmats, num_samples, smpl_model,
                          use_mean_shape=True, noise=None, rotmats_out=None):
    """B images x N samples in three launches (sampler, SMPL, statistics): returns dict with
    rotmats (B,N,23,3,3), vertices (B,N,6890,3), joints (B,N,90,3), betas, per_vertex_uncertainty (B,6890)."""
    B = pose_U.shape[0]
    R = pose_matrix_fisher_sampling_torch(pose_U, pose_S, pose_V, num_samples, noise=noise, out=rotmats_out)
    if use_mean_shape:
        betas = shape_distribution.loc                                  # (B,10), each image's row reused N times
    else:
        betas = shape_distribution.sample([num_samples]).transpose(0, 1).reshape(B * num_samples, -1)
    out = smpl_model(body_pose=R.view(B * num_samples, 23, 3, 3), global_orient=glob_rotmats.reshape(B, 1, 3, 3),
                     betas=betas, pose2rot=False)
    verts = out.vertices.view(B, num_samples, 6890, 3)
    mean, dist = vertex_uncertainty(verts)       # (HotPathPipeline gets both from ONE kernel: hp3d_smpl_forward_stats)
    return dict(rotmats=R, vertices=verts, joints=out.joints.view(B, num_samples, 90, 3), betas=betas,
                mean_vertices=mean, per_vertex_uncertainty=dist)
