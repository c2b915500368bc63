/* libhp3d -- C ABI of the B200-native probabilistic-pose inference hot path.
 *
 * The reference (akashsengupta1997/HierarchicalProbabilistic3DHuman) is pure Python and has no FFI;
 * the "interface" each entry point replaces is the Python call cited beside it (paths relative to
 * the reference root). INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - every function returns int: 0 = OK, <0 = argument error, >0 = cudaError_t of the failing call;
 *    hp3d_last_error() returns a thread-local message. Nothing throws or aborts.
 *  - all tensor memory is caller-owned DEVICE memory (fp32, contiguous, 16-byte aligned base
 *    pointers unless noted); model constants passed to *_create are HOST pointers and are copied /
 *    repacked into an immutable opaque handle.
 *  - kernels are enqueued on the given stream; no internal synchronisation. The library keeps no mutable state that changes
 *    results: its only process-level data are idempotent per-DEVICE caches (which kernels have been opted in to > 48 KB of
 *    shared memory on which device ordinal; the driver entry point of cuTensorMapEncodeTiled) and a thread-local error
 *    string, so one process may drive several GPUs and several host threads. Tuning knobs (HP3D_* environment variables,
 *    documented where they are read) are looked up per call, never cached.
 *  - `stream` is a cudaStream_t passed as void* so the header needs no CUDA include.
 */
#ifndef HP3D_H_
#define HP3D_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HP3D_VERSION 200
#define HP3D_NUM_VERTS 6890
#define HP3D_NUM_JOINTS 24          /* SMPL skeleton incl. root */
#define HP3D_NUM_BODY_JOINTS 23
#define HP3D_NUM_BETAS 10
#define HP3D_NUM_OUT_JOINTS 90      /* 24 posed + 21 picked + 45 regressed (models/smpl_official.py:30-34) */

int hp3d_version(void);
const char* hp3d_last_error(void);

/* ---------------------------------------------------------------- SMPL forward
 * replaces: models/smpl_official.py:13-41 (SMPL.__init__/forward) and, through it, smplx 0.1.26
 * lbs() / vertices2joints() / batch_rigid_transform() / VertexJointSelector. */
typedef struct hp3d_smpl hp3d_smpl;
typedef struct {
  const double* v_template;             /* [6890*3]                                   */
  const double* shapedirs;              /* [6890*3*10]  (vertex, xyz, beta)           */
  const double* posedirs;               /* [207 * 20670] (pose feature, vertex*3+xyz) */
  const double* J_regressor;            /* [24*6890]                                  */
  const double* lbs_weights;            /* [6890*24]                                  */
  const int32_t* parents;               /* [24], parents[0] = -1                      */
  const int32_t* extra_vertex_ids;      /* [21] vertices appended as joints 24..44    */
  const double* joint_regressors_extra; /* [45*6890] rows appended as joints 45..89   */
} hp3d_smpl_model;

int hp3d_smpl_create(const hp3d_smpl_model* model, hp3d_smpl** out);
void hp3d_smpl_destroy(hp3d_smpl* h);
/* bytes of scratch hp3d_smpl_forward needs for M meshes with Mb distinct shapes */
size_t hp3d_smpl_workspace_bytes(const hp3d_smpl* h, int M, int Mb);
/* betas [Mb*10]; global_orient [Mg*9] row-major rotmats; body_pose [M*23*9]; M % Mb == 0 and
 * M % Mg == 0: mesh m uses betas row m/(M/Mb) and global_orient row m/(M/Mg) (the reference's
 * per-image expand over N samples, utils/sampling_utils.py:178-185, train/...:304-308).
 * vertices [M*6890*3]; joints [M*90*3] (may be NULL). */
int hp3d_smpl_forward(const hp3d_smpl* h, const float* betas, int Mb, const float* global_orient, int Mg,
                      const float* body_pose, int M, float* vertices, float* joints,
                      void* workspace, size_t workspace_bytes, void* stream);
/* hp3d_smpl_forward for B images x N samples (M = B * samples_per_image, image-major) that ALSO returns the per-vertex
 * statistics of utils/sampling_utils.py:189-190 per image: avg_dist [B*6890] = mean over the image's samples of the
 * distance to the image's mean mesh, mean_vertices [B*6890*3] (may be NULL). With the default fused kernel
 * (csrc/smpl_fused.cu) and 8 <= samples_per_image <= 112 the statistics come out of the SMPL kernel itself (the sample
 * vertices are re-read from L2, never from HBM); otherwise (HP3D_SMPL=staged, other sample counts) it is hp3d_smpl_forward
 * followed by hp3d_vertex_uncertainty. */
int hp3d_smpl_forward_stats(const hp3d_smpl* h, const float* betas, int Mb, const float* global_orient, int Mg,
                            const float* body_pose, int M, int samples_per_image, float* vertices, float* joints,
                            float* avg_dist, float* mean_vertices, void* workspace, size_t workspace_bytes, void* stream);
/* which path a handle takes (fused = 1 unless HP3D_SMPL=staged or no tensor-map support), whether the fused plan's vertices were
 * re-ordered by dominant joint at create time, and the sum / maximum over the 216 32-vertex tiles of distinct skinning
 * joints (the fused kernel's skinning cost is proportional to the sum). */
int hp3d_smpl_layout_info(const hp3d_smpl* h, int* fused, int* permuted, int* tile_joint_sum, int* tile_joint_max);
/* stage-level entry points of the STAGED path (HP3D_SMPL=staged: hp3d_smpl_forward = these three in order; used by tests /
 * profiling). The default hp3d_smpl_forward is ONE tensor-core kernel (csrc/smpl_fused.cu): v_posed never exists in memory. */
int hp3d_smpl_shape_blend(const hp3d_smpl* h, const float* betas, int Mb, float* v_shaped /*[Mb*20672]*/,
                          float* J /*[Mb*24*3]*/, void* stream);
size_t hp3d_smpl_pose_blend_workspace_bytes(int M);   /* fp16 hi/lo pose features for the tensor-core blend */
int hp3d_smpl_pose_blend(const hp3d_smpl* h, const float* betas, const float* v_shaped, int Mb, const float* body_pose,
                         int M, float* v_posed /*[M*20672], row pitch 20672 floats*/, void* workspace, size_t workspace_bytes, void* stream);
int hp3d_smpl_lbs(const hp3d_smpl* h, const float* v_posed, const float* J, int Mb, const float* global_orient,
                  int Mg, const float* body_pose, int M, float* vertices, float* joints, void* stream);
/* smplx batch_rodrigues: axis-angle [n*3] -> rotmats [n*9] (pose2rot=True callers:
 * evaluate/...:176-178, predict/...:136) */
int hp3d_rodrigues(const float* axis_angle, int n, float* rotmats, void* stream);
/* replaces utils/rigid_transform_utils.py:80-94 */
int hp3d_rot6d_to_rotmat(const float* x6, int n, float* rotmats, void* stream);
/* per-vertex sample statistics, replaces utils/sampling_utils.py:189-190 for B images x N samples:
 * vertices [B*N*6890*3] -> mean_vertices [B*6890*3] (may be NULL), avg_dist [B*6890] */
int hp3d_vertex_uncertainty(const float* vertices, int B, int N, float* mean_vertices, float* avg_dist, void* stream);

/* sample ranking by 2D-joint consistency, replaces utils/sampling_utils.py:195-233 batched over B images:
 * joints [B*N*90*3]; heatmaps = pointer to the FIRST of 17 joint heat-maps (H*W floats each) of image 0, consecutive
 * images `heatmap_image_stride` floats apart (for a (B,18,H,W) proxy representation: base + H*W, stride 18*H*W);
 * cam [B*3] weak-perspective (s,tx,ty). Outputs: order [B*N] sample indices by ascending error, err [B*N] (max pixel
 * distance over visible COCO joints), joints2d_out [B*17*2] heat-map arg-max (x,y; -1 if invisible), vis_out [B*17].
 * heatmaps == NULL: joints2d_out / vis_out are INPUTS holding those arg-max joints already (image-space path:
 * hp3d_joints2d_heatmap_argmax), and the heat-maps are never read. */
int hp3d_rank_samples_by_joints2d(const float* joints, const float* heatmaps, long long heatmap_image_stride,
                                  const float* cam, int B, int N, int H, int W, float eps, int32_t* order, float* err,
                                  float* joints2d_out, int32_t* vis_out, void* stream);

/* ---------------------------------------------------------------- proxy-representation generation
 * replaces: models/canny_edge_detector.py:104-166 (CannyEdgeDetector.forward). img [B*C*H*W] fp32 NCHW; Gaussian
 * taps = scipy.signal.windows.gaussian(size, std)/sum (size odd, <= 9); every optional output may be NULL:
 * blurred [B*C*H*W]; grad_mag, grad_ori, thr_grad_mag, thin_edges, thr_thin_edges [B*H*W] (the reference's dict keys;
 * the two thin_* need nms != 0); edges = what predict/...:92 feeds the network (thr_thin_edges if nms else
 * thr_grad_mag), image b written at edges + b*edges_image_stride (e.g. channel 0 of a (B,18,H,W) tensor). */
int hp3d_canny_edges(const float* img, int B, int C, int H, int W, float gaussian_std, int gaussian_size,
                     float threshold, int nms, float* blurred, float* grad_mag, float* grad_ori, float* thr_grad_mag,
                     float* thin_edges, float* thr_thin_edges, float* edges, long long edges_image_stride, void* stream);
/* replaces: utils/label_conversions.py:105-124 (convert_2Djoints_to_gaussian_heatmaps_torch) and the visibility
 * mask of predict/...:97-99. joints2d [B*K*2] as (u = column, v = row); visibility [B*K] bytes or NULL;
 * heat-map k of image b written at out + b*out_image_stride + k*wh*wh. */
int hp3d_joints2d_to_heatmaps(const float* joints2d, const unsigned char* visibility, int B, int K, int img_wh,
                              float std, float* out, long long out_image_stride, void* stream);
/* replaces: predict/...:91-100 in one launch: out_nchw [B*(K+1)*wh*wh] = cat(edges, masked heat-maps). K <= 32. */
int hp3d_proxy_rep(const float* rgb, const float* joints2d, const unsigned char* visibility, int B, int C, int K,
                   int img_wh, float gaussian_std, int gaussian_size, float threshold, int nms, float heat_std,
                   float* out_nchw, void* stream);
/* arg-max pixel and visibility (max > eps) of each joint's heat-map WITHOUT materialising it: what
 * utils/label_conversions.py:127-155 returns for the maps of hp3d_joints2d_to_heatmaps. joints2d_px [B*K*2]
 * (x, y; -1 if invisible), vis_out [B*K]. */
int hp3d_joints2d_heatmap_argmax(const float* joints2d, const unsigned char* visibility, int B, int K, int img_wh,
                                 float std, float eps, float* joints2d_px, int32_t* vis_out, void* stream);

/* ---------------------------------------------------------------- crop / resample in front of it
 * replaces: utils/image_utils.py:234-378 (batch_crop_pytorch_affine) for a GIVEN bounding box (predict/...:84-93):
 * rgb [B*C*H*W] -> rgb_out [B*C*out_h*out_w] by F.affine_grid + bilinear F.grid_sample (align_corners=False, zeros),
 * joints2d [B*K*2] -> joints_out by the forward affine; either pair may be NULL. bbox_centres [B*2] as (vertical,
 * horizontal), bbox_heights / bbox_widths [B]; scale_factor = orig_scale_factor. Arithmetic: csrc/crop_math.h, verified
 * on the host bit for bit; the kernels have not yet run on hardware. */
int hp3d_crop_affine(const float* rgb, const float* joints2d, int B, int C, int H, int W, int K, const float* bbox_centres,
                     const float* bbox_heights, const float* bbox_widths, float scale_factor, int out_w, int out_h,
                     float* rgb_out, float* joints_out, void* stream);
/* replaces: predict/predict_hrnet.py:7-30 (get_kp_locations_confs_from_heatmaps): heatmaps [B*K*h*w] -> keypoints
 * [B*K*2] (x, y of the arg-max; 0 where the maximum is not positive), confs [B*K] (the maxima). */
int hp3d_heatmap_keypoints(const float* heatmaps, int B, int K, int h, int w, float* keypoints, float* confs, void* stream);

/* ---------------------------------------------------------------- matrix-Fisher sampler
 * replaces: utils/sampling_utils.py:74-143 (pose_matrix_fisher_sampling_torch) incl. :10-71
 * (bingham_sampling_for_matrix_fisher_torch) and utils/rigid_transform_utils.py:113-133.
 * U,V [B*J*9], S [B*J*3] (improper LAPACK factors as the head returns them); R_out [B*N*J*9].
 * Noise: if eps/w are non-NULL (eps [B*J*ov*N*4] normals, w [B*J*ov*N] uniforms) the kernel replays
 * the reference's "first N accepted of ov*N, in index order" rule exactly; otherwise Philox4x32-10
 * keyed by (seed, offset) draws proposals until N are accepted (at most max_rounds*32 proposals).
 * stats [3] (device, may be NULL): proposals, accepts, (image,joint) pairs that ran out of proposals. */
int hp3d_mf_sample(const float* U, const float* S, const float* V, int B, int J, int N, float b,
                   uint64_t seed, uint64_t offset, const float* eps, const float* w, int oversampling,
                   float* R_out, unsigned long long* stats, void* stream);
/* The same for one SHARD of a batch (the reference has no multi-GPU path): the B images passed are images
 * [image_offset, image_offset + B) of the global batch. The in-kernel Philox stream is keyed by the GLOBAL (image, joint,
 * lane), so with the same (seed, offset) on every rank the gathered samples are bit-identical to a single-GPU run on the
 * concatenated batch, whatever the world size. hp3d_mf_sample == image_offset 0. */
int hp3d_mf_sample_sharded(const float* U, const float* S, const float* V, int B, int J, int N, float b,
                           uint64_t seed, uint64_t offset, uint64_t image_offset, const float* eps, const float* w,
                           int oversampling, float* R_out, unsigned long long* stats, void* stream);

/* matrix-Fisher normalising constant, replaces losses/matrix_fisher_loss.py:134-192
 * (LogMFNormConstant.forward / backward): S_proper [n*3] proper singular values (s1 >= s2 >= |s3|) -> log_c [n] =
 * log c(S), dlogc_ds [n*3] = d log c / d s_k (may be NULL). 512-node trapezoid rule over scaled Bessel-I0 products.
 * Arithmetic: csrc/mf_norm_math.h, verified on the host; the kernel has not yet run on hardware. */
int hp3d_mf_log_norm_constant(const float* S_proper, int n, float* log_c, float* dlogc_ds, void* stream);

/* ---------------------------------------------------------------- distribution head
 * replaces: models/poseMF_shapeGaussian_net.py:95-160 (everything after the encoder). */
typedef struct hp3d_head hp3d_head;
typedef struct {
  const float *fc1_w, *fc1_b;           /* [512*512], [512]  (out, in) row-major like nn.Linear */
  const float *fc_shape_w, *fc_shape_b; /* [20*512], [20]  */
  const float *fc_glob_w, *fc_glob_b;   /* [6*512], [6]    */
  const float *fc_cam_w, *fc_cam_b;     /* [3*512], [3]    */
  const float *fc_embed_w, *fc_embed_b; /* [256*541], [256] */
  const float* const* fc_pose0_w;       /* 23 pointers, [128*(256+21*n_anc[j])] */
  const float* const* fc_pose0_b;       /* 23 pointers, [128] */
  const float* const* fc_pose2_w;       /* 23 pointers, [9*128] */
  const float* const* fc_pose2_b;       /* 23 pointers, [9] */
  const float *init_glob, *init_cam;    /* [6], [3] */
  const int32_t* parents;               /* [24] SMPL kinematic tree */
  float delta_i_weight;                 /* MODEL.DELTA_I_WEIGHT (0 when MODEL.DELTA_I is False) */
} hp3d_head_weights;
int hp3d_head_create(const hp3d_head_weights* w, hp3d_head** out);
void hp3d_head_destroy(hp3d_head* h);
size_t hp3d_head_workspace_bytes(const hp3d_head* h, int B);
/* feats [B*512] -> F,U,V,mode [B*23*9], S [B*23*3], shape_params [B*20] (mean | log_std),
 * glob [B*6], cam [B*3]. teacher_* (may be NULL) teacher-force the ancestors' inputs (parity tests). */
int hp3d_head_forward(const hp3d_head* h, const float* feats, int B, float* F, float* U, float* S, float* V,
                      float* mode, float* shape_params, float* glob, float* cam,
                      const float* teacher_Uproper, const float* teacher_Sproper, const float* teacher_mode,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- ResNet-18 encoder
 * replaces: models/resnet.py:202-217 (ResNet.forward) for resnet18(in_channels=18), eval-mode BN. */
typedef struct hp3d_encoder hp3d_encoder;
typedef struct {
  const float* w; int cout, cin, k, stride, pad;      /* conv weight [cout*cin*k*k] (OIHW) */
  const float *bn_w, *bn_b, *bn_mean, *bn_var;        /* [cout] each */
} hp3d_conv_bn;
typedef struct {
  hp3d_conv_bn stem;                    /* 7x7 s2 p3, 18->64 */
  hp3d_conv_bn conv[4][2][2];           /* [layer][block][conv1|conv2] */
  hp3d_conv_bn down[4];                 /* 1x1 s2 downsample of layers 2..4 (down[0] unused, w=NULL) */
  float bn_eps;
} hp3d_encoder_weights;
#define HP3D_ENC_PARITY 0   /* fp32 CUDA-core implicit GEMM (plain-fp32 cross-check of the contract)                      */
#define HP3D_ENC_FAST 1     /* fp16 operands, fp32 accumulate, tcgen05 implicit GEMM, ONE product: ~3e-4 on the features  */
#define HP3D_ENC_SPLIT 2    /* tcgen05 implicit GEMM on fp16 hi/lo pairs, three products in fp32 TMEM: <=1e-4 (default)    */
int hp3d_encoder_create(const hp3d_encoder_weights* w, int mode, hp3d_encoder** out);
void hp3d_encoder_destroy(hp3d_encoder* h);
size_t hp3d_encoder_workspace_bytes(const hp3d_encoder* h, int B, int H, int W);
/* x [B*18*H*W] fp32 NCHW (the reference's input layout, predict/...:100) -> feats [B*512] */
int hp3d_encoder_forward(const hp3d_encoder* h, const float* x_nchw, int B, int H, int W, float* feats,
                         void* workspace, size_t workspace_bytes, void* stream);
/* hp3d_encoder_forward that also returns the arg-max pixel (x, y; -1 if max <= eps) and visibility of the 17 joint
 * heat-maps in channels 1..17 (utils/label_conversions.py:127-155), a by-product of the input pass in fast mode:
 * joints2d_px [B*17*2], vis [B*17] -- the inputs hp3d_rank_samples_by_joints2d accepts with heatmaps == NULL. */
int hp3d_encoder_forward_argmax(const hp3d_encoder* h, const float* x_nchw, int B, int H, int W, float* feats,
                                void* workspace, size_t workspace_bytes, float eps, float* joints2d_px, int32_t* vis,
                                void* stream);
/* hp3d_encoder_forward / _argmax for a proxy representation the caller already holds in fp16 (x_nchw_f16 [B*18*H*W] halves;
 * tensor-core handles only). Opt-in: it halves the bytes a host has to push over PCIe (the fp32 contract input makes the
 * end-to-end path PCIe-bound); the arithmetic is unchanged, the input values are whatever fp16 holds. joints2d_px / vis may
 * both be NULL (no arg-max by-product). */
int hp3d_encoder_forward_f16in(const hp3d_encoder* h, const void* x_nchw_f16, int B, int H, int W, float* feats,
                               void* workspace, size_t workspace_bytes, float eps, float* joints2d_px, int32_t* vis,
                               void* stream);
/* image-space entry (tensor-core handles only: HP3D_ENC_SPLIT / HP3D_ENC_FAST): rgb [B*3*256*256] in [0,1], joints2d [B*17*2], visibility [B*17]
 * bytes or NULL -> feats. The Canny + heat-map kernel writes the stem's fp16 NHWC input records directly: the fp32
 * proxy representation of predict/...:100 never exists in memory. Same workspace as hp3d_encoder_forward. */
int hp3d_encoder_forward_image(const hp3d_encoder* h, const float* rgb, const float* joints2d,
                               const unsigned char* visibility, int B, int img_wh, float gaussian_std, int gaussian_size,
                               float threshold, int nms, float heat_std, float* feats, void* workspace,
                               size_t workspace_bytes, void* stream);
/* same, additionally dumping every post-activation tensor as fp32 NHWC into `taps` in the order
 * stem (B,H/2,W/2,64) | maxpool (B,H/4,W/4,64) | layer1.0 | layer1.1 | ... | layer4.1  (parity debugging of the
 * per-layer kernels against models/resnet.py:203-212; taps == NULL behaves like hp3d_encoder_forward). */
int hp3d_encoder_forward_taps(const hp3d_encoder* h, const float* x_nchw, int B, int H, int W, float* feats,
                              void* workspace, size_t workspace_bytes, float* taps, void* stream);

/* ---------------------------------------------------------------- multi-GPU gather by peer stores
 * New (the reference has no distributed code). Copies `bytes` (a multiple of 16) from `src` to each of the
 * `n_peers` (<= 15) destination pointers -- the same slice of every peer's gather buffer, mapped into this process through
 * CUDA peer / symmetric memory -- with a grid of `ctas` (<= 0: default) 128-thread, no-shared-memory CTAs that are small
 * enough to run NEXT TO the hot path's kernels on the same SMs, so the NVLink transfer of chunk c overlaps the kernels of
 * chunk c + 1 (NCCL's all-gather kernels cannot: DESIGN.md 6). Ends with a system-scope fence; the caller exchanges a
 * completion flag afterwards. */
int hp3d_peer_push(const void* src, void* const* peer_dsts, int n_peers, size_t bytes, int ctas, void* stream);
/* The same through an NVSwitch multicast (NVLS) address of the gather buffer: `multicast_dst` is the slice's address in the
 * multicast mapping (torch symmetric memory: handle.multicast_ptr + byte offset); one multimem.st reaches every GPU of the
 * group, so the slice leaves this GPU once instead of (world - 1) times. */
int hp3d_peer_push_multicast(const void* src, void* multicast_dst, size_t bytes, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HP3D_H_ */
