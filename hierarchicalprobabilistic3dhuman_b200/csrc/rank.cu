// Sample ranking by 2D-joint consistency for sm_100a.
//
// Replaces reference utils/sampling_utils.py:195-233 (`joints2D_error_sorted_verts_sampling`) batched over B images:
//   1. arg-max of the 17 joint heat-maps (utils/label_conversions.py:127-155) -> input 2D joints + visibility
//   2. per (image, sample): COCO joints (ALL_JOINTS_TO_COCO_MAP, :17) flipped 180 degrees about x, weak-perspective
//      projection s*(X+t) (utils/cam_utils.py:9-16), pixel space (p+1)*W/2 (utils/joints2d_utils.py:5-10),
//      max over visible joints of the L2 distance to the input joints (:222-227)
//   3. ascending sort of the N errors per image (:228) -> sample order
// The heat-maps are channels 1..17 of the proxy representation that is already on the device.
#include "common.cuh"
#include "encoder.cuh"
#include <math_constants.h>

using namespace hp3d;

namespace {

__constant__ int c_coco_map[17] = {24, 26, 25, 28, 27, 16, 17, 18, 19, 20, 21, 1, 2, 4, 5, 7, 8};

// grid (17, B): first index of the maximum of one heat-map (ties -> lowest index, like torch.max on CPU)
__global__ void __launch_bounds__(256) heatmap_argmax_kernel(const float* __restrict__ hm, size_t image_stride, int HW, int W,
                                                             float eps, float* __restrict__ j2d, int* __restrict__ vis) {
  const int k = blockIdx.x, b = blockIdx.y;
  const float* p = hm + (size_t)b * image_stride + (size_t)k * HW;
  float best = -CUDART_INF_F; int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < HW; i += 256) {
    const float v = p[i];
    if (v > best) { best = v; bi = i; }           // i increases per thread: keeps the first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  __shared__ float sv[8]; __shared__ int si[8];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    const bool v = best > eps;
    j2d[((size_t)b * 17 + k) * 2] = v ? (float)(bi % W) : -1.f;
    j2d[((size_t)b * 17 + k) * 2 + 1] = v ? floorf((float)bi / (float)W) : -1.f;
    vis[b * 17 + k] = v ? 1 : 0;
  }
}

// one CTA per image: errors of the N samples, then a bitonic sort of (error, index) in shared memory
__global__ void __launch_bounds__(256) rank_samples_kernel(const float* __restrict__ joints, const float* __restrict__ cam,
                                                           const float* __restrict__ j2d, const int* __restrict__ vis, int N,
                                                           int W, int* __restrict__ order, float* __restrict__ err_out) {
  extern __shared__ unsigned long long keys[];     // [P] (error bits << 32 | index), P = next pow2 >= N
  const int b = blockIdx.x;
  int P = 1; while (P < N) P <<= 1;
  __shared__ float sj[17][2]; __shared__ int sv[17];
  if (threadIdx.x < 17) { sj[threadIdx.x][0] = j2d[(b * 17 + threadIdx.x) * 2]; sj[threadIdx.x][1] = j2d[(b * 17 + threadIdx.x) * 2 + 1]; sv[threadIdx.x] = vis[b * 17 + threadIdx.x]; }
  __syncthreads();
  const float s = cam[b * 3], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2], half = (float)W / 2.0f;
  for (int n = threadIdx.x; n < P; n += 256) {
    float e = CUDART_INF_F;
    if (n < N) {
      e = 0.f;
      const float* J = joints + ((size_t)b * N + n) * NOUTJ * 3;
#pragma unroll
      for (int k = 0; k < 17; ++k) {
        if (!sv[k]) continue;
        const float X = J[c_coco_map[k] * 3], Y = -J[c_coco_map[k] * 3 + 1];          // 180 degrees about x: (X, -Y, -Z)
        const float px = (s * (X + tx) + 1.f) * half, py = (s * (Y + ty) + 1.f) * half;
        const float dx = px - sj[k][0], dy = py - sj[k][1];
        e = fmaxf(e, sqrtf(dx * dx + dy * dy));
      }
      err_out[(size_t)b * N + n] = e;
    }
    keys[n] = ((unsigned long long)__float_as_uint(e) << 32) | (unsigned)n;           // errors are >= 0: bit order == value order
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += 256) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], c = keys[l];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  for (int n = threadIdx.x; n < N; n += 256) order[(size_t)b * N + n] = (int)(keys[n] & 0xffffffffu);
}

}  // namespace

namespace hp3d {
int heatmap_argmax(const float* heatmaps, long long image_stride, int B, int H, int W, float eps, float* joints2d_px,
                   int* vis, cudaStream_t stream) {
  heatmap_argmax_kernel<<<dim3(17, B), 256, 0, stream>>>(heatmaps, (size_t)image_stride, H * W, W, eps, joints2d_px, vis);
  return launch_status("heatmap_argmax_kernel");
}
}  // namespace hp3d

extern "C" int hp3d_rank_samples_by_joints2d(const float* joints, const float* heatmaps, long long heatmap_image_stride,
                                             const float* cam, int B, int N, int H, int W, float eps, int32_t* order,
                                             float* err, float* joints2d_out, int32_t* vis_out, void* stream_) {
  HP3D_ARG(joints && cam && order && err && joints2d_out && vis_out, "null argument");
  HP3D_ARG(B > 0 && N > 0 && N <= 4096 && H > 0 && W > 0, "need B > 0, 0 < N <= 4096");
  cudaStream_t s = (cudaStream_t)stream_;
  int rc = 0;
  if (heatmaps) {     // else joints2d_out / vis_out already hold the input joints (hp3d_joints2d_heatmap_argmax)
    heatmap_argmax_kernel<<<dim3(17, B), 256, 0, s>>>(heatmaps, (size_t)heatmap_image_stride, H * W, W, eps, joints2d_out, vis_out);
    rc = launch_status("heatmap_argmax_kernel");
    if (rc) return rc;
  }
  int P = 1; while (P < N) P <<= 1;
  rank_samples_kernel<<<B, 256, (size_t)P * sizeof(unsigned long long), s>>>(joints, cam, joints2d_out, vis_out, N, W, order, err);
  return launch_status("rank_samples_kernel");
}
