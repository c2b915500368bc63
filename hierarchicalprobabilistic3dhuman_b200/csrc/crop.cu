// Crop / affine resample and HRNet key-point arg-max for sm_100a: what the reference runs between
// the 2D-pose network and the proxy-representation generator (predict/predict_poseMF_shapeGaussian_net.py:73-93).
//
// Replaces
//   * reference utils/image_utils.py:234-378 (`batch_crop_pytorch_affine`) for the given-bounding-box call of the predict
//     path: RGB resampled bilinearly through F.affine_grid + F.grid_sample (align_corners=False, zero padding), 2D joints
//     through the forward affine. The arithmetic lives in crop_math.h, which is also compiled for the host and checked
//     bit for bit against the reference-pinned oracle (tests/test_crop_host.py).
//   * reference predict/predict_hrnet.py:7-30 (`get_kp_locations_confs_from_heatmaps`).
// STATUS: compiled for sm_100a; the arithmetic is verified on the host; the kernels themselves have not yet run on
// hardware (round 1 ended with the GPU budget spent) -- the GPU tests in tests/test_gpu_crop.py are opt-in until then.
#include "common.cuh"
#include "crop_math.h"
#include <math_constants.h>

using namespace hp3d;

namespace {

// thread = output pixel (all channels): HBM-bound gather of 4 taps per channel, writes coalesced
__global__ void __launch_bounds__(256) crop_rgb_kernel(const float* __restrict__ rgb, int C, int H, int W,
                                                       const float* __restrict__ centres, const float* __restrict__ heights,
                                                       const float* __restrict__ widths, float scale, int out_w, int out_h,
                                                       float* __restrict__ out) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= out_w * out_h) return;
  const int oy = p / out_w, ox = p - oy * out_w;
  const CropXform X = crop_xform((float)W, (float)H, (float)out_w, (float)out_h, centres[2 * b], centres[2 * b + 1], heights[b],
                                 widths[b], scale);
  for (int c = 0; c < C; ++c)
    out[((size_t)b * C + c) * out_w * out_h + p] = crop_sample(rgb + ((size_t)b * C + c) * H * W, H, W, X, ox, oy, out_w, out_h);
}

__global__ void crop_joints_kernel(const float* __restrict__ joints, int n, int K, int H, int W,
                                   const float* __restrict__ centres, const float* __restrict__ heights,
                                   const float* __restrict__ widths, float scale, int out_w, int out_h, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = i / K;
  const CropXform X = crop_xform((float)W, (float)H, (float)out_w, (float)out_h, centres[2 * b], centres[2 * b + 1], heights[b],
                                 widths[b], scale);
  out[2 * i] = __fadd_rn(__fmul_rn(joints[2 * i], X.a00), X.a02);            // einsum as an unfused multiply + add
  out[2 * i + 1] = __fadd_rn(__fmul_rn(joints[2 * i + 1], X.a11), X.a12);
}

// grid (K, B): first index of the maximum of one heat-map, its value, key point zeroed unless the maximum is positive
__global__ void __launch_bounds__(256) heatmap_keypoints_kernel(const float* __restrict__ hm, int K, int hw, int w,
                                                                float* __restrict__ kps, float* __restrict__ confs) {
  const int k = blockIdx.x, b = blockIdx.y;
  const float* p = hm + ((size_t)b * K + k) * hw;
  float best = -CUDART_INF_F; int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < hw; i += 256) {
    const float v = p[i];
    if (v > best) { best = v; bi = i; }
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
    for (int q = 1; q < 8; ++q) if (sv[q] > best || (sv[q] == best && si[q] < bi)) { best = sv[q]; bi = si[q]; }
    const bool pos = best > 0.0f;
    const size_t o = (size_t)b * K + k;
    kps[2 * o] = pos ? (float)(bi % w) : 0.0f;
    kps[2 * o + 1] = pos ? floorf((float)bi / (float)w) : 0.0f;
    confs[o] = best;
  }
}

}  // namespace

extern "C" int hp3d_crop_affine(const float* rgb, const float* joints2d, int B, int C, int H, int W, int K,
                                const float* bbox_centres, const float* bbox_heights, const float* bbox_widths,
                                float scale_factor, int out_w, int out_h, float* rgb_out, float* joints_out, void* stream) {
  HP3D_ARG(bbox_centres && bbox_heights && bbox_widths, "null bounding box");
  HP3D_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && out_w > 1 && out_h > 1 && scale_factor > 0.f, "bad argument");
  HP3D_ARG((rgb == nullptr) == (rgb_out == nullptr) && (joints2d == nullptr) == (joints_out == nullptr), "input / output mismatch");
  HP3D_ARG(!rgb || C > 0, "C > 0");
  HP3D_ARG(!joints2d || K > 0, "K > 0");
  cudaStream_t s = (cudaStream_t)stream;
  if (rgb) {
    crop_rgb_kernel<<<dim3(cdiv(out_w * out_h, 256), B), 256, 0, s>>>(rgb, C, H, W, bbox_centres, bbox_heights, bbox_widths,
                                                                      scale_factor, out_w, out_h, rgb_out);
    const int rc = launch_status("crop_rgb_kernel");
    if (rc) return rc;
  }
  if (joints2d) {
    crop_joints_kernel<<<cdiv(B * K, 128), 128, 0, s>>>(joints2d, B * K, K, H, W, bbox_centres, bbox_heights, bbox_widths,
                                                        scale_factor, out_w, out_h, joints_out);
    return launch_status("crop_joints_kernel");
  }
  return 0;
}

extern "C" int hp3d_heatmap_keypoints(const float* heatmaps, int B, int K, int h, int w, float* keypoints, float* confs,
                                      void* stream) {
  HP3D_ARG(heatmaps && keypoints && confs && B > 0 && B <= 65535 && K > 0 && h > 0 && w > 0, "bad argument");
  heatmap_keypoints_kernel<<<dim3(K, B), 256, 0, (cudaStream_t)stream>>>(heatmaps, K, h * w, w, keypoints, confs);
  return launch_status("heatmap_keypoints_kernel");
}
