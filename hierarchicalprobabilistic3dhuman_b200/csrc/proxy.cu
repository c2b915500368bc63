// Proxy-representation generation for sm_100a: the step immediately before the encoder.
//
// Replaces
//   * reference models/canny_edge_detector.py:104-166 (CannyEdgeDetector.forward): per-channel separable Gaussian
//     blur -> Sobel gradients summed over channels / C -> magnitude, orientation binned to 45 degrees -> threshold
//     -> directional non-maximum suppression; every nn.Conv2d there zero-pads, so each stage sees ZERO outside the
//     image (not the previous stage evaluated on padding);
//   * reference utils/label_conversions.py:105-124 (convert_2Djoints_to_gaussian_heatmaps_torch) and the visibility
//     mask of predict/predict_poseMF_shapeGaussian_net.py:97-99;
//   * the torch.cat of predict/...:100 -- and, in the fused entry point, the fp32-NCHW -> fp16-NHWC(32) cast in front
//     of the tensor-core encoder: the kernel writes the encoder's input records directly, so the 4.7 MB/image fp32
//     proxy representation never exists in HBM (786 KB of RGB in, 4.2 MB of fp16 records out).
//
// One CTA = one 32x32 output tile of one image; the five stencil stages run out of shared memory with shrinking
// halos (raw 40x40 -> row-blurred 40x36 -> blurred 36x36 -> gradients / magnitude 34x34 -> edges 32x32 for the 5-tap
// filter). Arithmetic follows the reference's fp32 operation order (oneDNN accumulates filter taps in row-major order
// with FMAs from zero; the elementwise torch ops are unfused), so the blurred image is bit-identical and the rest
// differs only where the host's vectorised sqrt (1 ulp off IEEE on 0.6 % of inputs), atan2f or expf differ from
// CUDA's by an ulp.
#include "common.cuh"
#include "encoder.cuh"
#include <cuda_fp16.h>
#include <math.h>

using namespace hp3d;

namespace {

constexpr int CT = 32;                 // output tile edge
constexpr int MAXR = 4;                // Gaussian radius supported (filter size <= 9)
constexpr int MAXK = 32;               // joints per image supported by the fused heat-map path
constexpr int GR = CT + 2;             // gradient / magnitude region edge (tile + 1)
constexpr int BR = CT + 4;             // blurred region edge (tile + 2)
constexpr int RR = BR + 2 * MAXR;      // raw region edge at the largest radius

// directional filters 0..315 degrees: offset (dy, dx) of the neighbour subtracted from the centre (:62-100)
__constant__ int dyn[8] = {0, 1, 1, 1, 0, -1, -1, -1};
__constant__ int dxn[8] = {1, 1, 0, -1, -1, -1, 0, 1};

struct ProxyArgs {
  const float* img; int C, H, W;       // (B,C,H,W) fp32
  float g[2 * MAXR + 1]; int r;        // normalised Gaussian taps, radius
  float threshold; int nms;
  float *blurred, *mag, *ori, *thr_mag, *thin, *thr_thin;   // optional reference outputs
  float* edges; long long edges_stride;                     // optional final edge map, images `edges_stride` floats apart
  const float* joints2d; const unsigned char* vis; int K; float std;   // optional fused heat-maps
  float* heat; long long heat_stride;                       // (B,K,H,W)-like destination, images heat_stride apart
  __half* nhwc32;                                            // optional fp16 NHWC records of 32 channels (edge | K heat | 0)
  int nhwc_split;                                            // 1: 64-channel hi/lo records of the split stem (conv_tc.cu) instead
};

__device__ __forceinline__ float heat_value(float row, float col, float u, float v, float std) {
  // exp(-(((row - v)/std)^2)/2 - (((col - u)/std)^2)/2), unfused fp32 like the reference's tensor expression
  const float q1 = __fdiv_rn(__fsub_rn(row, v), std), q2 = __fdiv_rn(__fsub_rn(col, u), std);
  const float a1 = __fdiv_rn(__fmul_rn(q1, q1), 2.f), a2 = __fdiv_rn(__fmul_rn(q2, q2), 2.f);
  return expf(__fsub_rn(-a1, a2));
}

__global__ void __launch_bounds__(256) proxy_rep_kernel(const ProxyArgs a) {
  __shared__ float raw[RR * RR];
  __shared__ float bh[RR * BR];
  __shared__ float bv[BR * BR];
  __shared__ float sgx[GR * GR], sgy[GR * GR], smag[GR * GR];
  __shared__ float sj[MAXK * 2];
  __shared__ float svis[MAXK];
  const int tid = threadIdx.x;
  const int b = blockIdx.z, y0 = blockIdx.y * CT, x0 = blockIdx.x * CT;
  const int H = a.H, W = a.W, r = a.r, nt = 2 * r + 1;
  const int Rr = BR + 2 * r;                       // raw region edge; rows of bh
  const size_t HW = (size_t)H * W;
  for (int i = tid; i < GR * GR; i += 256) { sgx[i] = 0.f; sgy[i] = 0.f; }
  if (a.joints2d && tid < a.K) {
    sj[2 * tid] = a.joints2d[((size_t)b * a.K + tid) * 2];
    sj[2 * tid + 1] = a.joints2d[((size_t)b * a.K + tid) * 2 + 1];
    svis[tid] = a.vis ? (a.vis[(size_t)b * a.K + tid] ? 1.f : 0.f) : 1.f;
  }
  for (int c = 0; c < a.C; ++c) {
    const float* src = a.img + ((size_t)b * a.C + c) * HW;
    // raw tile + halo, zero outside the image
    for (int i = tid; i < Rr * Rr; i += 256) {
      const int ry = i / Rr, rx = i - ry * Rr;
      const int y = y0 - 2 - r + ry, x = x0 - 2 - r + rx;
      raw[ry * RR + rx] = (y >= 0 && y < H && x >= 0 && x < W) ? src[(size_t)y * W + x] : 0.f;
    }
    __syncthreads();
    // horizontal Gaussian (canny_edge_detector.py:141, inner call)
    for (int i = tid; i < Rr * BR; i += 256) {
      const int ry = i / BR, cx = i - ry * BR;
      const int y = y0 - 2 - r + ry, x = x0 - 2 + cx;
      float acc = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W)
        for (int k = 0; k < nt; ++k) acc = __fmaf_rn(a.g[k], raw[ry * RR + cx + k], acc);
      bh[ry * BR + cx] = acc;
    }
    __syncthreads();
    // vertical Gaussian (outer call)
    for (int i = tid; i < BR * BR; i += 256) {
      const int by = i / BR, cx = i - by * BR;
      const int y = y0 - 2 + by, x = x0 - 2 + cx;
      float acc = 0.f;
      const bool in = y >= 0 && y < H && x >= 0 && x < W;
      if (in)
        for (int k = 0; k < nt; ++k) acc = __fmaf_rn(a.g[k], bh[(by + k) * BR + cx], acc);
      bv[by * BR + cx] = acc;
      if (a.blurred && in && by >= 2 && by < 2 + CT && cx >= 2 && cx < 2 + CT)
        a.blurred[((size_t)b * a.C + c) * HW + (size_t)y * W + x] = acc;
    }
    __syncthreads();
    // Sobel, accumulated over channels (:145-146); taps in row-major filter order, zero weights skipped
    for (int i = tid; i < GR * GR; i += 256) {
      const int sy = i / GR, sx = i - sy * GR;
      const int y = y0 - 1 + sy, x = x0 - 1 + sx;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float* p = bv + (sy + 1) * BR + (sx + 1);     // centre in bv coordinates
        float gx = 0.f, gy = 0.f;
        gx = __fmaf_rn(1.f, p[-BR - 1], gx); gx = __fmaf_rn(-1.f, p[-BR + 1], gx);
        gx = __fmaf_rn(2.f, p[-1], gx);      gx = __fmaf_rn(-2.f, p[1], gx);
        gx = __fmaf_rn(1.f, p[BR - 1], gx);  gx = __fmaf_rn(-1.f, p[BR + 1], gx);
        gy = __fmaf_rn(1.f, p[-BR - 1], gy); gy = __fmaf_rn(2.f, p[-BR], gy); gy = __fmaf_rn(1.f, p[-BR + 1], gy);
        gy = __fmaf_rn(-1.f, p[BR - 1], gy); gy = __fmaf_rn(-2.f, p[BR], gy); gy = __fmaf_rn(-1.f, p[BR + 1], gy);
        sgx[i] = __fadd_rn(sgx[i], gx);
        sgy[i] = __fadd_rn(sgy[i], gy);
      }
    }
    __syncthreads();
  }
  // magnitude on the tile + 1 halo (:149-150); zero outside the image because the gradients are
  const float fc = (float)a.C;
  for (int i = tid; i < GR * GR; i += 256) {
    const float gx = __fdiv_rn(sgx[i], fc), gy = __fdiv_rn(sgy[i], fc);
    sgx[i] = gx; sgy[i] = gy;
    smag[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
  }
  __syncthreads();
  for (int i = tid; i < CT * CT; i += 256) {
    const int ty = i / CT, tx = i - ty * CT;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const int gi = (ty + 1) * GR + (tx + 1);
    const float m = smag[gi];
    // orientation in degrees, binned to multiples of 45 (:151-152)
    float o = __fadd_rn(__fmul_rn(atan2f(sgy[gi], sgx[gi]), 57.29577951308232f), 180.0f);
    o = __fmul_rn(rintf(__fdiv_rn(o, 45.0f)), 45.0f);
    const float tm = (m < a.threshold) ? 0.f : m;
    float thin = m;
    if (a.nms) {
      const int idx = ((int)__fdiv_rn(o, 45.0f)) & 7;                  // (o / 45) % 8
      const int p = idx & 3;
      const float d0 = __fsub_rn(m, smag[gi + dyn[p] * GR + dxn[p]]);
      const float d1 = __fsub_rn(m, smag[gi + dyn[p + 4] * GR + dxn[p + 4]]);
      if (!(fminf(d0, d1) > 0.0f)) thin = 0.f;
    }
    const float tthin = (thin < a.threshold) ? 0.f : thin;
    const size_t pix = (size_t)y * W + x;
    if (a.mag) a.mag[(size_t)b * HW + pix] = m;
    if (a.ori) a.ori[(size_t)b * HW + pix] = o;
    if (a.thr_mag) a.thr_mag[(size_t)b * HW + pix] = tm;
    if (a.thin) a.thin[(size_t)b * HW + pix] = thin;
    if (a.thr_thin) a.thr_thin[(size_t)b * HW + pix] = tthin;
    const float edge = a.nms ? tthin : tm;                             // predict/...:92
    if (a.edges) a.edges[(size_t)b * a.edges_stride + pix] = edge;
    if (a.heat) {
      for (int k = 0; k < a.K; ++k)
        a.heat[(size_t)b * a.heat_stride + (size_t)k * HW + pix] =
            __fmul_rn(heat_value((float)y, (float)x, sj[2 * k], sj[2 * k + 1], a.std), svis[k]);
    }
    if (a.nhwc32 && a.nhwc_split) {
      // [A_hi c0..15 | A_hi c0..15 | A_lo c0..15 | A_hi c16,17 | A_hi c16,17 | A_lo c16,17 | 0 x 10]  (stem2_kernel<true>)
      __half2 h[9], l[9];
      float prev = edge;
#pragma unroll
      for (int k = 0; k < 17; ++k) {                                   // channel k+1 = heat-map k
        const float v = __fmul_rn(heat_value((float)y, (float)x, sj[2 * k], sj[2 * k + 1], a.std), svis[k]);
        if (k & 1) prev = v;
        else {
          h[k >> 1] = __floats2half2_rn(prev, v);
          const float2 f = __half22float2(h[k >> 1]);
          l[k >> 1] = __floats2half2_rn(prev - f.x, v - f.y);
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(a.nhwc32 + ((size_t)b * HW + pix) * 64);
      dst[0] = *reinterpret_cast<const uint4*>(&h[0]); dst[1] = *reinterpret_cast<const uint4*>(&h[4]);
      dst[2] = dst[0]; dst[3] = dst[1];
      dst[4] = *reinterpret_cast<const uint4*>(&l[0]); dst[5] = *reinterpret_cast<const uint4*>(&l[4]);
      const __half2 t[4] = {h[8], h[8], l[8], __float2half2_rn(0.f)};
      dst[6] = *reinterpret_cast<const uint4*>(&t[0]);
      dst[7] = make_uint4(0u, 0u, 0u, 0u);
    } else if (a.nhwc32) {
      __half2 h[16];
      float prev = edge;
#pragma unroll
      for (int k = 0; k < 31; ++k) {                                   // channel k+1 = heat-map k (zero padding past K)
        const float v = (k < a.K) ? __fmul_rn(heat_value((float)y, (float)x, sj[2 * k], sj[2 * k + 1], a.std), svis[k]) : 0.f;
        if (k & 1) prev = v; else h[k >> 1] = __floats2half2_rn(prev, v);
      }
      uint4* dst = reinterpret_cast<uint4*>(a.nhwc32 + ((size_t)b * HW + pix) * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = *reinterpret_cast<const uint4*>(&h[4 * q]);
    }
  }
}

// stand-alone heat-maps for any K / size (label_conversions.py:105-124): thread = pixel, loops over joints
__global__ void __launch_bounds__(256) heatmaps_kernel(const float* __restrict__ j2d, const unsigned char* __restrict__ vis,
                                                       int K, int wh, float std, float* __restrict__ out, long long image_stride) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= wh * wh) return;
  const int y = pix / wh, x = pix - y * wh;
  for (int k = 0; k < K; ++k) {
    const float u = j2d[((size_t)b * K + k) * 2], v = j2d[((size_t)b * K + k) * 2 + 1];
    float h = heat_value((float)y, (float)x, u, v, std);
    if (vis) h = __fmul_rn(h, vis[(size_t)b * K + k] ? 1.f : 0.f);
    out[(size_t)b * image_stride + (size_t)k * wh * wh + pix] = h;
  }
}

// arg-max of the heat-map of joint (u,v) without materialising it (utils/label_conversions.py:127-155 applied to
// :105-124): the maximum of the separable Gaussian lies within one pixel of the rounded joint clamped to the image, so
// a 4x4 neighbourhood evaluated with the SAME fp32 expression as the heat-map kernels, first index winning ties, gives
// the arg-max torch.max would return; visible iff that maximum exceeds eps.
__global__ void joints2d_argmax_kernel(const float* __restrict__ j2d, const unsigned char* __restrict__ vis, int n, int wh,
                                       float std, float eps, float* __restrict__ out, int* __restrict__ vis_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = j2d[2 * i], v = j2d[2 * i + 1];
  const float visf = vis ? (vis[i] ? 1.f : 0.f) : 1.f;
  const float fu = fminf(fmaxf(floorf(u), -2.f), (float)wh + 1.f), fv = fminf(fmaxf(floorf(v), -2.f), (float)wh + 1.f);
  const int cu = (int)fu, cv = (int)fv;
  float best = -1.f; int bi = 0;
  for (int dy = -1; dy <= 2; ++dy)
    for (int dx = -1; dx <= 2; ++dx) {
      const int y = min(max(cv + dy, 0), wh - 1), x = min(max(cu + dx, 0), wh - 1);
      const float h = __fmul_rn(heat_value((float)y, (float)x, u, v, std), visf);
      const int idx = y * wh + x;
      if (h > best || (h == best && idx < bi)) { best = h; bi = idx; }
    }
  if (best == 0.f) bi = 0;                      // an all-zero map: torch.max returns index 0 (irrelevant, invisible)
  const bool ok = best > eps;
  out[2 * i] = ok ? (float)(bi % wh) : -1.f;
  out[2 * i + 1] = ok ? floorf((float)bi / (float)wh) : -1.f;
  vis_out[i] = ok ? 1 : 0;
}

int fill_gauss(ProxyArgs& a, float gaussian_std, int gaussian_size) {
  if (gaussian_size < 1 || gaussian_size > 2 * MAXR + 1 || !(gaussian_size & 1) || !(gaussian_std > 0.f)) {
    set_error("canny: gaussian_filter_size must be odd and <= %d, std > 0", 2 * MAXR + 1);
    return -1;
  }
  // scipy.signal.windows.gaussian(size, std) / sum, in float64, then float32 (canny_edge_detector.py:23-24,31)
  double g[2 * MAXR + 1], s = 0.0;
  for (int k = 0; k < gaussian_size; ++k) {
    const double n = (double)k - (gaussian_size - 1) / 2.0;
    g[k] = exp(-0.5 * (n / (double)gaussian_std) * (n / (double)gaussian_std));
    s += g[k];
  }
  for (int k = 0; k < gaussian_size; ++k) a.g[k] = (float)(g[k] / s);
  a.r = gaussian_size / 2;
  return 0;
}

}  // namespace

extern "C" int hp3d_canny_edges(const float* img, int B, int C, int H, int W, float gaussian_std, int gaussian_size,
                                float threshold, int nms, float* blurred, float* grad_mag, float* grad_ori,
                                float* thr_grad_mag, float* thin_edges, float* thr_thin_edges, float* edges,
                                long long edges_image_stride, void* stream) {
  HP3D_ARG(img && B > 0 && C > 0 && H > 0 && W > 0, "bad argument");
  HP3D_ARG(B <= 65535, "B <= 65535");
  HP3D_ARG(!edges || edges_image_stride >= (long long)H * W, "edges_image_stride < H*W");
  ProxyArgs a = {};
  a.img = img; a.C = C; a.H = H; a.W = W; a.threshold = threshold; a.nms = nms ? 1 : 0;
  if (fill_gauss(a, gaussian_std, gaussian_size)) return -1;
  a.blurred = blurred; a.mag = grad_mag; a.ori = grad_ori; a.thr_mag = thr_grad_mag;
  a.thin = nms ? thin_edges : nullptr; a.thr_thin = nms ? thr_thin_edges : nullptr;
  a.edges = edges; a.edges_stride = edges_image_stride;
  proxy_rep_kernel<<<dim3(cdiv(W, CT), cdiv(H, CT), B), 256, 0, (cudaStream_t)stream>>>(a);
  return launch_status("proxy_rep_kernel");
}

extern "C" int hp3d_joints2d_to_heatmaps(const float* joints2d, const unsigned char* visibility, int B, int K, int img_wh,
                                         float std, float* out, long long out_image_stride, void* stream) {
  HP3D_ARG(joints2d && out && B > 0 && K > 0 && img_wh > 0 && std > 0.f, "bad argument");
  HP3D_ARG(B <= 65535, "B <= 65535");
  HP3D_ARG(out_image_stride >= (long long)K * img_wh * img_wh, "out_image_stride < K*wh*wh");
  heatmaps_kernel<<<dim3(cdiv(img_wh * img_wh, 256), B), 256, 0, (cudaStream_t)stream>>>(joints2d, visibility, K, img_wh, std,
                                                                                        out, out_image_stride);
  return launch_status("heatmaps_kernel");
}

extern "C" int hp3d_proxy_rep(const float* rgb, const float* joints2d, const unsigned char* visibility, int B, int C, int K,
                              int img_wh, float gaussian_std, int gaussian_size, float threshold, int nms, float heat_std,
                              float* out_nchw, void* stream) {
  HP3D_ARG(rgb && joints2d && out_nchw && B > 0 && C > 0 && K > 0 && K <= MAXK && img_wh > 0 && heat_std > 0.f, "bad argument");
  HP3D_ARG(B <= 65535, "B <= 65535");
  ProxyArgs a = {};
  a.img = rgb; a.C = C; a.H = img_wh; a.W = img_wh; a.threshold = threshold; a.nms = nms ? 1 : 0;
  if (fill_gauss(a, gaussian_std, gaussian_size)) return -1;
  const long long stride = (long long)(K + 1) * img_wh * img_wh;
  a.edges = out_nchw; a.edges_stride = stride;
  a.joints2d = joints2d; a.vis = visibility; a.K = K; a.std = heat_std;
  a.heat = out_nchw + (size_t)img_wh * img_wh; a.heat_stride = stride;
  proxy_rep_kernel<<<dim3(cdiv(img_wh, CT), cdiv(img_wh, CT), B), 256, 0, (cudaStream_t)stream>>>(a);
  return launch_status("proxy_rep_kernel");
}

namespace hp3d {
// fused producer of the tensor-core encoder's input (conv_tc.cu): fp16 NHWC records of 32 channels (fast mode) or the
// split stem's 64-channel hi/lo records (split = 1)
int proxy_rep_nhwc_f16(const float* rgb, const float* joints2d, const unsigned char* visibility, int B, int img_wh,
                       float gaussian_std, int gaussian_size, float threshold, int nms, float heat_std, void* nhwc32,
                       int split, cudaStream_t stream) {
  ProxyArgs a = {};
  a.img = rgb; a.C = 3; a.H = img_wh; a.W = img_wh; a.threshold = threshold; a.nms = nms ? 1 : 0;
  if (fill_gauss(a, gaussian_std, gaussian_size)) return -1;
  a.joints2d = joints2d; a.vis = visibility; a.K = 17; a.std = heat_std;
  a.nhwc32 = (__half*)nhwc32; a.nhwc_split = split;
  proxy_rep_kernel<<<dim3(cdiv(img_wh, CT), cdiv(img_wh, CT), B), 256, 0, stream>>>(a);
  return launch_status("proxy_rep_kernel");
}
}  // namespace hp3d

extern "C" int hp3d_joints2d_heatmap_argmax(const float* joints2d, const unsigned char* visibility, int B, int K, int img_wh,
                                            float std, float eps, float* joints2d_px, int32_t* vis_out, void* stream) {
  HP3D_ARG(joints2d && joints2d_px && vis_out && B > 0 && K > 0 && img_wh > 0 && std > 0.f, "bad argument");
  const int n = B * K;
  joints2d_argmax_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(joints2d, visibility, n, img_wh, std, eps, joints2d_px, vis_out);
  return launch_status("joints2d_argmax_kernel");
}
