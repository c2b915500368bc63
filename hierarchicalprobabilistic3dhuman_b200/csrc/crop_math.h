// Crop / affine-resample arithmetic shared by the device kernel (crop.cu) and the host test shim (host_shim.cpp), so the
// numerics are checked on CPU against the bit-pinned oracle (oracle/crop_oracle.py) without a GPU.
//
// Restates reference utils/image_utils.py:305-378 (`batch_crop_pytorch_affine` with a given bounding box: aspect-ratio
// fix, scale, forward affine for the joints, normalised inverse affine) and the ATen operators it calls, in THEIR fp32
// operation order as measured in the build container:
//   * torch.linspace: symmetric halves, fused multiply-add  (i < n/2 ? fma(i, step, -1) : fma(-(n-1-i), step, 1));
//   * F.affine_grid: base = linspace * (n-1) / n, grid = base @ theta^T as an UNFUSED multiply + add (K = 3 sgemm);
//   * F.grid_sample (bilinear, zeros, align_corners=False): source coordinate = fma(g + 1, size/2, -0.5); corner weights
//     w = x - floor(x), e = 1 - w, n = y - floor(y), s = 1 - n; value = nw*v_nw, then fused multiply-adds for ne, sw, se.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HP3D_CROP_HD __host__ __device__ __forceinline__
#else
#define HP3D_CROP_HD static inline
#endif

namespace hp3d {

#if defined(__CUDA_ARCH__)
#define HP3D_CMUL(a, b) __fmul_rn((a), (b))
#define HP3D_CADD(a, b) __fadd_rn((a), (b))
#define HP3D_CSUB(a, b) __fsub_rn((a), (b))
#define HP3D_CDIV(a, b) __fdiv_rn((a), (b))
#define HP3D_CFMA(a, b, c) __fmaf_rn((a), (b), (c))
#else   // host build: compiled with -ffp-contract=off, so the plain operators stay unfused
#define HP3D_CMUL(a, b) ((a) * (b))
#define HP3D_CADD(a, b) ((a) + (b))
#define HP3D_CSUB(a, b) ((a) - (b))
#define HP3D_CDIV(a, b) ((a) / (b))
#define HP3D_CFMA(a, b, c) fmaf((a), (b), (c))
#endif

struct CropXform {
  float a00, a11, a02, a12;   // forward pixel transform (joints):  x' = a00 x + a02,  y' = a11 y + a12
  float t00, t11, t02, t12;   // normalised inverse transform handed to affine_grid
};

// bbox centre (vertical, horizontal), height, width; image_utils.py:305-349 without the random augmentations
HP3D_CROP_HD CropXform crop_xform(float in_w, float in_h, float out_w, float out_h, float c_v, float c_h, float bh, float bw,
                                  float scale) {
  const float aspect = HP3D_CDIV(out_h, out_w);
  if (bh > HP3D_CMUL(bw, aspect)) bw = HP3D_CDIV(bh, aspect);
  if (bh < HP3D_CMUL(bw, aspect)) bh = HP3D_CMUL(bw, aspect);
  bh = HP3D_CMUL(bh, scale);
  bw = HP3D_CMUL(bw, scale);
  CropXform X;
  const float sx = HP3D_CDIV(out_w, bw), sy = HP3D_CDIV(out_h, bh);
  X.a00 = sx; X.a11 = sy;
  X.a02 = HP3D_CSUB(HP3D_CMUL(out_w, 0.5f), HP3D_CMUL(sx, c_h));
  X.a12 = HP3D_CSUB(HP3D_CMUL(out_h, 0.5f), HP3D_CMUL(sy, c_v));
  X.t00 = HP3D_CDIV(bw, in_w); X.t11 = HP3D_CDIV(bh, in_h);
  const float u = HP3D_CDIV(-X.a02, sx), v = HP3D_CDIV(-X.a12, sy);
  X.t02 = HP3D_CSUB(HP3D_CADD(HP3D_CDIV(u, HP3D_CMUL(in_w, 0.5f)), X.t00), 1.0f);
  X.t12 = HP3D_CSUB(HP3D_CADD(HP3D_CDIV(v, HP3D_CMUL(in_h, 0.5f)), X.t11), 1.0f);
  return X;
}

// affine_grid base coordinate of output index i along an axis of n samples (align_corners=False)
HP3D_CROP_HD float crop_base_coord(int i, int n) {
  const float step = HP3D_CDIV(2.0f, (float)(n - 1));
  const float lin = (i < n / 2) ? HP3D_CFMA((float)i, step, -1.0f) : HP3D_CFMA(-(float)(n - 1 - i), step, 1.0f);
  return HP3D_CDIV(HP3D_CMUL(lin, (float)(n - 1)), (float)n);
}

// one bilinear sample of a (H, W) plane at output pixel (ox, oy) of an (out_h, out_w) crop
HP3D_CROP_HD float crop_sample(const float* plane, int H, int W, const CropXform& X, int ox, int oy, int out_w, int out_h) {
  const float gx = HP3D_CADD(HP3D_CMUL(crop_base_coord(ox, out_w), X.t00), X.t02);
  const float gy = HP3D_CADD(HP3D_CMUL(crop_base_coord(oy, out_h), X.t11), X.t12);
  const float ix = HP3D_CFMA(HP3D_CADD(gx, 1.0f), (float)W / 2.0f, -0.5f);
  const float iy = HP3D_CFMA(HP3D_CADD(gy, 1.0f), (float)H / 2.0f, -0.5f);
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float w = HP3D_CSUB(ix, x0), e = HP3D_CSUB(1.0f, w), n = HP3D_CSUB(iy, y0), s = HP3D_CSUB(1.0f, n);
  // float -> int conversions of far-out-of-range coordinates are clamped first (the taps are zero there anyway)
  const float xc = fminf(fmaxf(x0, -2.0f), (float)W + 1.0f), yc = fminf(fmaxf(y0, -2.0f), (float)H + 1.0f);
  const int xi = (int)xc, yi = (int)yc;
  const bool xin0 = xi >= 0 && xi < W, xin1 = xi + 1 >= 0 && xi + 1 < W;
  const bool yin0 = yi >= 0 && yi < H, yin1 = yi + 1 >= 0 && yi + 1 < H;
  const float v_nw = (xin0 && yin0) ? plane[(long)yi * W + xi] : 0.0f;
  const float v_ne = (xin1 && yin0) ? plane[(long)yi * W + xi + 1] : 0.0f;
  const float v_sw = (xin0 && yin1) ? plane[(long)(yi + 1) * W + xi] : 0.0f;
  const float v_se = (xin1 && yin1) ? plane[(long)(yi + 1) * W + xi + 1] : 0.0f;
  float acc = HP3D_CMUL(v_nw, HP3D_CMUL(s, e));
  acc = HP3D_CFMA(v_ne, HP3D_CMUL(s, w), acc);
  acc = HP3D_CFMA(v_sw, HP3D_CMUL(n, e), acc);
  acc = HP3D_CFMA(v_se, HP3D_CMUL(n, w), acc);
  return acc;
}

}  // namespace hp3d
