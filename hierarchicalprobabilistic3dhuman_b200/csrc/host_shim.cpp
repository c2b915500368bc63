// Host build of the shared host/device math headers so the numerics that cannot be exercised
// without a GPU (LAPACK-convention 3x3 SVD) are checked on CPU by `-m "not gpu"` tests.
// Test shim only; the product path never calls it.
#include "svd3.h"
extern "C" void hp3d_host_svd3(const float* A, long n, float* U, float* S, float* V) {
  for (long i = 0; i < n; ++i) hp3d::svd3_lapack(A + 9 * i, U + 9 * i, S + 3 * i, V + 9 * i);
}

// crop / affine resample (csrc/crop_math.h) on the host: checked bit for bit against oracle/crop_oracle.py
#include "crop_math.h"
extern "C" void hp3d_host_crop(const float* rgb, const float* joints, int B, int C, int H, int W, int K, const float* centres,
                               const float* heights, const float* widths, float scale, int out_w, int out_h, float* rgb_out,
                               float* joints_out) {
  for (int b = 0; b < B; ++b) {
    const hp3d::CropXform X = hp3d::crop_xform((float)W, (float)H, (float)out_w, (float)out_h, centres[2 * b], centres[2 * b + 1],
                                               heights[b], widths[b], scale);
    if (rgb)
      for (int c = 0; c < C; ++c)
        for (int oy = 0; oy < out_h; ++oy)
          for (int ox = 0; ox < out_w; ++ox)
            rgb_out[(((long)b * C + c) * out_h + oy) * out_w + ox] =
                hp3d::crop_sample(rgb + ((long)b * C + c) * H * W, H, W, X, ox, oy, out_w, out_h);
    if (joints)
      for (int k = 0; k < K; ++k) {
        joints_out[((long)b * K + k) * 2] = joints[((long)b * K + k) * 2] * X.a00 + X.a02;
        joints_out[((long)b * K + k) * 2 + 1] = joints[((long)b * K + k) * 2 + 1] * X.a11 + X.a12;
      }
  }
}
