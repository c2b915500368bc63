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

// matrix-Fisher normalising constant (csrc/mf_norm_math.h) on the host, nodes summed sequentially
#include "mf_norm_math.h"
extern "C" void hp3d_host_mf_log_norm(const float* S, long n, float* log_c, float* dlogc_ds) {
  for (long r = 0; r < n; ++r) {
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < hp3d::MF_NORM_TRAPS; ++i) {
      float t[4];
      hp3d::mf_norm_node_terms(i, S[3 * r], S[3 * r + 1], S[3 * r + 2], t);
      for (int q = 0; q < 4; ++q) sum[q] += t[q];
    }
    hp3d::mf_norm_finish(sum, S[3 * r], S[3 * r + 1], S[3 * r + 2], log_c + r, dlogc_ds ? dlogc_ds + 3 * r : nullptr);
  }
}
