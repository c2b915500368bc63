// Matrix-Fisher normalising constant: log c(S) and d log c / d s_k for proper singular values S = (s1 >= s2 >= |s3|),
// shared by the device kernel (mf_norm.cu) and the host test shim so the numerics are checked on CPU without a GPU.
//
// Restates reference losses/matrix_fisher_loss.py:9-192 (`bessel0_exp_scaled`, `torch_trapezoid_integral`,
// `integrand_normconst_forward_exp_scaled`, `integrand_dlognormconst_ds_backward`, `LogMFNormConstant`): the
// exponentially scaled constant c_bar(S) = c(S) / exp(tr S) is a 1-D integral over u in [-1, 1] of
//   I0_bar((s_i - s_j)(1 - u)/2) * I0_bar((s_i + s_j)(1 + u)/2) * exp((s_j + s_k)(u - 1))
// (Lee 2017, arXiv:1710.03746 eq. 85-90), evaluated by the trapezoid rule on 512 nodes with polynomial approximations of
// the scaled modified Bessel function I0_bar(x) = I0(x) / exp(|x|) (Numerical Recipes `bessi0`).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HP3D_MFN_HD __host__ __device__ __forceinline__
#else
#define HP3D_MFN_HD static inline
#endif

namespace hp3d {

#if defined(__CUDA_ARCH__)
#define HP3D_NMUL(a, b) __fmul_rn((a), (b))
#define HP3D_NADD(a, b) __fadd_rn((a), (b))
#define HP3D_NSUB(a, b) __fsub_rn((a), (b))
#define HP3D_NDIV(a, b) __fdiv_rn((a), (b))
#else   // host build: -ffp-contract=off keeps the plain operators unfused, like the reference's tensor ops
#define HP3D_NMUL(a, b) ((a) * (b))
#define HP3D_NADD(a, b) ((a) + (b))
#define HP3D_NSUB(a, b) ((a) - (b))
#define HP3D_NDIV(a, b) ((a) / (b))
#endif

constexpr int MF_NORM_TRAPS = 512;          // matrix_fisher_loss.py:151,180

// I0(x) / exp(|x|), matrix_fisher_loss.py:31-48 (Horner's rule as unfused multiply + add, :15-28)
HP3D_MFN_HD float bessel0_exp_scaled(float x) {
  const float ax = fabsf(x);
  if (ax <= 3.75f) {
    const float q = HP3D_NDIV(ax, 3.75f), t = HP3D_NMUL(q, q);
    float z = 0.45813e-2f;
    z = HP3D_NADD(HP3D_NMUL(z, t), 0.360768e-1f);
    z = HP3D_NADD(HP3D_NMUL(z, t), 0.2659732f);
    z = HP3D_NADD(HP3D_NMUL(z, t), 1.2067492f);
    z = HP3D_NADD(HP3D_NMUL(z, t), 3.0899424f);
    z = HP3D_NADD(HP3D_NMUL(z, t), 3.5156229f);
    z = HP3D_NADD(HP3D_NMUL(z, t), 1.0f);
    return HP3D_NDIV(z, expf(ax));
  }
  const float t = HP3D_NDIV(3.75f, ax);
  float z = 0.392377e-2f;
  z = HP3D_NADD(HP3D_NMUL(z, t), -0.1647633e-1f);
  z = HP3D_NADD(HP3D_NMUL(z, t), 0.2635537e-1f);
  z = HP3D_NADD(HP3D_NMUL(z, t), -0.2057706e-1f);
  z = HP3D_NADD(HP3D_NMUL(z, t), 0.916281e-2f);
  z = HP3D_NADD(HP3D_NMUL(z, t), -0.157565e-2f);
  z = HP3D_NADD(HP3D_NMUL(z, t), 0.225319e-2f);
  z = HP3D_NADD(HP3D_NMUL(z, t), 0.1328592e-1f);
  z = HP3D_NADD(HP3D_NMUL(z, t), 0.39894228f);
  return HP3D_NDIV(z, sqrtf(ax));
}

// trapezoid node i of 512 on [-1, 1] and its weight (matrix_fisher_loss.py:66-71)
HP3D_MFN_HD float mf_norm_node(int i) { return HP3D_NADD(HP3D_NMUL((float)i, (float)(2.0 / (MF_NORM_TRAPS - 1))), -1.0f); }
HP3D_MFN_HD float mf_norm_weight(int i) { return (i == 0 || i == MF_NORM_TRAPS - 1) ? 0.5f : 1.0f; }

// shared integrand: (s_i, s_j) feed the Bessel factors, (s_j + s_k) the exponential (:76-99 with (i,j,k) = (2,3,1); :102-131)
HP3D_MFN_HD float mf_norm_integrand(float u, float s_i, float s_j, float s_k) {
  const float f1 = bessel0_exp_scaled(HP3D_NMUL(HP3D_NMUL(HP3D_NSUB(s_i, s_j), 0.5f), HP3D_NSUB(1.0f, u)));
  const float f2 = bessel0_exp_scaled(HP3D_NMUL(HP3D_NMUL(HP3D_NADD(s_i, s_j), 0.5f), HP3D_NADD(1.0f, u)));
  const float f3 = expf(HP3D_NMUL(HP3D_NADD(s_j, s_k), HP3D_NSUB(u, 1.0f)));
  return HP3D_NMUL(HP3D_NMUL(f1, f2), f3);
}

// weighted integrand values of node i for the forward integral and the three backward integrals (cyclic shifts of S,
// :183-189): out[0] -> c_bar, out[1 + k] -> d c_bar / d s_k + c_bar
HP3D_MFN_HD void mf_norm_node_terms(int i, float s0, float s1, float s2, float out[4]) {
  const float u = mf_norm_node(i), w = mf_norm_weight(i);
  out[0] = HP3D_NMUL(mf_norm_integrand(u, s1, s2, s0), w);
  const float s[3] = {s0, s1, s2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float a = s[(k + 1) % 3], b = s[(k + 2) % 3];
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    out[1 + k] = HP3D_NMUL(HP3D_NMUL(mf_norm_integrand(u, hi, lo, s[k]), u), w);
  }
}

// sums over the 512 nodes -> log c(S) and d log c / d s (matrix_fisher_loss.py:153-164, 183-191)
HP3D_MFN_HD void mf_norm_finish(const float sum[4], float s0, float s1, float s2, float* log_c, float dlogc_ds[3]) {
  const float scale = 1.0f / (float)(MF_NORM_TRAPS - 1);             // sum * (to - from) / (n - 1), then * 0.5
  const float c_bar = HP3D_NMUL(0.5f, HP3D_NMUL(HP3D_NMUL(sum[0], 2.0f), scale));
  *log_c = HP3D_NADD(logf(c_bar), HP3D_NADD(HP3D_NADD(s0, s1), s2));
  if (dlogc_ds)
    for (int k = 0; k < 3; ++k)
      dlogc_ds[k] = HP3D_NDIV(HP3D_NMUL(0.5f, HP3D_NMUL(HP3D_NMUL(sum[1 + k], 2.0f), scale)), c_bar);
}

}  // namespace hp3d
