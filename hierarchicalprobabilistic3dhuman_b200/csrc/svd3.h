// 3x3 single-precision SVD with LAPACK *gesdd sign/ordering conventions, host+device.
//
// The reference decomposes every joint's F on the CPU with torch.svd (reference
// models/poseMF_shapeGaussian_net.py:137), i.e. LAPACK sgesdd('S'), and feeds the *sign-convention
// dependent* factor U (made proper) to the children's MLPs (:126-130,148-152). A "canonical sign"
// SVD would therefore change predictions. This file re-implements, for n = 3 only and fully
// unrolled in registers, the path sgesdd takes for a small square matrix:
//   sgebd2 (Householder bidiagonalisation: H1 (3), G1 (2), H2 (2))
//   -> sbdsdc/slasdq -> sbdsqr on the 3x3 upper bidiagonal, rotations accumulated on identity
//      (slartg, slas2 shift, slasv2 2x2 blocks, zero-shift and shifted implicit QR sweeps, both
//      chase directions, negative-sigma fix on rows of VT, descending sort)
//   -> sormbr: U = H1 H2 U_b,  VT = VT_b G1.
// The restatement is from the published LAPACK 3.x algorithm (netlib); no LAPACK source is
// vendored. Compiled for the device by nvcc and for the host by g++ (tests/ build a tiny host shim
// to check sign agreement against torch.svd on CPU without a GPU).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HP3D_HD __host__ __device__ __forceinline__
#else
#define HP3D_HD static inline
#endif

namespace hp3d {

#define HP3D_SVD_EPS 5.9604644775390625e-8f     /* slamch('E') = 2^-24 */
#define HP3D_SVD_SAFMIN 1.17549435e-38f         /* slamch('S') */

HP3D_HD float svd_sign(float a, float b) { return copysignf(fabsf(a), b); }  // Fortran SIGN(a,b)

HP3D_HD float svd_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);      // keep LAPACK's un-fused rounding on the device
#else
  return a * b;
#endif
}
HP3D_HD float svd_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
HP3D_HD float svd_sub(float a, float b) { return svd_add(a, -b); }
HP3D_HD float svd_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
HP3D_HD float svd_sqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// slapy2: sqrt(x^2+y^2) without unnecessary overflow
HP3D_HD float svd_lapy2(float x, float y) {
  float xa = fabsf(x), ya = fabsf(y);
  float w = fmaxf(xa, ya), z = fminf(xa, ya);
  if (z == 0.f || w > 3.0e38f) return w;
  float q = svd_div(z, w);
  return svd_mul(w, svd_sqrt(svd_add(1.f, svd_mul(q, q))));
}

// slartg (LAPACK >= 3.10 semantics: c >= 0, r carries the sign of f)
HP3D_HD void svd_lartg(float f, float g, float& c, float& s, float& r) {
  if (g == 0.f) { c = 1.f; s = 0.f; r = f; return; }
  if (f == 0.f) { c = 0.f; s = svd_sign(1.f, g); r = fabsf(g); return; }
  float f1 = fabsf(f);
  float d = svd_sqrt(svd_add(svd_mul(f, f), svd_mul(g, g)));
  c = svd_div(f1, d);
  r = svd_sign(d, f);
  s = svd_div(g, r);
}

// slas2: singular values of [[f,g],[0,h]]
HP3D_HD void svd_las2(float f, float g, float h, float& ssmin, float& ssmax) {
  float fa = fabsf(f), ga = fabsf(g), ha = fabsf(h);
  float fhmn = fminf(fa, ha), fhmx = fmaxf(fa, ha);
  if (fhmn == 0.f) {
    ssmin = 0.f;
    if (fhmx == 0.f) ssmax = ga;
    else {
      float mx = fmaxf(fhmx, ga), mn = fminf(fhmx, ga);
      float q = svd_div(mn, mx);
      ssmax = svd_mul(mx, svd_sqrt(svd_add(1.f, svd_mul(q, q))));
    }
  } else if (ga < fhmx) {
    float as = svd_add(1.f, svd_div(fhmn, fhmx));
    float at = svd_div(svd_sub(fhmx, fhmn), fhmx);
    float q = svd_div(ga, fhmx);
    float au = svd_mul(q, q);
    float c = svd_div(2.f, svd_add(svd_sqrt(svd_add(svd_mul(as, as), au)), svd_sqrt(svd_add(svd_mul(at, at), au))));
    ssmin = svd_mul(fhmn, c);
    ssmax = svd_div(fhmx, c);
  } else {
    float au = svd_div(fhmx, ga);
    if (au == 0.f) {
      ssmin = svd_div(svd_mul(fhmn, fhmx), ga);
      ssmax = ga;
    } else {
      float as = svd_add(1.f, svd_div(fhmn, fhmx));
      float at = svd_div(svd_sub(fhmx, fhmn), fhmx);
      float x = svd_mul(as, au), y = svd_mul(at, au);
      float c = svd_div(1.f, svd_add(svd_sqrt(svd_add(1.f, svd_mul(x, x))), svd_sqrt(svd_add(1.f, svd_mul(y, y)))));
      ssmin = svd_mul(svd_mul(fhmn, c), au);
      ssmin = svd_add(ssmin, ssmin);
      ssmax = svd_div(ga, svd_add(c, c));
    }
  }
}

// slasv2: SVD of the 2x2 upper triangular [[f,g],[0,h]]
HP3D_HD void svd_lasv2(float f, float g, float h, float& ssmin, float& ssmax,
                       float& snr, float& csr, float& snl, float& csl) {
  float ft = f, fa = fabsf(ft), ht = h, ha = fabsf(h);
  int pmax = 1;
  bool swp = ha > fa;
  if (swp) { pmax = 3; float t = ft; ft = ht; ht = t; t = fa; fa = ha; ha = t; }
  float gt = g, ga = fabsf(gt);
  float clt, crt, slt, srt;
  if (ga == 0.f) {
    ssmin = ha; ssmax = fa; clt = 1.f; crt = 1.f; slt = 0.f; srt = 0.f;
  } else {
    bool gasmal = true;
    if (ga > fa) {
      pmax = 2;
      if (svd_div(fa, ga) < HP3D_SVD_EPS) {
        gasmal = false;
        ssmax = ga;
        ssmin = (ha > 1.f) ? svd_div(fa, svd_div(ga, ha)) : svd_mul(svd_div(fa, ga), ha);
        clt = 1.f; slt = svd_div(ht, gt); srt = 1.f; crt = svd_div(ft, gt);
      }
    }
    if (gasmal) {
      float d = svd_sub(fa, ha);
      float l = (d == fa) ? 1.f : svd_div(d, fa);
      float m = svd_div(gt, ft);
      float t = svd_sub(2.f, l);
      float mm = svd_mul(m, m), tt = svd_mul(t, t);
      float s = svd_sqrt(svd_add(tt, mm));
      float r = (l == 0.f) ? fabsf(m) : svd_sqrt(svd_add(svd_mul(l, l), mm));
      float a = svd_mul(0.5f, svd_add(s, r));
      ssmin = svd_div(ha, a);
      ssmax = svd_mul(fa, a);
      if (mm == 0.f) {
        if (l == 0.f) t = svd_mul(svd_sign(2.f, ft), svd_sign(1.f, gt));
        else t = svd_add(svd_div(gt, svd_sign(d, ft)), svd_div(m, t));
      } else {
        t = svd_mul(svd_add(svd_div(m, svd_add(s, t)), svd_div(m, svd_add(r, l))), svd_add(1.f, a));
      }
      l = svd_sqrt(svd_add(svd_mul(t, t), 4.f));
      crt = svd_div(2.f, l);
      srt = svd_div(t, l);
      clt = svd_div(svd_add(crt, svd_mul(srt, m)), a);
      slt = svd_div(svd_mul(svd_div(ht, ft), srt), a);
    }
  }
  if (swp) { csl = srt; snl = crt; csr = slt; snr = clt; }
  else     { csl = clt; snl = slt; csr = crt; snr = srt; }
  float tsign;
  if (pmax == 1) tsign = svd_sign(1.f, csr) * svd_sign(1.f, csl) * svd_sign(1.f, f);
  else if (pmax == 2) tsign = svd_sign(1.f, snr) * svd_sign(1.f, csl) * svd_sign(1.f, g);
  else tsign = svd_sign(1.f, snr) * svd_sign(1.f, snl) * svd_sign(1.f, h);
  ssmax = svd_sign(ssmax, tsign);
  ssmin = svd_sign(ssmin, tsign * svd_sign(1.f, f) * svd_sign(1.f, h));
}

// srot on rows (i, i+1) of VT: x' = c x + s y ; y' = c y - s x
HP3D_HD void svd_rot_rows(float* vt, int i, float c, float s) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = vt[i * 3 + k], y = vt[(i + 1) * 3 + k];
    vt[i * 3 + k] = svd_add(svd_mul(c, x), svd_mul(s, y));
    vt[(i + 1) * 3 + k] = svd_sub(svd_mul(c, y), svd_mul(s, x));
  }
}
// srot on columns (j, j+1) of U
HP3D_HD void svd_rot_cols(float* u, int j, float c, float s) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = u[k * 3 + j], y = u[k * 3 + j + 1];
    u[k * 3 + j] = svd_add(svd_mul(c, x), svd_mul(s, y));
    u[k * 3 + j + 1] = svd_sub(svd_mul(c, y), svd_mul(s, x));
  }
}
// slasr plane rotation of rows (j, j+1): A(j+1,:) = c*A(j+1,:) - s*A(j,:); A(j,:) = s*A(j+1,:) + c*A(j,:)
HP3D_HD void svd_lasr_rows(float* a, int j, float c, float s) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float temp = a[(j + 1) * 3 + k], aj = a[j * 3 + k];
    a[(j + 1) * 3 + k] = svd_sub(svd_mul(c, temp), svd_mul(s, aj));
    a[j * 3 + k] = svd_add(svd_mul(s, temp), svd_mul(c, aj));
  }
}
HP3D_HD void svd_lasr_cols(float* a, int j, float c, float s) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float temp = a[k * 3 + j + 1], aj = a[k * 3 + j];
    a[k * 3 + j + 1] = svd_sub(svd_mul(c, temp), svd_mul(s, aj));
    a[k * 3 + j] = svd_add(svd_mul(s, temp), svd_mul(c, aj));
  }
}

// sbdsqr for the 3x3 upper bidiagonal (d[3], e[2]); vt and u (row-major 3x3) receive the rotations.
// Indices are 0-based versions of LAPACK's ll..m (inclusive).
HP3D_HD void svd_bdsqr3(float* d, float* e, float* vt, float* u) {
  const float eps = HP3D_SVD_EPS, unfl = HP3D_SVD_SAFMIN;
  const float tol = 10.f * eps;                 // tolmul = max(10, min(100, eps^-1/8 = 8)) = 10
  const int n = 3, maxitr = 6;
  // thresh from the relative-accuracy estimate of the smallest singular value
  float smax = fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fabsf(d[2]));
  smax = fmaxf(smax, fmaxf(fabsf(e[0]), fabsf(e[1])));
  float sminoa = fabsf(d[0]);
  if (sminoa != 0.f) {
    float mu = sminoa;
    for (int i = 1; i < n; ++i) {
      mu = svd_mul(fabsf(d[i]), svd_div(mu, svd_add(mu, fabsf(e[i - 1]))));
      sminoa = fminf(sminoa, mu);
      if (sminoa == 0.f) break;
    }
  }
  sminoa = svd_div(sminoa, svd_sqrt(3.f));
  const float thresh = fmaxf(svd_mul(tol, sminoa), (float)(maxitr * n * n) * unfl);
  const int maxit = maxitr * n * n;
  int iter = 0, oldll = -2, oldm = -2, m = n - 1, idir = 0;
  float sminl = 0.f;
  while (m > 0) {
    if (iter > maxit) break;
    // find diagonal block ll..m with non-negligible off-diagonals
    smax = fabsf(d[m]);
    int ll = -1;
    bool split = false;
    for (int l = m - 1; l >= 0; --l) {
      float abss = fabsf(d[l]), abse = fabsf(e[l]);
      if (abse <= thresh) { e[l] = 0.f; ll = l; split = true; break; }
      smax = fmaxf(smax, fmaxf(abss, abse));
    }
    if (split) {
      if (ll == m - 1) { m -= 1; continue; }    // bottom singular value converged
      ll = ll + 1;
    } else {
      ll = 0;
    }
    if (ll == m - 1) {                           // 2x2 block
      float sigmn, sigmx, sinr, cosr, sinl, cosl;
      svd_lasv2(d[m - 1], e[m - 1], d[m], sigmn, sigmx, sinr, cosr, sinl, cosl);
      d[m - 1] = sigmx; e[m - 1] = 0.f; d[m] = sigmn;
      svd_rot_rows(vt, m - 1, cosr, sinr);
      svd_rot_cols(u, m - 1, cosl, sinl);
      m -= 2;
      continue;
    }
    // here the block is the full 3x3 (ll = 0, m = 2)
    if (ll > oldm || m < oldll) idir = (fabsf(d[ll]) >= fabsf(d[m])) ? 1 : 2;
    bool deflated = false;
    if (idir == 1) {
      if (fabsf(e[m - 1]) <= svd_mul(tol, fabsf(d[m]))) { e[m - 1] = 0.f; continue; }
      float mu = fabsf(d[ll]);
      sminl = mu;
      for (int l = ll; l < m; ++l) {
        if (fabsf(e[l]) <= svd_mul(tol, mu)) { e[l] = 0.f; deflated = true; break; }
        mu = svd_mul(fabsf(d[l + 1]), svd_div(mu, svd_add(mu, fabsf(e[l]))));
        sminl = fminf(sminl, mu);
      }
    } else {
      if (fabsf(e[ll]) <= svd_mul(tol, fabsf(d[ll]))) { e[ll] = 0.f; continue; }
      float mu = fabsf(d[m]);
      sminl = mu;
      for (int l = m - 1; l >= ll; --l) {
        if (fabsf(e[l]) <= svd_mul(tol, mu)) { e[l] = 0.f; deflated = true; break; }
        mu = svd_mul(fabsf(d[l]), svd_div(mu, svd_add(mu, fabsf(e[l]))));
        sminl = fminf(sminl, mu);
      }
    }
    if (deflated) continue;
    oldll = ll; oldm = m;
    // shift
    float shift, rdum;
    if (svd_mul(svd_mul((float)n, tol), svd_div(sminl, smax)) <= fmaxf(eps, 0.01f * tol)) {
      shift = 0.f;
    } else {
      float sll;
      if (idir == 1) { sll = fabsf(d[ll]); svd_las2(d[m - 1], e[m - 1], d[m], shift, rdum); }
      else           { sll = fabsf(d[m]);  svd_las2(d[ll], e[ll], d[ll + 1], shift, rdum); }
      if (sll > 0.f) { float q = svd_div(shift, sll); if (svd_mul(q, q) < eps) shift = 0.f; }
    }
    iter += m - ll;
    float wc[2], ws[2], wc2[2], ws2[2];
    if (shift == 0.f) {
      if (idir == 1) {
        float cs = 1.f, oldcs = 1.f, sn = 0.f, oldsn = 0.f, r;
        for (int i = ll; i < m; ++i) {
          svd_lartg(svd_mul(d[i], cs), e[i], cs, sn, r);
          if (i > ll) e[i - 1] = svd_mul(oldsn, r);
          svd_lartg(svd_mul(oldcs, r), svd_mul(d[i + 1], sn), oldcs, oldsn, d[i]);
          wc[i - ll] = cs; ws[i - ll] = sn; wc2[i - ll] = oldcs; ws2[i - ll] = oldsn;
        }
        float h = svd_mul(d[m], cs);
        d[m] = svd_mul(h, oldcs);
        e[m - 1] = svd_mul(h, oldsn);
        for (int j = 0; j < m - ll; ++j) svd_lasr_rows(vt, ll + j, wc[j], ws[j]);
        for (int j = 0; j < m - ll; ++j) svd_lasr_cols(u, ll + j, wc2[j], ws2[j]);
        if (fabsf(e[m - 1]) <= thresh) e[m - 1] = 0.f;
      } else {
        float cs = 1.f, oldcs = 1.f, sn = 0.f, oldsn = 0.f, r;
        for (int i = m; i > ll; --i) {
          svd_lartg(svd_mul(d[i], cs), e[i - 1], cs, sn, r);
          if (i < m) e[i] = svd_mul(oldsn, r);
          svd_lartg(svd_mul(oldcs, r), svd_mul(d[i - 1], sn), oldcs, oldsn, d[i]);
          wc[i - ll - 1] = cs; ws[i - ll - 1] = -sn; wc2[i - ll - 1] = oldcs; ws2[i - ll - 1] = -oldsn;
        }
        float h = svd_mul(d[ll], cs);
        d[ll] = svd_mul(h, oldcs);
        e[ll] = svd_mul(h, oldsn);
        for (int j = m - ll - 1; j >= 0; --j) svd_lasr_rows(vt, ll + j, wc2[j], ws2[j]);
        for (int j = m - ll - 1; j >= 0; --j) svd_lasr_cols(u, ll + j, wc[j], ws[j]);
        if (fabsf(e[ll]) <= thresh) e[ll] = 0.f;
      }
    } else {
      if (idir == 1) {
        float f = svd_mul(svd_sub(fabsf(d[ll]), shift), svd_add(svd_sign(1.f, d[ll]), svd_div(shift, d[ll])));
        float g = e[ll];
        float cosr, sinr, cosl, sinl, r;
        for (int i = ll; i < m; ++i) {
          svd_lartg(f, g, cosr, sinr, r);
          if (i > ll) e[i - 1] = r;
          f = svd_add(svd_mul(cosr, d[i]), svd_mul(sinr, e[i]));
          e[i] = svd_sub(svd_mul(cosr, e[i]), svd_mul(sinr, d[i]));
          g = svd_mul(sinr, d[i + 1]);
          d[i + 1] = svd_mul(cosr, d[i + 1]);
          svd_lartg(f, g, cosl, sinl, r);
          d[i] = r;
          f = svd_add(svd_mul(cosl, e[i]), svd_mul(sinl, d[i + 1]));
          d[i + 1] = svd_sub(svd_mul(cosl, d[i + 1]), svd_mul(sinl, e[i]));
          if (i < m - 1) { g = svd_mul(sinl, e[i + 1]); e[i + 1] = svd_mul(cosl, e[i + 1]); }
          wc[i - ll] = cosr; ws[i - ll] = sinr; wc2[i - ll] = cosl; ws2[i - ll] = sinl;
        }
        e[m - 1] = f;
        for (int j = 0; j < m - ll; ++j) svd_lasr_rows(vt, ll + j, wc[j], ws[j]);
        for (int j = 0; j < m - ll; ++j) svd_lasr_cols(u, ll + j, wc2[j], ws2[j]);
        if (fabsf(e[m - 1]) <= thresh) e[m - 1] = 0.f;
      } else {
        float f = svd_mul(svd_sub(fabsf(d[m]), shift), svd_add(svd_sign(1.f, d[m]), svd_div(shift, d[m])));
        float g = e[m - 1];
        float cosr, sinr, cosl, sinl, r;
        for (int i = m; i > ll; --i) {
          svd_lartg(f, g, cosr, sinr, r);
          if (i < m) e[i] = r;
          f = svd_add(svd_mul(cosr, d[i]), svd_mul(sinr, e[i - 1]));
          e[i - 1] = svd_sub(svd_mul(cosr, e[i - 1]), svd_mul(sinr, d[i]));
          g = svd_mul(sinr, d[i - 1]);
          d[i - 1] = svd_mul(cosr, d[i - 1]);
          svd_lartg(f, g, cosl, sinl, r);
          d[i] = r;
          f = svd_add(svd_mul(cosl, e[i - 1]), svd_mul(sinl, d[i - 1]));
          d[i - 1] = svd_sub(svd_mul(cosl, d[i - 1]), svd_mul(sinl, e[i - 1]));
          if (i > ll + 1) { g = svd_mul(sinl, e[i - 2]); e[i - 2] = svd_mul(cosl, e[i - 2]); }
          wc[i - ll - 1] = cosr; ws[i - ll - 1] = -sinr; wc2[i - ll - 1] = cosl; ws2[i - ll - 1] = -sinl;
        }
        e[ll] = f;
        if (fabsf(e[ll]) <= thresh) e[ll] = 0.f;
        for (int j = m - ll - 1; j >= 0; --j) svd_lasr_rows(vt, ll + j, wc2[j], ws2[j]);
        for (int j = m - ll - 1; j >= 0; --j) svd_lasr_cols(u, ll + j, wc[j], ws[j]);
      }
    }
  }
  // make singular values positive (negate rows of VT)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (d[i] < 0.f || (d[i] == 0.f && signbit(d[i]))) {
      d[i] = -d[i];
      vt[i * 3 + 0] = -vt[i * 3 + 0]; vt[i * 3 + 1] = -vt[i * 3 + 1]; vt[i * 3 + 2] = -vt[i * 3 + 2];
    }
  }
  // sort into decreasing order: smallest of the leading part goes to the end (one swap per pass)
  for (int i = 0; i < n - 1; ++i) {
    int isub = 0; float smin = d[0];
    for (int j = 1; j < n - i; ++j) if (d[j] <= smin) { isub = j; smin = d[j]; }
    int last = n - 1 - i;
    if (isub != last) {
      d[isub] = d[last]; d[last] = smin;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float t = vt[isub * 3 + k]; vt[isub * 3 + k] = vt[last * 3 + k]; vt[last * 3 + k] = t;
        t = u[k * 3 + isub]; u[k * 3 + isub] = u[k * 3 + last]; u[k * 3 + last] = t;
      }
    }
  }
}

// slarfg for a length-n vector (alpha, x[0..n-2]); returns tau, overwrites alpha with beta, x with v.
HP3D_HD float svd_larfg2(float& alpha, float& x0) {            // n = 2
  float xnorm = fabsf(x0);
  if (xnorm == 0.f) return 0.f;
  float beta = -svd_sign(svd_lapy2(alpha, xnorm), alpha);
  float tau = svd_div(svd_sub(beta, alpha), beta);
  x0 = svd_mul(x0, svd_div(1.f, svd_sub(alpha, beta)));
  alpha = beta;
  return tau;
}
HP3D_HD float svd_nrm2_2(float a, float b) {
  // scaled 2-norm as reference BLAS snrm2 does (scale/ssq form)
  float aa = fabsf(a), ab = fabsf(b);
  float scale = fmaxf(aa, ab);
  if (scale == 0.f) return 0.f;
  float qa = svd_div(aa, scale), qb = svd_div(ab, scale);
  return svd_mul(scale, svd_sqrt(svd_add(svd_mul(qa, qa), svd_mul(qb, qb))));
}
HP3D_HD float svd_larfg3(float& alpha, float& x0, float& x1) {  // n = 3
  float xnorm = svd_nrm2_2(x0, x1);
  if (xnorm == 0.f) return 0.f;
  float beta = -svd_sign(svd_lapy2(alpha, xnorm), alpha);
  float tau = svd_div(svd_sub(beta, alpha), beta);
  float sc = svd_div(1.f, svd_sub(alpha, beta));
  x0 = svd_mul(x0, sc); x1 = svd_mul(x1, sc);
  alpha = beta;
  return tau;
}

// Full SVD A = U diag(S) V^T of a row-major 3x3 float matrix, LAPACK-convention signs/order.
// U, V row-major 3x3 (V, not V^T, like torch.svd); S descending non-negative.
HP3D_HD void svd3_lapack(const float* A, float* U, float* S, float* V) {
  float a00 = A[0], a01 = A[1], a02 = A[2], a10 = A[3], a11 = A[4], a12 = A[5], a20 = A[6], a21 = A[7], a22 = A[8];
  float d[3], e[2];
  // H1 annihilates a10, a20
  float v1a = a10, v1b = a20;
  float tauq1 = svd_larfg3(a00, v1a, v1b);
  d[0] = a00;
  if (tauq1 != 0.f) {   // apply H1 = I - tau v v^T (v = [1, v1a, v1b]) to columns 1, 2
    float w1 = svd_add(svd_add(a01, svd_mul(v1a, a11)), svd_mul(v1b, a21));
    float w2 = svd_add(svd_add(a02, svd_mul(v1a, a12)), svd_mul(v1b, a22));
    float t1 = svd_mul(tauq1, w1), t2 = svd_mul(tauq1, w2);
    a01 = svd_sub(a01, t1); a11 = svd_sub(a11, svd_mul(v1a, t1)); a21 = svd_sub(a21, svd_mul(v1b, t1));
    a02 = svd_sub(a02, t2); a12 = svd_sub(a12, svd_mul(v1a, t2)); a22 = svd_sub(a22, svd_mul(v1b, t2));
  }
  // G1 annihilates a02 (row 0, columns 1:2)
  float g1 = a02;
  float taup1 = svd_larfg2(a01, g1);
  e[0] = a01;
  if (taup1 != 0.f) {   // apply G1 = I - tau u u^T (u = [1, g1]) from the right to rows 1, 2
    float w1 = svd_add(a11, svd_mul(g1, a12));
    float w2 = svd_add(a21, svd_mul(g1, a22));
    float t1 = svd_mul(taup1, w1), t2 = svd_mul(taup1, w2);
    a11 = svd_sub(a11, t1); a12 = svd_sub(a12, svd_mul(g1, t1));
    a21 = svd_sub(a21, t2); a22 = svd_sub(a22, svd_mul(g1, t2));
  }
  // H2 annihilates a21
  float v2 = a21;
  float tauq2 = svd_larfg2(a11, v2);
  d[1] = a11;
  if (tauq2 != 0.f) {   // apply H2 (v = [1, v2]) to column 2, rows 1:2
    float w = svd_add(a12, svd_mul(v2, a22));
    float t = svd_mul(tauq2, w);
    a12 = svd_sub(a12, t); a22 = svd_sub(a22, svd_mul(v2, t));
  }
  e[1] = a12;
  d[2] = a22;
  float ub[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  float vt[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  svd_bdsqr3(d, e, vt, ub);
  // U = H1 H2 U_b : apply H2 to rows 1:2, then H1 to rows 0:2 (sorm2r order for 'L','N')
  if (tauq2 != 0.f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float w = svd_add(ub[3 + c], svd_mul(v2, ub[6 + c]));
      float t = svd_mul(tauq2, w);
      ub[3 + c] = svd_sub(ub[3 + c], t); ub[6 + c] = svd_sub(ub[6 + c], svd_mul(v2, t));
    }
  }
  if (tauq1 != 0.f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float w = svd_add(svd_add(ub[c], svd_mul(v1a, ub[3 + c])), svd_mul(v1b, ub[6 + c]));
      float t = svd_mul(tauq1, w);
      ub[c] = svd_sub(ub[c], t); ub[3 + c] = svd_sub(ub[3 + c], svd_mul(v1a, t)); ub[6 + c] = svd_sub(ub[6 + c], svd_mul(v1b, t));
    }
  }
  // VT = VT_b G1 : columns 1:2 of every row
  if (taup1 != 0.f) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float w = svd_add(vt[r * 3 + 1], svd_mul(g1, vt[r * 3 + 2]));
      float t = svd_mul(taup1, w);
      vt[r * 3 + 1] = svd_sub(vt[r * 3 + 1], t); vt[r * 3 + 2] = svd_sub(vt[r * 3 + 2], svd_mul(g1, t));
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) U[i] = ub[i];
  S[0] = d[0]; S[1] = d[1]; S[2] = d[2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) V[c * 3 + r] = vt[r * 3 + c];
}

HP3D_HD float det3(const float* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

}  // namespace hp3d
