// Matrix-Fisher normalising constant for sm_100a: log c(S) and d log c / d s for a batch of
// proper singular values -- the "Bessel" term of the reference's pose NLL (losses/matrix_fisher_loss.py:134-192,
// `LogMFNormConstant`), and E[R] = U diag(d log c / d s) V^T of the matrix-Fisher distribution the sampler draws from.
//
// One warp per (image x joint) row: the 512 trapezoid nodes of the four integrals (c_bar and its three derivatives) are
// split over the lanes (16 nodes each, 2 Bessel polynomials + 1 exp per integrand), partial sums meet in a shuffle
// reduction. Pure ALU/SFU work: 4 x 512 integrand evaluations (~120 kFLOP) per row against 12 B in / 16 B out.
// The arithmetic lives in mf_norm_math.h, which is also compiled for the host and checked against the reference-pinned
// oracle (tests/test_mf_norm_host.py).
// STATUS: compiled for sm_100a, arithmetic verified on the host; not yet run on hardware (round 1 ended with the GPU
// budget spent) -- tests/test_gpu_mf_norm.py is opt-in until then.
#include "common.cuh"
#include "mf_norm_math.h"

using namespace hp3d;

namespace {

__global__ void __launch_bounds__(256) mf_log_norm_kernel(const float* __restrict__ S, int n, float* __restrict__ log_c,
                                                          float* __restrict__ dlogc_ds) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;                                   // warp-uniform
  const float s0 = S[3 * row], s1 = S[3 * row + 1], s2 = S[3 * row + 2];
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = lane; i < MF_NORM_TRAPS; i += 32) {
    float t[4];
    mf_norm_node_terms(i, s0, s1, s2, t);
#pragma unroll
    for (int q = 0; q < 4; ++q) sum[q] += t[q];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int q = 0; q < 4; ++q) sum[q] += __shfl_xor_sync(0xffffffffu, sum[q], o);
  if (lane == 0) {
    float lc, d[3];
    mf_norm_finish(sum, s0, s1, s2, &lc, d);
    log_c[row] = lc;
    if (dlogc_ds) { dlogc_ds[3 * row] = d[0]; dlogc_ds[3 * row + 1] = d[1]; dlogc_ds[3 * row + 2] = d[2]; }
  }
}

}  // namespace

extern "C" int hp3d_mf_log_norm_constant(const float* S_proper, int n, float* log_c, float* dlogc_ds, void* stream) {
  HP3D_ARG(S_proper && log_c && n > 0, "bad argument");
  mf_log_norm_kernel<<<cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(S_proper, n, log_c, dlogc_ds);
  return launch_status("mf_log_norm_kernel");
}
