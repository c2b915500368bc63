// Host build of the shared host/device math headers so the numerics that cannot be exercised
// without a GPU (LAPACK-convention 3x3 SVD) are checked on CPU by `-m "not gpu"` tests.
// Test shim only; the product path never calls it.
#include "svd3.h"
extern "C" void hp3d_host_svd3(const float* A, long n, float* U, float* S, float* V) {
  for (long i = 0; i < n; ++i) hp3d::svd3_lapack(A + 9 * i, U + 9 * i, S + 3 * i, V + 9 * i);
}
