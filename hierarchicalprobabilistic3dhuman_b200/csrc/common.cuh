// Shared helpers for libhp3d (error plumbing, constants, small device math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "../../include/hp3d.h"

namespace hp3d {

constexpr int NV = HP3D_NUM_VERTS;          // 6890
constexpr int NV3 = NV * 3;                 // 20670
constexpr int VPITCH = 20672;               // v_shaped row pitch (16-byte multiple)
constexpr int NJ = HP3D_NUM_JOINTS;         // 24
constexpr int NBJ = HP3D_NUM_BODY_JOINTS;   // 23
constexpr int NBETA = HP3D_NUM_BETAS;       // 10
constexpr int NPF = 207;                    // pose feature length
constexpr int NOUTJ = HP3D_NUM_OUT_JOINTS;  // 90
constexpr int NPICK = 21;
constexpr int NREG = 45;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define HP3D_CUDA(call)                                        \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return ::hp3d::cuda_fail(e__, #call); \
  } while (0)

#define HP3D_ARG(cond, msg)                                    \
  do {                                                         \
    if (!(cond)) { ::hp3d::set_error("%s: %s", __func__, msg); return -1; } \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return 0;
}

template <typename T>
int upload(T** dptr, const T* host, size_t n) {
  cudaError_t e = cudaMalloc((void**)dptr, n * sizeof(T));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  e = cudaMemcpy(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy H2D");
  return 0;
}

// Opt a kernel in to more than 48 KB of dynamic shared memory. The attribute is per DEVICE (context), so the "already done"
// cache is a bit per device ordinal, not a process-wide flag: one process may drive several GPUs (handles and workspaces
// are keyed by device on the Python side). A lost update under a race only repeats the (idempotent) call.
#define HP3D_SMEM_OPT_IN(kernel, bytes)                                                                        \
  do {                                                                                                         \
    static unsigned long long done__ = 0ull;                                                                   \
    int dev__ = 0;                                                                                             \
    HP3D_CUDA(cudaGetDevice(&dev__));                                                                          \
    const unsigned long long bit__ = 1ull << (dev__ & 63);                                                     \
    if (!(__atomic_load_n(&done__, __ATOMIC_RELAXED) & bit__)) {                                               \
      HP3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      __atomic_fetch_or(&done__, bit__, __ATOMIC_RELAXED);                                                     \
    }                                                                                                          \
  } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
// CTAs a persistent one-CTA-per-SM kernel may launch: the SM count, or HP3D_SM_LIMIT if set lower (multi-GPU runs leave a
// few SMs to the NCCL all-gather kernels, which otherwise cannot start until a whole persistent kernel has drained and
// then stall a statically partitioned successor)
inline int persistent_ctas(int num_sms) {
  const char* e = getenv("HP3D_SM_LIMIT");
  const int lim = e ? atoi(e) : 0;
  return (lim > 0 && lim < num_sms) ? lim : num_sms;
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace hp3d
