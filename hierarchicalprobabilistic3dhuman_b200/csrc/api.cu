// libhp3d: version + thread-local error plumbing (C ABI declared in include/hp3d.h).
#include "common.cuh"
#include <stdarg.h>

namespace hp3d {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  cudaGetLastError();   // clear the sticky launch error so later calls report their own status
  return (int)e;
}
}  // namespace hp3d

extern "C" int hp3d_version(void) { return HP3D_VERSION; }
extern "C" const char* hp3d_last_error(void) { return hp3d::g_err; }
