// Temporary stub until the tcgen05 pose-blend GEMM lands.
#include "common.cuh"
namespace hp3d {
int blend_tc_create(const double*, void** out) { *out = nullptr; return 0; }
void blend_tc_destroy(void*) {}
int blend_tc_forward(void*, const float*, int, const float*, int, float*, cudaStream_t) { return -2; }
}
