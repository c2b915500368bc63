// Temporary stubs until the tcgen05 kernels land.
#include "encoder.cuh"
namespace hp3d {
int encoder_tc_create(const hp3d_encoder_weights*, void**) { set_error("HP3D_ENC_FAST not built yet"); return -2; }
void encoder_tc_destroy(void*) {}
size_t encoder_tc_workspace_bytes(const void*, int, int, int) { return 0; }
int encoder_tc_forward(const void*, const float*, int, int, int, float*, void*, size_t, cudaStream_t) { return -2; }
int blend_tc_create(const double*, void** out) { *out = nullptr; return 0; }
void blend_tc_destroy(void*) {}
int blend_tc_forward(void*, const float*, int, const float*, int, float*, cudaStream_t) { return -2; }
}
