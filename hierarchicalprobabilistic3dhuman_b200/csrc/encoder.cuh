// Internal interface between the encoder front (encoder.cu) and the tensor-core plan (conv_tc.cu).
#pragma once
#include "common.cuh"
#include <vector>
namespace hp3d {
int fold_conv_bn(const hp3d_conv_bn& c, float eps, int cin_pad, std::vector<float>& w_khwc, std::vector<float>& bias);
int encoder_tc_create(const hp3d_encoder_weights* w, bool split, void** out);
void encoder_tc_destroy(void* p);
size_t encoder_tc_workspace_bytes(const void* p, int B, int H, int W);
// image-space input of the fused proxy-representation producer (proxy.cu); when given, x_nchw is ignored
struct ImageInput {
  const float* rgb; const float* joints2d; const unsigned char* visibility;
  float gaussian_std; int gaussian_size; float threshold; int nms; float heat_std;
};
// optional by-product of the input cast: arg-max pixel / visibility of the 17 joint heat-maps (channels 1..17)
struct ArgmaxOut { float eps; float* joints2d_px; int* vis; };
int encoder_tc_forward(const void* p, const float* x_nchw, int B, int H, int W, float* feats, void* workspace,
                       size_t workspace_bytes, float* taps, cudaStream_t stream, const ImageInput* image = nullptr,
                       const ArgmaxOut* argmax = nullptr, bool x_half = false);
// rank.cu: stand-alone heat-map arg-max (17 maps per image, images `image_stride` floats apart)
int heatmap_argmax(const float* heatmaps, long long image_stride, int B, int H, int W, float eps, float* joints2d_px,
                   int* vis, cudaStream_t stream);
int proxy_rep_nhwc_f16(const float* rgb, const float* joints2d, const unsigned char* visibility, int B, int img_wh,
                       float gaussian_std, int gaussian_size, float threshold, int nms, float heat_std, void* nhwc32,
                       int split, cudaStream_t stream);
// append `count` activation values (converted to fp32) to the debug tap buffer
int tap_copy_f32(const float* src, size_t count, float** taps, cudaStream_t s);
int tap_copy_f16(const void* src, size_t count, float** taps, cudaStream_t s);
}  // namespace hp3d
