// ResNet-18 proxy-representation encoder for sm_100a (eval-mode BatchNorm folded into the convs).
//
// Replaces reference models/resnet.py:202-217 (+ BasicBlock.forward :62-78) for resnet18(18).
// Three arithmetic modes behind one handle:
//   HP3D_ENC_SPLIT   tcgen05 tensor-core implicit GEMM on fp16 hi/lo pairs, three products per k-block accumulated in
//                    fp32 TMEM (conv_tc.cu) -- the default: meets the <=1e-4 contract on the tensor cores;
//   HP3D_ENC_FAST    single fp16 product per k-block (conv_tc.cu) -- opt-in, ~3e-4 on the features;
//   HP3D_ENC_PARITY  fp32 NHWC activations, fp32 CUDA-core implicit GEMM (this file) -- the plain-fp32 cross-check.
// Layer plan (18x256x256 input): stem 7x7/2 -> maxpool 3x3/2 -> 4 stages x 2 BasicBlocks -> global
// average pool -> (B,512). Layout in HBM: activations NHWC (channels innermost, stem input padded
// 18 -> 20/32 channels), weights [kh][kw][cin][cout] so both GEMM operands are contiguous along K/N.
#include "common.cuh"
#include "encoder.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <math.h>

using namespace hp3d;

namespace {

// ---------------------------------------------------------------- layout change NCHW fp32 -> NHWC
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int HW, int Cp,
                                                           T* __restrict__ y) {
  // one CTA: 32 pixels x all channels of one image, transposed through shared memory
  __shared__ float tile[32][33];
  const int n = blockIdx.y, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows of 32
  for (int c = ty; c < 32; c += 8) tile[c][tx] = (c < C && p0 + tx < HW) ? x[((size_t)n * C + c) * HW + p0 + tx] : 0.f;
  __syncthreads();
  for (int p = ty; p < 32; p += 8)
    if (tx < Cp && p0 + p < HW) y[((size_t)n * HW + p0 + p) * Cp + tx] = (T)tile[tx][p];
}

// ---------------------------------------------------------------- fp32 implicit-GEMM convolution
// out[p][co] = relu?( sum_{kh,kw,ci} in[n][oy*s-pad+kh][ox*s-pad+kw][ci] * w[kh][kw][ci][co] + bias[co] (+ res[p][co]) )
// CTA tile 64 output pixels x 64 output channels, 256 threads x (4 px x 4 co), K chunk = 16 input channels.
struct ConvGeom { int H, W, Cin, Ho, Wo, Cout, k, stride, pad; };

__global__ void __launch_bounds__(256) conv_fp32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ res,
                                                        float* __restrict__ out, ConvGeom g, int P, int relu) {
  __shared__ float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int p0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  // A-load role: pixel tid/4, channel quad tid%4
  const int lp = p0 + (tid >> 2);
  int ln = 0, loy = 0, lox = 0;
  const bool lvalid = lp < P;
  if (lvalid) { ln = lp / (g.Ho * g.Wo); const int r = lp - ln * g.Ho * g.Wo; loy = r / g.Wo; lox = r - loy * g.Wo; }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int kh = 0; kh < g.k; ++kh) {
    for (int kw = 0; kw < g.k; ++kw) {
      const int iy = loy * g.stride - g.pad + kh, ix = lox * g.stride - g.pad + kw;
      const bool inb = lvalid && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      const float* src = in + (((size_t)ln * g.H + iy) * g.W + ix) * g.Cin;
      const float* wt = w + (size_t)(kh * g.k + kw) * g.Cin * g.Cout;
      for (int c0 = 0; c0 < g.Cin; c0 += 16) {
        {
          const int c = c0 + (tid & 3) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (inb && c < g.Cin) v = *reinterpret_cast<const float4*>(src + c);
          const int kk = (tid & 3) * 4, px = tid >> 2;
          As[kk][px] = v.x; As[kk + 1][px] = v.y; As[kk + 2][px] = v.z; As[kk + 3][px] = v.w;
        }
        {
          const int kk = tid >> 4, c = c0 + kk;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < g.Cin) v = *reinterpret_cast<const float4*>(wt + (size_t)c * g.Cout + n0 + (tid & 15) * 4);
          *reinterpret_cast<float4*>(&Bs[kk][(tid & 15) * 4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
  const float4 bz = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty * 4 + i;
    if (p >= P) continue;
    float4 o = make_float4(acc[i][0] + bz.x, acc[i][1] + bz.y, acc[i][2] + bz.z, acc[i][3] + bz.w);
    const size_t off = (size_t)p * g.Cout + n0 + tx * 4;
    if (res) { const float4 r = *reinterpret_cast<const float4*>(res + off); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4*>(out + off) = o;
  }
}

// ---------------------------------------------------------------- maxpool 3x3 s2 p1 (NHWC), avgpool
template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const T* __restrict__ in, int H, int W, int C, int Ho, int Wo,
                                                           T* __restrict__ out, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  size_t r = i / C;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int n = (int)(r / Ho);
  float m = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= W) continue;
      m = fmaxf(m, (float)in[(((size_t)n * H + iy) * W + ix) * C + c]);
    }
  }
  out[i] = (T)m;
}

template <typename T>
__global__ void __launch_bounds__(256) avgpool_kernel(const T* __restrict__ in, int HW, int C, float* __restrict__ out) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += (float)in[((size_t)n * HW + p) * C + c];
    out[(size_t)n * C + c] = s / (float)HW;
  }
}

template <typename T>
__global__ void tap_copy_kernel(const T* __restrict__ src, size_t n, float* __restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}

}  // namespace

namespace hp3d {

int tap_copy_f32(const float* src, size_t count, float** taps, cudaStream_t s) {
  if (!*taps) return 0;
  tap_copy_kernel<float><<<1184, 256, 0, s>>>(src, count, *taps);
  *taps += count;
  return launch_status("tap_copy_kernel");
}
int tap_copy_f16(const void* src, size_t count, float** taps, cudaStream_t s) {
  if (!*taps) return 0;
  tap_copy_kernel<__half><<<1184, 256, 0, s>>>((const __half*)src, count, *taps);
  *taps += count;
  return launch_status("tap_copy_kernel");
}

int fold_conv_bn(const hp3d_conv_bn& c, float eps, int cin_pad, std::vector<float>& w_khwc, std::vector<float>& bias) {
  if (!c.w || !c.bn_w || !c.bn_b || !c.bn_mean || !c.bn_var) { set_error("encoder: null conv/bn pointer"); return -1; }
  const int k = c.k;
  w_khwc.assign((size_t)k * k * cin_pad * c.cout, 0.f);
  bias.assign(c.cout, 0.f);
  for (int o = 0; o < c.cout; ++o) {
    const double scale = (double)c.bn_w[o] / sqrt((double)c.bn_var[o] + (double)eps);
    bias[o] = (float)((double)c.bn_b[o] - (double)c.bn_mean[o] * scale);
    for (int i = 0; i < c.cin; ++i)
      for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b)
          w_khwc[(((size_t)a * k + b) * cin_pad + i) * c.cout + o] =
              (float)((double)c.w[(((size_t)o * c.cin + i) * k + a) * k + b] * scale);
  }
  return 0;
}

}  // namespace hp3d

struct ConvLayer {
  float *w = nullptr, *bias = nullptr;
  int cin, cin_pad, cout, k, stride, pad;
};

struct hp3d_encoder {
  int mode = 0;
  ConvLayer stem, conv[4][2][2], down[4];
  bool has_down[4] = {false, false, false, false};
  void* tc = nullptr;    // tensor-core plan (conv_tc.cu)
};

static int make_layer(const hp3d_conv_bn& c, float eps, int cin_pad, ConvLayer& L) {
  std::vector<float> w, b;
  int rc = fold_conv_bn(c, eps, cin_pad, w, b);
  if (rc) return rc;
  L.cin = c.cin; L.cin_pad = cin_pad; L.cout = c.cout; L.k = c.k; L.stride = c.stride; L.pad = c.pad;
  rc = upload(&L.w, w.data(), w.size());
  return rc ? rc : upload(&L.bias, b.data(), b.size());
}

extern "C" int hp3d_encoder_create(const hp3d_encoder_weights* w, int mode, hp3d_encoder** out) {
  HP3D_ARG(w && out, "null argument");
  HP3D_ARG(mode == HP3D_ENC_PARITY || mode == HP3D_ENC_FAST || mode == HP3D_ENC_SPLIT, "unknown mode");
  HP3D_ARG(w->stem.cin == 18 && w->stem.cout == 64 && w->stem.k == 7 && w->stem.stride == 2 && w->stem.pad == 3,
           "stem must be 7x7/2 pad 3, 18->64");
  hp3d_encoder* h = new hp3d_encoder();
  h->mode = mode;
  int rc = 0;
  if (mode != HP3D_ENC_PARITY) {
    rc = encoder_tc_create(w, mode == HP3D_ENC_SPLIT, &h->tc);
  } else {
    rc = make_layer(w->stem, w->bn_eps, 20, h->stem);
    const int planes[4] = {64, 128, 256, 512};
    int inpl = 64;
    for (int l = 0; l < 4 && !rc; ++l) {
      for (int b = 0; b < 2 && !rc; ++b) {
        const hp3d_conv_bn& c1 = w->conv[l][b][0];
        const hp3d_conv_bn& c2 = w->conv[l][b][1];
        const int stride = (l > 0 && b == 0) ? 2 : 1;
        if (c1.cin != inpl || c1.cout != planes[l] || c1.k != 3 || c1.stride != stride || c1.pad != 1 ||
            c2.cin != planes[l] || c2.cout != planes[l] || c2.k != 3 || c2.stride != 1 || c2.pad != 1) {
          set_error("hp3d_encoder_create: layer%d.%d is not a ResNet-18 BasicBlock", l + 1, b); rc = -1; break;
        }
        rc = make_layer(c1, w->bn_eps, c1.cin, h->conv[l][b][0]);
        rc = rc ? rc : make_layer(c2, w->bn_eps, c2.cin, h->conv[l][b][1]);
        if (b == 0 && l > 0) {
          const hp3d_conv_bn& d = w->down[l];
          if (!d.w || d.cin != inpl || d.cout != planes[l] || d.k != 1 || d.stride != 2 || d.pad != 0) {
            set_error("hp3d_encoder_create: layer%d downsample must be 1x1/2", l + 1); rc = -1; break;
          }
          rc = rc ? rc : make_layer(d, w->bn_eps, d.cin, h->down[l]);
          h->has_down[l] = true;
        }
        inpl = planes[l];
      }
    }
  }
  if (rc) { hp3d_encoder_destroy(h); return rc; }
  *out = h;
  return 0;
}

extern "C" void hp3d_encoder_destroy(hp3d_encoder* h) {
  if (!h) return;
  auto fr = [](ConvLayer& L) { cudaFree(L.w); cudaFree(L.bias); };
  fr(h->stem);
  for (int l = 0; l < 4; ++l) { for (int b = 0; b < 2; ++b) { fr(h->conv[l][b][0]); fr(h->conv[l][b][1]); } fr(h->down[l]); }
  if (h->tc) encoder_tc_destroy(h->tc);
  delete h;
}

static size_t enc_ws_parity(int B, int H, int W) {
  const size_t in = align_up((size_t)B * H * W * 20 * 4, 256);
  const size_t stem = align_up((size_t)B * (H / 2) * (W / 2) * 64 * 4, 256);
  const size_t act = align_up((size_t)B * (H / 4) * (W / 4) * 64 * 4, 256);   // largest block activation
  return in + stem + 4 * act;
}

extern "C" size_t hp3d_encoder_workspace_bytes(const hp3d_encoder* h, int B, int H, int W) {
  if (!h || B <= 0) return 0;
  if (h->mode != HP3D_ENC_PARITY) return encoder_tc_workspace_bytes(h->tc, B, H, W);
  return enc_ws_parity(B, H, W);
}

static int run_conv(const ConvLayer& L, const float* in, int B, int H, int W, const float* res, int relu, float* out,
                    cudaStream_t s) {
  ConvGeom g;
  g.H = H; g.W = W; g.Cin = L.cin_pad; g.k = L.k; g.stride = L.stride; g.pad = L.pad; g.Cout = L.cout;
  g.Ho = (H + 2 * L.pad - L.k) / L.stride + 1;
  g.Wo = (W + 2 * L.pad - L.k) / L.stride + 1;
  const int P = B * g.Ho * g.Wo;
  dim3 grid(cdiv(P, 64), L.cout / 64);
  conv_fp32_kernel<<<grid, 256, 0, s>>>(in, L.w, L.bias, res, out, g, P, relu);
  return launch_status("conv_fp32_kernel");
}

extern "C" int hp3d_encoder_forward(const hp3d_encoder* h, const float* x, int B, int H, int W, float* feats,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  return hp3d_encoder_forward_taps(h, x, B, H, W, feats, workspace, workspace_bytes, nullptr, stream_);
}

extern "C" int hp3d_encoder_forward_image(const hp3d_encoder* h, const float* rgb, const float* joints2d,
                                          const unsigned char* visibility, int B, int img_wh, float gaussian_std,
                                          int gaussian_size, float threshold, int nms, float heat_std, float* feats,
                                          void* workspace, size_t workspace_bytes, void* stream_) {
  HP3D_ARG(h && rgb && joints2d && feats && workspace, "null argument");
  HP3D_ARG(h->mode != HP3D_ENC_PARITY, "fused image input needs a tensor-core handle (HP3D_ENC_SPLIT / HP3D_ENC_FAST); HP3D_ENC_PARITY: hp3d_proxy_rep + hp3d_encoder_forward");
  HP3D_ARG(B > 0 && img_wh == 256, "256x256 proxy representations (DATA.PROXY_REP_SIZE)");
  HP3D_ARG(workspace_bytes >= hp3d_encoder_workspace_bytes(h, B, img_wh, img_wh), "workspace too small");
  const ImageInput im = {rgb, joints2d, visibility, gaussian_std, gaussian_size, threshold, nms, heat_std};
  return encoder_tc_forward(h->tc, nullptr, B, img_wh, img_wh, feats, workspace, workspace_bytes, nullptr, (cudaStream_t)stream_, &im);
}

extern "C" int hp3d_encoder_forward_argmax(const hp3d_encoder* h, const float* x, int B, int H, int W, float* feats,
                                           void* workspace, size_t workspace_bytes, float eps, float* joints2d_px,
                                           int32_t* vis, void* stream_) {
  HP3D_ARG(h && x && feats && workspace && joints2d_px && vis, "null argument");
  if (h->mode != HP3D_ENC_PARITY) {
    HP3D_ARG(B > 0 && H >= 32 && W >= 32 && H % 32 == 0 && W % 32 == 0, "H and W must be multiples of 32");
    HP3D_ARG(workspace_bytes >= hp3d_encoder_workspace_bytes(h, B, H, W), "workspace too small");
    const ArgmaxOut am = {eps, joints2d_px, vis};
    return encoder_tc_forward(h->tc, x, B, H, W, feats, workspace, workspace_bytes, nullptr, (cudaStream_t)stream_, nullptr, &am);
  }
  int rc = heatmap_argmax(x + (size_t)H * W, (long long)18 * H * W, B, H, W, eps, joints2d_px, vis, (cudaStream_t)stream_);
  if (rc) return rc;
  return hp3d_encoder_forward_taps(h, x, B, H, W, feats, workspace, workspace_bytes, nullptr, stream_);
}

extern "C" int hp3d_encoder_forward_f16in(const hp3d_encoder* h, const void* x_f16, int B, int H, int W, float* feats,
                                          void* workspace, size_t workspace_bytes, float eps, float* joints2d_px,
                                          int32_t* vis, void* stream_) {
  HP3D_ARG(h && x_f16 && feats && workspace, "null argument");
  HP3D_ARG(h->mode != HP3D_ENC_PARITY, "fp16 input needs a tensor-core handle (HP3D_ENC_SPLIT / HP3D_ENC_FAST)");
  HP3D_ARG((joints2d_px == nullptr) == (vis == nullptr), "joints2d_px and vis must be given together");
  HP3D_ARG(B > 0 && H >= 32 && W >= 32 && H % 32 == 0 && W % 32 == 0, "H and W must be multiples of 32");
  HP3D_ARG(workspace_bytes >= hp3d_encoder_workspace_bytes(h, B, H, W), "workspace too small");
  const ArgmaxOut am = {eps, joints2d_px, vis};
  return encoder_tc_forward(h->tc, (const float*)x_f16, B, H, W, feats, workspace, workspace_bytes, nullptr, (cudaStream_t)stream_,
                            nullptr, joints2d_px ? &am : nullptr, true);
}

extern "C" int hp3d_encoder_forward_taps(const hp3d_encoder* h, const float* x, int B, int H, int W, float* feats,
                                         void* workspace, size_t workspace_bytes, float* taps, void* stream_) {
  HP3D_ARG(h && x && feats && workspace, "null argument");
  HP3D_ARG(B > 0 && H >= 32 && W >= 32 && H % 32 == 0 && W % 32 == 0, "H and W must be multiples of 32");
  HP3D_ARG(workspace_bytes >= hp3d_encoder_workspace_bytes(h, B, H, W), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  if (h->mode != HP3D_ENC_PARITY) return encoder_tc_forward(h->tc, x, B, H, W, feats, workspace, workspace_bytes, taps, s);
  char* ws = (char*)workspace;
  float* xin = (float*)ws; ws += align_up((size_t)B * H * W * 20 * 4, 256);
  float* stem = (float*)ws; ws += align_up((size_t)B * (H / 2) * (W / 2) * 64 * 4, 256);
  const size_t act = align_up((size_t)B * (H / 4) * (W / 4) * 64 * 4, 256);
  float* buf[4];
  for (int i = 0; i < 4; ++i) { buf[i] = (float*)ws; ws += act; }
  nchw_to_nhwc_kernel<float><<<dim3(cdiv(H * W, 32), B), 256, 0, s>>>(x, 18, H * W, 20, xin);
  int rc = launch_status("nchw_to_nhwc_kernel");
  if (rc) return rc;
  rc = run_conv(h->stem, xin, B, H, W, nullptr, 1, stem, s);
  if (rc) return rc;
  int ch = H / 2, cw = W / 2;
  rc = tap_copy_f32(stem, (size_t)B * ch * cw * 64, &taps, s);
  if (rc) return rc;
  {
    const size_t total = (size_t)B * (ch / 2) * (cw / 2) * 64;
    maxpool3x3s2_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(stem, ch, cw, 64, ch / 2, cw / 2, buf[0], total);
    rc = launch_status("maxpool3x3s2_kernel");
    if (rc) return rc;
    ch /= 2; cw /= 2;
    rc = tap_copy_f32(buf[0], total, &taps, s);
    if (rc) return rc;
  }
  float* cur = buf[0];
  int free_idx[3] = {1, 2, 3};
  int C = 64;
  for (int l = 0; l < 4; ++l) {
    for (int b = 0; b < 2; ++b) {
      const ConvLayer& c1 = h->conv[l][b][0];
      const ConvLayer& c2 = h->conv[l][b][1];
      float* t = buf[free_idx[0]];
      float* y = buf[free_idx[1]];
      float* d = buf[free_idx[2]];
      rc = run_conv(c1, cur, B, ch, cw, nullptr, 1, t, s);
      if (rc) return rc;
      const int oh = ch / c1.stride, ow = cw / c1.stride;
      const float* identity = cur;
      if (b == 0 && h->has_down[l]) {
        rc = run_conv(h->down[l], cur, B, ch, cw, nullptr, 0, d, s);
        if (rc) return rc;
        identity = d;
      }
      rc = run_conv(c2, t, B, oh, ow, identity, 1, y, s);
      if (rc) return rc;
      // rotate buffers: y becomes current, old current becomes free
      int cur_idx = 0;
      for (int i = 0; i < 4; ++i) if (buf[i] == cur) cur_idx = i;
      const int y_idx = free_idx[1];
      free_idx[1] = cur_idx;
      cur = buf[y_idx];
      ch = oh; cw = ow; C = c1.cout;
      rc = tap_copy_f32(cur, (size_t)B * ch * cw * C, &taps, s);
      if (rc) return rc;
    }
  }
  avgpool_kernel<float><<<B, 256, 0, s>>>(cur, ch * cw, C, feats);
  return launch_status("avgpool_kernel");
}
