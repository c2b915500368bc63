// SMPL forward for sm_100a: shape blend -> pose-corrective blend -> fused FK + linear blend skinning
// (+ 90-joint epilogue), and the per-vertex sample statistics.
//
// Replaces reference models/smpl_official.py:27-41 and the smplx 0.1.26 lbs() it calls
// (SURVEY.md §8c steps 1-9). Data layout in HBM (all fp32):
//   v_shaped  [Mb][20672]   one row per distinct shape (pitch padded to 16 B)
//   v_posed   [M][20672]    written by the blend stage (TMA stores: 16-byte row pitch), read once by the LBS kernel
//   vertices  [M][6890][3]  the reference's output layout; joints [M][90][3]
// Model constants are repacked once at create time:
//   shapedirs -> [10][20672]; posedirs -> [207][20672] (row pitch padded for float4 loads);
//   J_regressor is folded into J_template [24*3] + J_shapedirs [24*3][10] (J is linear in beta);
//   lbs_weights -> per-vertex top-K (idx,w) in SoA [K][6890] (K = max nnz, 4 for SMPL);
//   the three extra joint regressors -> one CSR (255 nnz).
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <math.h>

using namespace hp3d;

struct SmplTree {
  int8_t parent[NJ];
  int8_t depth[NJ];
  int max_depth;
};

struct hp3d_smpl {
  float* v_template = nullptr;   // [VPITCH]
  float* shapedirs_t = nullptr;  // [10][VPITCH]
  float* posedirs = nullptr;     // [207][VPITCH]
  float* J_template = nullptr;   // [72]
  float* J_shapedirs = nullptr;  // [72][10]
  int skin_k = 0;
  uint8_t* skin_idx = nullptr;   // [skin_k][NV]
  float* skin_w = nullptr;       // [skin_k][NV]
  int* reg_rowptr = nullptr;     // [NREG+1]
  int* reg_col = nullptr;
  float* reg_val = nullptr;
  int* pick_ids = nullptr;       // [NPICK]
  SmplTree tree;
  void* blend_tc = nullptr;      // tensor-core pose-blend operands (gemm_tc.cu), optional
  void* fused = nullptr;         // fused blend + skinning + statistics plan (smpl_fused.cu), optional
  // tile-local skinning tables (64-vertex tiles): distinct joints of the tile + dense per-vertex weights
  int tile_nq_max = 0;           // 0 => tables unavailable (some tile touches > 12 joints): generic kernel
  int* tile_nq = nullptr;        // [NT]
  int* tile_joff = nullptr;      // [NT][12]  joint * 3 (float4 units inside one mesh's A block)
  float* tile_w = nullptr;       // [NT][12][64]
  int* tile_ustart = nullptr;    // [NT+1] per-tile range into tile_uent
  int* tile_uent = nullptr;      // (vertex_in_tile | slot << 8): vertices the 66 extra joints depend on
  int* reg_slot = nullptr;       // CSR column -> parked-vertex slot
  int* pick_slot = nullptr;      // [NPICK]
};

// ------------------------------------------------------------------ shape blend (K=10, fp32 FFMA)
// v_shaped[mb][c] = v_template[c] + sum_l beta[mb][l] * shapedirs[c][l];  J[mb] = J_t + J_s beta.
__global__ void __launch_bounds__(256) shape_blend_kernel(const float* __restrict__ betas, int Mb,
                                                          const float* __restrict__ v_template,
                                                          const float* __restrict__ shapedirs_t,
                                                          const float* __restrict__ J_template,
                                                          const float* __restrict__ J_shapedirs,
                                                          float* __restrict__ v_shaped, float* __restrict__ J) {
  const int mb = blockIdx.y;
  __shared__ float sb[NBETA];
  if (threadIdx.x < NBETA) sb[threadIdx.x] = betas[mb * NBETA + threadIdx.x];
  __syncthreads();
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;   // float4 index within the padded row
  if (c4 < VPITCH / 4) {
    float4 acc = reinterpret_cast<const float4*>(v_template)[c4];
#pragma unroll
    for (int l = 0; l < NBETA; ++l) {
      const float4 s = reinterpret_cast<const float4*>(shapedirs_t + (size_t)l * VPITCH)[c4];
      const float b = sb[l];
      acc.x = fmaf(b, s.x, acc.x); acc.y = fmaf(b, s.y, acc.y); acc.z = fmaf(b, s.z, acc.z); acc.w = fmaf(b, s.w, acc.w);
    }
    reinterpret_cast<float4*>(v_shaped + (size_t)mb * VPITCH)[c4] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x < NJ * 3) {
    float a = J_template[threadIdx.x];
#pragma unroll
    for (int l = 0; l < NBETA; ++l) a = fmaf(sb[l], J_shapedirs[threadIdx.x * NBETA + l], a);
    J[(size_t)mb * NJ * 3 + threadIdx.x] = a;
  }
}

// ------------------------------------------------------------------ pose blend, fp32 CUDA-core GEMM
// v_posed[m][c] = v_shaped[m/rep][c] + sum_k (R[m][k] - I[k]) * posedirs[k][c],  K = 207.
// Exact-fp32 variant (parity path). The tensor-core variant lives in gemm_tc.cu.
constexpr int PB_BM = 128, PB_BN = 128, PB_BK = 8;
__global__ void __launch_bounds__(256) pose_blend_fp32_kernel(const float* __restrict__ body_pose,
                                                              const float* __restrict__ posedirs,
                                                              const float* __restrict__ v_shaped, int M, int rep,
                                                              float* __restrict__ v_posed) {
  __shared__ float As[PB_BK][PB_BM + 4];
  __shared__ __align__(16) float Bs[PB_BK][PB_BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * PB_BM, n0 = blockIdx.x * PB_BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < NPF; k0 += PB_BK) {
    // A tile: 128 meshes x 8 features, pose feature computed on the fly
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (tid >> 3) + 32 * i, k = k0 + (tid & 7), m = m0 + r;
      float v = 0.f;
      if (m < M && k < NPF) {
        const int e = k % 9;
        v = body_pose[(size_t)m * NPF + k] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
      }
      As[tid & 7][r] = v;
    }
    {
      const int r = tid >> 5, c4 = tid & 31, k = k0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < NPF) v = *reinterpret_cast<const float4*>(posedirs + (size_t)k * VPITCH + n0 + c4 * 4);
      *reinterpret_cast<float4*>(&Bs[r][c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PB_BK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; a[4 + i] = As[kk][64 + ty * 4 + i]; }
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const float* vs = v_shaped + (size_t)(m / rep) * VPITCH;
    float* out = v_posed + (size_t)m * VPITCH;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int nn = n + 2 * p;
        if (nn < NV3) {
          float2 o;
          o.x = vs[nn] + acc[i][h * 4 + 2 * p];
          o.y = vs[nn + 1] + acc[i][h * 4 + 2 * p + 1];
          *reinterpret_cast<float2*>(out + nn) = o;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ fused FK + LBS + joints
// One CTA per mesh (grid-stride). Warp 0 walks the 24-joint tree level by level in shared memory
// (G_j = G_parent * [R_j | J_j - J_parent]), forms A_j = [R^G_j | t^G_j - R^G_j J_j], then all
// threads skin the 6890 vertices: v' = (sum_k w_k A_{idx_k}) [v;1]. Epilogue: 24 posed joints,
// 21 picked vertices, 45 sparse-regressed joints (read back through L2).
__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* c) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c[i * 3 + j] = fmaf(a[i * 3 + 2], b[6 + j], fmaf(a[i * 3 + 1], b[3 + j], a[i * 3] * b[j]));
}

__global__ void __launch_bounds__(256) lbs_kernel(const float* __restrict__ v_posed, const float* __restrict__ J,
                                                  int Mb, const float* __restrict__ global_orient, int Mg,
                                                  const float* __restrict__ body_pose, int M,
                                                  const uint8_t* __restrict__ skin_idx,
                                                  const float* __restrict__ skin_w, int skin_k,
                                                  const int* __restrict__ reg_rowptr, const int* __restrict__ reg_col,
                                                  const float* __restrict__ reg_val, const int* __restrict__ pick_ids,
                                                  SmplTree tree, float* __restrict__ vertices,
                                                  float* __restrict__ joints) {
  __shared__ float4 sA[NJ][3];
  __shared__ float sG[NJ][12];
  const int tid = threadIdx.x;
  const int repb = M / Mb, repg = M / Mg;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    if (tid < 32) {
      const int j = tid;
      float R[9], Jj[3] = {0.f, 0.f, 0.f}, rel[3] = {0.f, 0.f, 0.f};
      int par = -1, dep = 99;
      if (j < NJ) {
        const float* src = (j == 0) ? (global_orient + (size_t)(m / repg) * 9)
                                    : (body_pose + ((size_t)m * NBJ + (j - 1)) * 9);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = src[e];
        const float* Jm = J + (size_t)(m / repb) * NJ * 3;
        par = tree.parent[j]; dep = tree.depth[j];
#pragma unroll
        for (int e = 0; e < 3; ++e) { Jj[e] = Jm[j * 3 + e]; rel[e] = (par >= 0) ? (Jj[e] - Jm[par * 3 + e]) : Jj[e]; }
      }
      float G[12];
      for (int d = 0; d <= tree.max_depth; ++d) {
        if (dep == d) {
          if (par < 0) {
#pragma unroll
            for (int e = 0; e < 9; ++e) G[e] = R[e];
            G[9] = rel[0]; G[10] = rel[1]; G[11] = rel[2];
          } else {
            float P[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) P[e] = sG[par][e];
            mat3_mul(P, R, G);
#pragma unroll
            for (int i = 0; i < 3; ++i)
              G[9 + i] = fmaf(P[i * 3 + 2], rel[2], fmaf(P[i * 3 + 1], rel[1], P[i * 3] * rel[0])) + P[9 + i];
          }
#pragma unroll
          for (int e = 0; e < 12; ++e) sG[j][e] = G[e];
        }
        __syncwarp();
      }
      if (j < NJ) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float t = G[9 + i] - fmaf(G[i * 3 + 2], Jj[2], fmaf(G[i * 3 + 1], Jj[1], G[i * 3] * Jj[0]));
          sA[j][i] = make_float4(G[i * 3], G[i * 3 + 1], G[i * 3 + 2], t);
        }
        if (joints) {
          float* jo = joints + ((size_t)m * NOUTJ + j) * 3;
          jo[0] = G[9]; jo[1] = G[10]; jo[2] = G[11];
        }
      }
    }
    __syncthreads();
    const float* vp = v_posed + (size_t)m * VPITCH;
    float* vo = vertices + (size_t)m * NV3;
    for (int v = tid; v < NV; v += blockDim.x) {
      const float x = vp[3 * v], y = vp[3 * v + 1], z = vp[3 * v + 2];
      float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0, t2 = t0;
      for (int k = 0; k < skin_k; ++k) {
        const int j = skin_idx[k * NV + v];
        const float w = skin_w[k * NV + v];
        const float4 a0 = sA[j][0], a1 = sA[j][1], a2 = sA[j][2];
        t0.x = fmaf(w, a0.x, t0.x); t0.y = fmaf(w, a0.y, t0.y); t0.z = fmaf(w, a0.z, t0.z); t0.w = fmaf(w, a0.w, t0.w);
        t1.x = fmaf(w, a1.x, t1.x); t1.y = fmaf(w, a1.y, t1.y); t1.z = fmaf(w, a1.z, t1.z); t1.w = fmaf(w, a1.w, t1.w);
        t2.x = fmaf(w, a2.x, t2.x); t2.y = fmaf(w, a2.y, t2.y); t2.z = fmaf(w, a2.z, t2.z); t2.w = fmaf(w, a2.w, t2.w);
      }
      vo[3 * v]     = fmaf(t0.z, z, fmaf(t0.y, y, t0.x * x)) + t0.w;
      vo[3 * v + 1] = fmaf(t1.z, z, fmaf(t1.y, y, t1.x * x)) + t1.w;
      vo[3 * v + 2] = fmaf(t2.z, z, fmaf(t2.y, y, t2.x * x)) + t2.w;
    }
    __syncthreads();   // vertices of this mesh visible to the whole CTA; sA reusable
    if (joints && tid < NPICK + NREG) {
      float ax = 0.f, ay = 0.f, az = 0.f;
      if (tid < NPICK) {
        const int v = pick_ids[tid];
        ax = __ldcg(vo + 3 * v); ay = __ldcg(vo + 3 * v + 1); az = __ldcg(vo + 3 * v + 2);
      } else {
        const int r = tid - NPICK;
        for (int p = reg_rowptr[r]; p < reg_rowptr[r + 1]; ++p) {
          const int v = reg_col[p];
          const float w = reg_val[p];
          ax = fmaf(w, __ldcg(vo + 3 * v), ax); ay = fmaf(w, __ldcg(vo + 3 * v + 1), ay); az = fmaf(w, __ldcg(vo + 3 * v + 2), az);
        }
      }
      float* jo = joints + ((size_t)m * NOUTJ + NJ + tid) * 3;
      jo[0] = ax; jo[1] = ay; jo[2] = az;
    }
  }
}

// ------------------------------------------------------------------ fused FK + LBS + joints, tile-local variant
// The generic kernel above is bound by shared-memory wavefronts: every vertex gathers 4 x 48 B of A with
// lane-varying addresses (12 LDS.128 = 48 wavefronts per 32 vertices, vs ~33 cycles of HBM time).
// Skinning weights are spatially coherent (a run of consecutive vertices touches a handful of joints), so at
// create time each 64-vertex tile gets the list of joints it touches (nq <= 12) and dense per-vertex weights
// over that list. Then A_{joint q} is a WARP-UNIFORM shared-memory address (broadcast, one wavefront per
// LDS.128), weights/indices live in registers across the G meshes a CTA owns, and the vertex stream moves with
// 8-byte vector loads/stores (2 vertices = 24 B per lane; the 82,680-byte mesh rows are 8-byte aligned for
// every mesh), the loads of the next two meshes being issued before the current two are skinned.
// One CTA = G = 8 consecutive meshes: warp g runs the FK of mesh g; the 8 warps sweep 108 tiles x G meshes;
// the <= 276 vertices the 66 extra joints depend on are parked in shared memory as they are produced, so the
// joint epilogue never re-reads vertices from HBM (they would already have been evicted from L2).
// History (profiles/README.md): generic 2.02 ms (32 % of HBM peak) -> tile-local 1.09 ms (60 %) -> this.
constexpr int TV = 64;                      // vertices per tile
constexpr int NT = (NV + TV - 1) / TV;      // 108
constexpr int NQCAP = 12;
constexpr int LBS_GMAX = 8;                 // meshes per CTA (<= warps per CTA); small batches use 2 so every SM gets a CTA
constexpr int NU_MAX = NPICK + 255;         // unique vertices feeding the 66 extra joints (<= 276)
constexpr int LBS_DEFAULT_MODE = 1;         // FFMA2 blending, 2 meshes ahead (see hp3d_smpl_lbs; profiles/r01i_lbs_sweep.jsonl)

struct LbsTileCtx {
  int tile, lane, m0, Gv, park0, park1;
  const float* v_posed; float* vertices;
  const int* tile_joff; const float* tile_w;
  const float4* sA;     // [G][NJ*3]
  float* sV;            // [G][NU_MAX][3]
};

// Packed fp32 FMA (Blackwell FFMA2): acc.xy += s * v.xy in ONE issue slot; ptxas encodes the scalar as a
// broadcast operand (FFMA2 R, R.F32, R.F32x2, R.F32x2), so (s,s) costs no extra register.
__device__ __forceinline__ void ffma2(float2& acc, float s, float x, float y) {
  unsigned long long a, b, c, d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(x), "f"(y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(d));
}

// One 64-vertex tile x Gv meshes with exactly NQ joints (compile-time): no per-joint predicates or branches.
// F2: blend the joint transforms with packed FFMA2 (halves the FMA issue slots, the kernel's dominant
// instruction); PF: meshes whose vertex loads are in flight ahead of the one being skinned (bytes in flight
// per warp = PF x 768 B).
template <int NQ, bool F2, int PF>
__device__ __forceinline__ void lbs_tile_body(const LbsTileCtx& c) {
  const int lane = c.lane, tile = c.tile;
  const int v0 = tile * TV + 2 * lane;
  const bool valid = v0 < NV;                         // NV is even: a lane's two vertices are both in or out
  float w0[NQ], w1[NQ];
  int joff[NQ];                                       // float4 offset of joint q inside one mesh's A block
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float2 w = *reinterpret_cast<const float2*>(c.tile_w + ((size_t)tile * NQCAP + q) * TV + 2 * lane);
    w0[q] = w.x; w1[q] = w.y;
    joff[q] = c.tile_joff[tile * NQCAP + q];
  }
  const size_t voff = (size_t)3 * v0;
  const float* src0 = c.v_posed + (size_t)c.m0 * VPITCH + voff;
  float* dst0 = c.vertices + (size_t)c.m0 * NV3 + voff;
  const int Gv = c.Gv;
  auto ld = [&](int g, float2 (&p)[3]) {
    p[0] = p[1] = p[2] = make_float2(0.f, 0.f);
    if (valid && g < Gv) {
      const float* s_ = src0 + (size_t)g * VPITCH;
      p[0] = *reinterpret_cast<const float2*>(s_);
      p[1] = *reinterpret_cast<const float2*>(s_ + 2);
      p[2] = *reinterpret_cast<const float2*>(s_ + 4);
    }
  };
  auto emit = [&](int g, const float2 (&p)[3], const float4& a0, const float4& a1, const float4& a2,
                  const float4& b0, const float4& b1, const float4& b2) {
    if (valid) {
      const float x0 = p[0].x, y0 = p[0].y, z0 = p[1].x, x1 = p[1].y, y1 = p[2].x, z1 = p[2].y;
      float2 o0, o1, o2;
      o0.x = fmaf(a0.z, z0, fmaf(a0.y, y0, a0.x * x0)) + a0.w;
      o0.y = fmaf(a1.z, z0, fmaf(a1.y, y0, a1.x * x0)) + a1.w;
      o1.x = fmaf(a2.z, z0, fmaf(a2.y, y0, a2.x * x0)) + a2.w;
      o1.y = fmaf(b0.z, z1, fmaf(b0.y, y1, b0.x * x1)) + b0.w;
      o2.x = fmaf(b1.z, z1, fmaf(b1.y, y1, b1.x * x1)) + b1.w;
      o2.y = fmaf(b2.z, z1, fmaf(b2.y, y1, b2.x * x1)) + b2.w;
      float* d_ = dst0 + (size_t)g * NV3;
      *reinterpret_cast<float2*>(d_) = o0;
      *reinterpret_cast<float2*>(d_ + 2) = o1;
      *reinterpret_cast<float2*>(d_ + 4) = o2;
      if (c.park0 >= 0) { float* sv = c.sV + ((size_t)g * NU_MAX + c.park0) * 3; sv[0] = o0.x; sv[1] = o0.y; sv[2] = o1.x; }
      if (c.park1 >= 0) { float* sv = c.sV + ((size_t)g * NU_MAX + c.park1) * 3; sv[0] = o1.y; sv[1] = o2.x; sv[2] = o2.y; }
    }
  };
  auto skin = [&](int g, const float2 (&p)[3]) {
    const float4* Ag = c.sA + g * (NJ * 3);
    if constexpr (F2) {
      float2 a[6], b[6];                               // rows 0..2 as (xy, zw) pairs, vertex 0 / vertex 1
#pragma unroll
      for (int e = 0; e < 6; ++e) a[e] = b[e] = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 r0 = Ag[joff[q]], r1 = Ag[joff[q] + 1], r2 = Ag[joff[q] + 2];
        const float u = w0[q], v = w1[q];
        ffma2(a[0], u, r0.x, r0.y); ffma2(a[1], u, r0.z, r0.w);
        ffma2(a[2], u, r1.x, r1.y); ffma2(a[3], u, r1.z, r1.w);
        ffma2(a[4], u, r2.x, r2.y); ffma2(a[5], u, r2.z, r2.w);
        ffma2(b[0], v, r0.x, r0.y); ffma2(b[1], v, r0.z, r0.w);
        ffma2(b[2], v, r1.x, r1.y); ffma2(b[3], v, r1.z, r1.w);
        ffma2(b[4], v, r2.x, r2.y); ffma2(b[5], v, r2.z, r2.w);
      }
      emit(g, p, make_float4(a[0].x, a[0].y, a[1].x, a[1].y), make_float4(a[2].x, a[2].y, a[3].x, a[3].y),
           make_float4(a[4].x, a[4].y, a[5].x, a[5].y), make_float4(b[0].x, b[0].y, b[1].x, b[1].y),
           make_float4(b[2].x, b[2].y, b[3].x, b[3].y), make_float4(b[4].x, b[4].y, b[5].x, b[5].y));
    } else {
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, b0 = a0, b1 = a0, b2 = a0;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 r0 = Ag[joff[q]], r1 = Ag[joff[q] + 1], r2 = Ag[joff[q] + 2];
        const float u = w0[q], v = w1[q];
        a0.x = fmaf(u, r0.x, a0.x); a0.y = fmaf(u, r0.y, a0.y); a0.z = fmaf(u, r0.z, a0.z); a0.w = fmaf(u, r0.w, a0.w);
        a1.x = fmaf(u, r1.x, a1.x); a1.y = fmaf(u, r1.y, a1.y); a1.z = fmaf(u, r1.z, a1.z); a1.w = fmaf(u, r1.w, a1.w);
        a2.x = fmaf(u, r2.x, a2.x); a2.y = fmaf(u, r2.y, a2.y); a2.z = fmaf(u, r2.z, a2.z); a2.w = fmaf(u, r2.w, a2.w);
        b0.x = fmaf(v, r0.x, b0.x); b0.y = fmaf(v, r0.y, b0.y); b0.z = fmaf(v, r0.z, b0.z); b0.w = fmaf(v, r0.w, b0.w);
        b1.x = fmaf(v, r1.x, b1.x); b1.y = fmaf(v, r1.y, b1.y); b1.z = fmaf(v, r1.z, b1.z); b1.w = fmaf(v, r1.w, b1.w);
        b2.x = fmaf(v, r2.x, b2.x); b2.y = fmaf(v, r2.y, b2.y); b2.z = fmaf(v, r2.z, b2.z); b2.w = fmaf(v, r2.w, b2.w);
      }
      emit(g, p, a0, a1, a2, b0, b1, b2);
    }
  };
  float2 cur[PF][3], nxt[PF][3];
#pragma unroll
  for (int i = 0; i < PF; ++i) ld(i, cur[i]);
  for (int g = 0; g < Gv; g += PF) {
#pragma unroll
    for (int i = 0; i < PF; ++i) ld(g + PF + i, nxt[i]);
#pragma unroll
    for (int i = 0; i < PF; ++i)
      if (g + i < Gv) skin(g + i, cur[i]);
#pragma unroll
    for (int i = 0; i < PF; ++i)
#pragma unroll
      for (int e = 0; e < 3; ++e) cur[i][e] = nxt[i][e];
  }
}

template <int NQMAX, bool F2, int PF, int LBS_G>
__global__ void __launch_bounds__(256, 2) lbs_tile_kernel(const float* __restrict__ v_posed, const float* __restrict__ J,
                                                          int Mb, const float* __restrict__ global_orient, int Mg,
                                                          const float* __restrict__ body_pose, int M,
                                                          const int* __restrict__ tile_nq, const int* __restrict__ tile_joff,
                                                          const float* __restrict__ tile_w,
                                                          const int* __restrict__ tile_ustart, const int* __restrict__ tile_uent,
                                                          const int* __restrict__ reg_rowptr, const int* __restrict__ reg_slot,
                                                          const float* __restrict__ reg_val, const int* __restrict__ pick_slot,
                                                          SmplTree tree, float* __restrict__ vertices,
                                                          float* __restrict__ joints) {
  __shared__ float4 sA[LBS_G][NJ * 3];
  __shared__ float sV[LBS_G][NU_MAX][3];                               // parked vertices for the joint epilogue
  float (*sG)[NJ][12] = reinterpret_cast<float (*)[NJ][12]>(&sV[0][0][0]);   // FK scratch (phase 1) aliases sV
  __shared__ int next_tile;
  if (threadIdx.x == 0) next_tile = 0;
  static_assert(sizeof(float) * LBS_G * NJ * 12 <= sizeof(float) * LBS_G * NU_MAX * 3, "FK scratch must fit");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int repb = M / Mb, repg = M / Mg;
  const int m0 = blockIdx.x * LBS_G;
  const int Gv = min(LBS_G, M - m0);
  // ---- phase 1: forward kinematics, warp g <-> mesh m0 + g
  if (warp < Gv) {   // LBS_G <= 8 warps
    const int m = m0 + warp, j = lane;
    float R[9], Jj[3] = {0.f, 0.f, 0.f}, rel[3] = {0.f, 0.f, 0.f};
    int par = -1, dep = 99;
    if (j < NJ) {
      const float* src = (j == 0) ? (global_orient + (size_t)(m / repg) * 9) : (body_pose + ((size_t)m * NBJ + (j - 1)) * 9);
#pragma unroll
      for (int e = 0; e < 9; ++e) R[e] = src[e];
      const float* Jm = J + (size_t)(m / repb) * NJ * 3;
      par = tree.parent[j]; dep = tree.depth[j];
#pragma unroll
      for (int e = 0; e < 3; ++e) { Jj[e] = Jm[j * 3 + e]; rel[e] = (par >= 0) ? (Jj[e] - Jm[par * 3 + e]) : Jj[e]; }
    }
    float G[12];
    for (int d = 0; d <= tree.max_depth; ++d) {
      if (dep == d) {
        if (par < 0) {
#pragma unroll
          for (int e = 0; e < 9; ++e) G[e] = R[e];
          G[9] = rel[0]; G[10] = rel[1]; G[11] = rel[2];
        } else {
          float P[12];
#pragma unroll
          for (int e = 0; e < 12; ++e) P[e] = sG[warp][par][e];
          mat3_mul(P, R, G);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            G[9 + i] = fmaf(P[i * 3 + 2], rel[2], fmaf(P[i * 3 + 1], rel[1], P[i * 3] * rel[0])) + P[9 + i];
        }
#pragma unroll
        for (int e = 0; e < 12; ++e) sG[warp][j][e] = G[e];
      }
      __syncwarp();
    }
    if (j < NJ) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float t = G[9 + i] - fmaf(G[i * 3 + 2], Jj[2], fmaf(G[i * 3 + 1], Jj[1], G[i * 3] * Jj[0]));
        sA[warp][j * 3 + i] = make_float4(G[i * 3], G[i * 3 + 1], G[i * 3 + 2], t);
      }
      if (joints) {
        float* jo = joints + ((size_t)m * NOUTJ + j) * 3;
        jo[0] = G[9]; jo[1] = G[10]; jo[2] = G[11];
      }
    }
  }
  __syncthreads();
  // ---- phase 2: skinning, warp sweeps its tiles, meshes innermost. The tile's joint count is warp-uniform:
  // dispatch once per tile to a fully unrolled, branch-free body for exactly that count.
  // tiles cost between 1 and 8(12) joints: warps take the next unprocessed tile from a CTA-wide counter
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(&next_tile, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= NT) break;
    const int nq = tile_nq[tile];
    LbsTileCtx c;
    c.tile = tile; c.lane = lane; c.m0 = m0; c.Gv = Gv; c.v_posed = v_posed; c.vertices = vertices;
    c.tile_joff = tile_joff; c.tile_w = tile_w; c.sA = &sA[0][0]; c.sV = &sV[0][0][0];
    c.park0 = -1; c.park1 = -1;
    if (joints) {      // which of my two vertices feed the joint epilogue, and into which shared-memory slot
      for (int e = tile_ustart[tile]; e < tile_ustart[tile + 1]; ++e) {
        const int ent = tile_uent[e], vl = ent & 0xFF, slot = ent >> 8;
        if ((vl >> 1) == lane) { if (vl & 1) c.park1 = slot; else c.park0 = slot; }
      }
    }
    switch (nq) {
      case 1: lbs_tile_body<1, F2, PF>(c); break;
      case 2: lbs_tile_body<2, F2, PF>(c); break;
      case 3: lbs_tile_body<3, F2, PF>(c); break;
      case 4: lbs_tile_body<4, F2, PF>(c); break;
      case 5: lbs_tile_body<5, F2, PF>(c); break;
      case 6: lbs_tile_body<6, F2, PF>(c); break;
      case 7: lbs_tile_body<7, F2, PF>(c); break;
      case 8: lbs_tile_body<8, F2, PF>(c); break;
      default:
        if constexpr (NQMAX > 8) {
          switch (nq) {
            case 9: lbs_tile_body<9, F2, PF>(c); break;
            case 10: lbs_tile_body<10, F2, PF>(c); break;
            case 11: lbs_tile_body<11, F2, PF>(c); break;
            case 12: lbs_tile_body<12, F2, PF>(c); break;
            default: break;
          }
        }
        break;
    }
  }
  if (!joints) return;
  __syncthreads();
  // ---- phase 3: picked + regressed joints of the G meshes from the parked vertices
  for (int it = tid; it < Gv * (NPICK + NREG); it += 256) {
    const int g = it / (NPICK + NREG), r = it - g * (NPICK + NREG);
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (r < NPICK) {
      const int sl = pick_slot[r];
      ax = sV[g][sl][0]; ay = sV[g][sl][1]; az = sV[g][sl][2];
    } else {
      const int rr = r - NPICK;
      for (int p = reg_rowptr[rr]; p < reg_rowptr[rr + 1]; ++p) {
        const int sl = reg_slot[p];
        const float w = reg_val[p];
        ax = fmaf(w, sV[g][sl][0], ax); ay = fmaf(w, sV[g][sl][1], ay); az = fmaf(w, sV[g][sl][2], az);
      }
    }
    float* jo = joints + ((size_t)(m0 + g) * NOUTJ + NJ + r) * 3;
    jo[0] = ax; jo[1] = ay; jo[2] = az;
  }
}

// ------------------------------------------------------------------ small rotation kernels
__global__ void rodrigues_kernel(const float* __restrict__ aa, int n, float* __restrict__ R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // smplx batch_rodrigues: angle = ||r + 1e-8||, dir = r / angle, R = I + sin K + (1 - cos) K^2
  const float rx = aa[3 * i], ry = aa[3 * i + 1], rz = aa[3 * i + 2];
  const float ex = rx + 1e-8f, ey = ry + 1e-8f, ez = rz + 1e-8f;
  const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
  const float dx = rx / angle, dy = ry / angle, dz = rz / angle;
  float s, c;
  sincosf(angle, &s, &c);
  const float oc = 1.f - c;
  // K = [[0,-dz,dy],[dz,0,-dx],[-dy,dx,0]];  K^2 = d d^T - |d|^2 I
  const float dd = dx * dx + dy * dy + dz * dz;
  float* o = R + 9 * (size_t)i;
  o[0] = 1.f + oc * (dx * dx - dd); o[1] = -s * dz + oc * dx * dy;    o[2] = s * dy + oc * dx * dz;
  o[3] = s * dz + oc * dx * dy;     o[4] = 1.f + oc * (dy * dy - dd); o[5] = -s * dx + oc * dy * dz;
  o[6] = -s * dy + oc * dx * dz;    o[7] = s * dx + oc * dy * dz;     o[8] = 1.f + oc * (dz * dz - dd);
}

__global__ void rot6d_kernel(const float* __restrict__ x, int n, float* __restrict__ R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // reference utils/rigid_transform_utils.py:88-94: x.view(-1,3,2): a1 = elements 0,2,4; a2 = 1,3,5;
  // F.normalize uses max(||a||, 1e-12); columns stacked (b1,b2,b3).
  const float* p = x + 6 * (size_t)i;
  float a1[3] = {p[0], p[2], p[4]}, a2[3] = {p[1], p[3], p[5]};
  float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
  float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
  const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
  float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
  float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
  float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
  float* o = R + 9 * (size_t)i;
#pragma unroll
  for (int r = 0; r < 3; ++r) { o[r * 3] = b1[r]; o[r * 3 + 1] = b2[r]; o[r * 3 + 2] = b3[r]; }
}

// ------------------------------------------------------------------ per-vertex sample statistics
// reference utils/sampling_utils.py:189-190, batched over images: thread = (image, vertex).
__global__ void __launch_bounds__(256) vertex_uncertainty_kernel(const float* __restrict__ verts, int B, int N,
                                                                 float* __restrict__ mean_out,
                                                                 float* __restrict__ dist_out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= NV) return;
  const float* base = verts + ((size_t)b * N) * NV3 + 3 * v;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* p = base + (size_t)n * NV3;
    sx += p[0]; sy += p[1]; sz += p[2];
  }
  const float inv = 1.f / (float)N;
  const float mx = sx * inv, my = sy * inv, mz = sz * inv;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* p = base + (size_t)n * NV3;
    const float dx = p[0] - mx, dy = p[1] - my, dz = p[2] - mz;
    acc += sqrtf(dx * dx + dy * dy + dz * dz);
  }
  dist_out[(size_t)b * NV + v] = acc * inv;
  if (mean_out) {
    float* mo = mean_out + ((size_t)b * NV + v) * 3;
    mo[0] = mx; mo[1] = my; mo[2] = mz;
  }
}

// Single-read variant: a CTA stages ALL N samples of 64 consecutive vertices (N x 768 B, contiguous per sample)
// in shared memory with one coalesced pass over HBM, then computes the mean and the mean distance from it out of
// shared memory. Halves the HBM traffic of the two-pass kernel above (which remains the fallback for N > 280).
constexpr int UNC_DEFAULT_VARIANT = 3;     // 64-vertex tiles staged by cp.async: 0.50 ms vs 0.52 (variant 1) / 0.63 (variant 2)
constexpr int UNC_TV = 32;                 // vertices per CTA: 96 floats (384 B) per sample, N x 384 B of shared memory
constexpr int UNC_F = UNC_TV * 3;
__global__ void __launch_bounds__(256) vertex_uncertainty_smem_kernel(const float* __restrict__ verts, int B, int N,
                                                                      float* __restrict__ mean_out,
                                                                      float* __restrict__ dist_out) {
  extern __shared__ float us[];                       // [N][96] samples | [96] mean | [8][32] partial distances
  const int b = blockIdx.y, v0 = blockIdx.x * UNC_TV;
  const int nfl = min(UNC_TV, NV - v0) * 3;           // floats per sample in this tile (96, last tile 30)
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  float* mean = us + (size_t)N * UNC_F;
  float* part = mean + UNC_F;
  const float* base = verts + (size_t)b * N * NV3 + (size_t)v0 * 3;
  // warp w stages samples w, w+8, ...; eight samples (24 independent row loads per lane) in flight at a time
  for (int n0 = w; n0 < N; n0 += 64) {
    float r[8][3];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = n0 + 8 * u;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int f = lane + 32 * i;
        r[u][i] = (n < N && f < nfl) ? base[(size_t)n * NV3 + f] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = n0 + 8 * u;
      if (n < N) {
#pragma unroll
        for (int i = 0; i < 3; ++i) us[n * UNC_F + lane + 32 * i] = r[u][i];
      }
    }
  }
  __syncthreads();
  if (t < UNC_F) {
    float s0 = 0.f, s1 = 0.f;
    int n = 0;
    for (; n + 1 < N; n += 2) { s0 += us[n * UNC_F + t]; s1 += us[(n + 1) * UNC_F + t]; }
    if (n < N) s0 += us[n * UNC_F + t];
    mean[t] = (s0 + s1) / (float)N;
  }
  __syncthreads();
  {
    const int v = lane, q = w;                        // 8 sample groups x 32 vertices
    const float mx = mean[3 * v], my = mean[3 * v + 1], mz = mean[3 * v + 2];
    float acc = 0.f;
    for (int n = q; n < N; n += 8) {
      const float* p = us + n * UNC_F + 3 * v;
      const float dx = p[0] - mx, dy = p[1] - my, dz = p[2] - mz;
      acc += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    part[q * 32 + v] = acc;
  }
  __syncthreads();
  if (t < 32 && v0 + t < NV) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) a += part[q * 32 + t];
    dist_out[(size_t)b * NV + v0 + t] = a / (float)N;
    if (mean_out) {
      float* mo = mean_out + ((size_t)b * NV + v0 + t) * 3;
      mo[0] = mean[3 * t]; mo[1] = mean[3 * t + 1]; mo[2] = mean[3 * t + 2];
    }
  }
}

// 64-vertex variant with 8-byte loads: a sample row of the tile is 192 floats = 96 float2 (rows start 8-byte aligned:
// 82,680 = 8 * 10,335 and 64 vertices = 768 B), three coalesced 256-byte loads per warp instead of six 128-byte ones,
// half the load / shared-store instructions per byte. N x 768 B of shared memory (2 CTAs/SM at N = 100).
constexpr int UNC2_TV = 64;
constexpr int UNC2_F = UNC2_TV * 3;        // 192
// ASYNC: stage with 8-byte cp.async (LDGSTS) instead of register round trips: no registers or STS per byte and every row
// of the tile is in flight at once (ptxas interleaves the register variant's loads and stores in groups of ~8).
template <bool ASYNC>
__global__ void __launch_bounds__(256) vertex_uncertainty_smem2_kernel(const float* __restrict__ verts, int B, int N,
                                                                       float* __restrict__ mean_out,
                                                                       float* __restrict__ dist_out) {
  extern __shared__ __align__(8) float us2[];          // [N][192] samples | [192] mean | [4][64] partial distances
  const int b = blockIdx.y, v0 = blockIdx.x * UNC2_TV;
  const int nv = min(UNC2_TV, NV - v0);                // 64, last tile 42 (even: whole float2s)
  const int nf2 = nv * 3 / 2;                          // float2s per sample row in this tile
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  float* mean = us2 + (size_t)N * UNC2_F;
  float* part = mean + UNC2_F;
  const float* base = verts + (size_t)b * N * NV3 + (size_t)v0 * 3;
  if (ASYNC) {
    for (int n = w; n < N; n += 8) {
      const float* row = base + (size_t)n * NV3;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int f = lane + 32 * i;
        if (f < nf2) {
          const unsigned dst = (unsigned)__cvta_generic_to_shared(us2 + (size_t)n * UNC2_F + 2 * f);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(row + 2 * f) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else
  for (int n0 = w; n0 < N; n0 += 64) {                 // warp w stages samples w, w+8, ...; eight rows in flight
    float2 r[8][3];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = n0 + 8 * u;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int f = lane + 32 * i;
        r[u][i] = (n < N && f < nf2) ? reinterpret_cast<const float2*>(base + (size_t)n * NV3)[f] : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = n0 + 8 * u;
      if (n < N) {
#pragma unroll
        for (int i = 0; i < 3; ++i) reinterpret_cast<float2*>(us2 + (size_t)n * UNC2_F)[lane + 32 * i] = r[u][i];
      }
    }
  }
  __syncthreads();
  if (t < UNC2_F) {
    float s0 = 0.f, s1 = 0.f;
    int n = 0;
    for (; n + 1 < N; n += 2) { s0 += us2[n * UNC2_F + t]; s1 += us2[(n + 1) * UNC2_F + t]; }
    if (n < N) s0 += us2[n * UNC2_F + t];
    mean[t] = (s0 + s1) / (float)N;
  }
  __syncthreads();
  {
    const int v = t & 63, q = t >> 6;                  // 4 sample groups x 64 vertices
    const float mx = mean[3 * v], my = mean[3 * v + 1], mz = mean[3 * v + 2];
    float acc = 0.f;
    for (int n = q; n < N; n += 4) {
      const float* p = us2 + n * UNC2_F + 3 * v;
      const float dx = p[0] - mx, dy = p[1] - my, dz = p[2] - mz;
      acc += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    part[q * 64 + v] = acc;
  }
  __syncthreads();
  if (t < nv) {
    const float a = (part[t] + part[64 + t]) + (part[128 + t] + part[192 + t]);
    dist_out[(size_t)b * NV + v0 + t] = a / (float)N;
    if (mean_out) {
      float* mo = mean_out + ((size_t)b * NV + v0 + t) * 3;
      mo[0] = mean[3 * t]; mo[1] = mean[3 * t + 1]; mo[2] = mean[3 * t + 2];
    }
  }
}

// ------------------------------------------------------------------ host side
namespace hp3d {
int blend_tc_create(const double* posedirs, const double* shapedirs, const double* v_template, void** out);   // gemm_tc.cu
void blend_tc_destroy(void* p);
size_t blend_tc_workspace_bytes(int M);
int smpl_fused_create(const hp3d_smpl_model* md, const float* Jt, const float* Js, void** out);       // smpl_fused.cu
void smpl_fused_destroy(void* p);
size_t smpl_fused_workspace_bytes(int M);
void smpl_fused_info(const void* p, int* permuted, int* nq_sum, int* nq_max);
int smpl_fused_forward(void* p, const float* betas, int Mb, const float* global_orient, int Mg, const float* body_pose, int M,
                       int samples_per_image, float* vertices, float* joints, float* unc, float* mean, void* workspace,
                       cudaStream_t stream);
int blend_tc_forward(void* p, const float* betas, int Mb, const float* body_pose, int M, float* v_posed,
                     void* workspace, cudaStream_t stream);
}

extern "C" int hp3d_smpl_create(const hp3d_smpl_model* md, hp3d_smpl** out) {
  HP3D_ARG(md && out, "null argument");
  HP3D_ARG(md->v_template && md->shapedirs && md->posedirs && md->J_regressor && md->lbs_weights && md->parents &&
           md->extra_vertex_ids && md->joint_regressors_extra, "null model field");
  hp3d_smpl* h = new hp3d_smpl();
  int rc = 0;
  // tree
  HP3D_ARG(md->parents[0] < 0, "parents[0] must be -1");
  h->tree.max_depth = 0;
  for (int j = 0; j < NJ; ++j) {
    const int p = md->parents[j];
    if (j > 0 && (p < 0 || p >= j)) { delete h; set_error("hp3d_smpl_create: parents must satisfy 0 <= parents[j] < j"); return -1; }
    h->tree.parent[j] = (int8_t)p;
    h->tree.depth[j] = (j == 0) ? 0 : (int8_t)(h->tree.depth[p] + 1);
    h->tree.max_depth = std::max<int>(h->tree.max_depth, h->tree.depth[j]);
  }
  std::vector<float> vt(VPITCH, 0.f), sd((size_t)NBETA * VPITCH, 0.f), pd((size_t)NPF * VPITCH, 0.f);
  for (int c = 0; c < NV3; ++c) vt[c] = (float)md->v_template[c];
  for (int c = 0; c < NV3; ++c)
    for (int l = 0; l < NBETA; ++l) sd[(size_t)l * VPITCH + c] = (float)md->shapedirs[(size_t)c * NBETA + l];
  for (int k = 0; k < NPF; ++k)
    for (int c = 0; c < NV3; ++c) pd[(size_t)k * VPITCH + c] = (float)md->posedirs[(size_t)k * NV3 + c];
  // J = J_regressor (v_template + shapedirs beta) is linear in beta: fold in fp64
  std::vector<float> Jt(NJ * 3), Js((size_t)NJ * 3 * NBETA);
  for (int j = 0; j < NJ; ++j)
    for (int e = 0; e < 3; ++e) {
      double a = 0.0, s[NBETA] = {0};
      for (int v = 0; v < NV; ++v) {
        const double w = md->J_regressor[(size_t)j * NV + v];
        if (w == 0.0) continue;
        a += w * md->v_template[v * 3 + e];
        for (int l = 0; l < NBETA; ++l) s[l] += w * md->shapedirs[((size_t)v * 3 + e) * NBETA + l];
      }
      Jt[j * 3 + e] = (float)a;
      for (int l = 0; l < NBETA; ++l) Js[(size_t)(j * 3 + e) * NBETA + l] = (float)s[l];
    }
  // skinning weights: per-vertex non-zeros, padded to the max count
  int K = 1;
  for (int v = 0; v < NV; ++v) {
    int c = 0;
    for (int j = 0; j < NJ; ++j) c += md->lbs_weights[(size_t)v * NJ + j] != 0.0;
    K = std::max(K, c);
  }
  std::vector<uint8_t> sidx((size_t)K * NV, 0);
  std::vector<float> sw((size_t)K * NV, 0.f);
  for (int v = 0; v < NV; ++v) {
    int c = 0;
    for (int j = 0; j < NJ; ++j) {
      const double w = md->lbs_weights[(size_t)v * NJ + j];
      if (w != 0.0) { sidx[(size_t)c * NV + v] = (uint8_t)j; sw[(size_t)c * NV + v] = (float)w; ++c; }
    }
    for (; c < K; ++c) sidx[(size_t)c * NV + v] = sidx[v];   // zero weight, benign index
  }
  h->skin_k = K;
  // tile-local tables
  std::vector<int> tnq(NT, 0), tjoff((size_t)NT * NQCAP, 0);
  std::vector<float> tw((size_t)NT * NQCAP * TV, 0.f);
  int nq_max = 0;
  bool tiles_ok = true;
  for (int t = 0; t < NT && tiles_ok; ++t) {
    int slot[NJ];
    for (int j = 0; j < NJ; ++j) slot[j] = -1;
    int nq = 0;
    for (int v = t * TV; v < std::min(NV, (t + 1) * TV); ++v)
      for (int j = 0; j < NJ; ++j)
        if (md->lbs_weights[(size_t)v * NJ + j] != 0.0 && slot[j] < 0) slot[j] = nq++;
    if (nq > NQCAP) { tiles_ok = false; break; }
    // renumber in joint order so the table is deterministic
    int q = 0;
    for (int j = 0; j < NJ; ++j) if (slot[j] >= 0) { slot[j] = q; tjoff[(size_t)t * NQCAP + q] = j * 3; ++q; }
    for (int v = t * TV; v < std::min(NV, (t + 1) * TV); ++v)
      for (int j = 0; j < NJ; ++j) {
        const double w = md->lbs_weights[(size_t)v * NJ + j];
        if (w != 0.0) tw[((size_t)t * NQCAP + slot[j]) * TV + (v - t * TV)] = (float)w;
      }
    tnq[t] = nq;
    nq_max = std::max(nq_max, nq);
  }
  h->tile_nq_max = tiles_ok ? nq_max : 0;
  std::vector<int> rp(NREG + 1, 0), rcol;
  std::vector<float> rval;
  for (int r = 0; r < NREG; ++r) {
    for (int v = 0; v < NV; ++v) {
      const double w = md->joint_regressors_extra[(size_t)r * NV + v];
      if (w != 0.0) { rcol.push_back(v); rval.push_back((float)w); }
    }
    rp[r + 1] = (int)rcol.size();
  }
  if (rcol.empty()) { rcol.push_back(0); rval.push_back(0.f); }
  std::vector<int> picks(md->extra_vertex_ids, md->extra_vertex_ids + NPICK);
  for (int p : picks) if (p < 0 || p >= NV) { delete h; set_error("hp3d_smpl_create: extra_vertex_ids out of range"); return -1; }
  // vertices the joint epilogue needs: unique(pick_ids U regressor columns) -> slots, listed per tile
  std::vector<int> slot_of(NV, -1), ustart(NT + 1, 0), uent, rslot(rcol.size(), 0), pslot(NPICK, 0);
  {
    int nu = 0;
    for (int v = 0; v < NV; ++v) {
      bool need = false;
      for (int p_ : picks) need |= (p_ == v);
      if (!need) for (size_t e = 0; e < rcol.size() && !need; ++e) need = (rcol[e] == v && rval[e] != 0.f);
      if (need) slot_of[v] = nu++;
    }
    if (nu > NU_MAX) h->tile_nq_max = 0;     // cannot park that many: generic kernel
    for (int t = 0; t < NT; ++t) {
      ustart[t] = (int)uent.size();
      for (int v = t * TV; v < std::min(NV, (t + 1) * TV); ++v)
        if (slot_of[v] >= 0) uent.push_back((v - t * TV) | (slot_of[v] << 8));
    }
    ustart[NT] = (int)uent.size();
    if (uent.empty()) uent.push_back(0);
    for (size_t e = 0; e < rcol.size(); ++e) rslot[e] = std::max(0, slot_of[rcol[e]]);
    for (int p_ = 0; p_ < NPICK; ++p_) pslot[p_] = std::max(0, slot_of[picks[p_]]);
  }
  rc = rc ? rc : upload(&h->tile_ustart, ustart.data(), ustart.size());
  rc = rc ? rc : upload(&h->tile_uent, uent.data(), uent.size());
  rc = rc ? rc : upload(&h->reg_slot, rslot.data(), rslot.size());
  rc = rc ? rc : upload(&h->pick_slot, pslot.data(), pslot.size());
  rc = rc ? rc : upload(&h->v_template, vt.data(), vt.size());
  rc = rc ? rc : upload(&h->shapedirs_t, sd.data(), sd.size());
  rc = rc ? rc : upload(&h->posedirs, pd.data(), pd.size());
  rc = rc ? rc : upload(&h->J_template, Jt.data(), Jt.size());
  rc = rc ? rc : upload(&h->J_shapedirs, Js.data(), Js.size());
  rc = rc ? rc : upload(&h->skin_idx, sidx.data(), sidx.size());
  rc = rc ? rc : upload(&h->skin_w, sw.data(), sw.size());
  rc = rc ? rc : upload(&h->reg_rowptr, rp.data(), rp.size());
  rc = rc ? rc : upload(&h->reg_col, rcol.data(), rcol.size());
  rc = rc ? rc : upload(&h->reg_val, rval.data(), rval.size());
  rc = rc ? rc : upload(&h->pick_ids, picks.data(), picks.size());
  rc = rc ? rc : upload(&h->tile_nq, tnq.data(), tnq.size());
  rc = rc ? rc : upload(&h->tile_joff, tjoff.data(), tjoff.size());
  rc = rc ? rc : upload(&h->tile_w, tw.data(), tw.size());
  rc = rc ? rc : blend_tc_create(md->posedirs, md->shapedirs, md->v_template, &h->blend_tc);
  rc = rc ? rc : smpl_fused_create(md, Jt.data(), Js.data(), &h->fused);
  if (rc) { hp3d_smpl_destroy(h); return rc; }
  *out = h;
  return 0;
}

extern "C" void hp3d_smpl_destroy(hp3d_smpl* h) {
  if (!h) return;
  cudaFree(h->v_template); cudaFree(h->shapedirs_t); cudaFree(h->posedirs); cudaFree(h->J_template);
  cudaFree(h->J_shapedirs); cudaFree(h->skin_idx); cudaFree(h->skin_w); cudaFree(h->reg_rowptr);
  cudaFree(h->reg_col); cudaFree(h->reg_val); cudaFree(h->pick_ids);
  cudaFree(h->tile_nq); cudaFree(h->tile_joff); cudaFree(h->tile_w);
  cudaFree(h->tile_ustart); cudaFree(h->tile_uent); cudaFree(h->reg_slot); cudaFree(h->pick_slot);
  blend_tc_destroy(h->blend_tc);
  smpl_fused_destroy(h->fused);
  delete h;
}

static size_t ws_vshaped(int Mb) { return align_up((size_t)Mb * VPITCH * sizeof(float), 256); }
static size_t ws_J(int Mb) { return align_up((size_t)Mb * NJ * 3 * sizeof(float), 256); }
static size_t ws_vposed(int M) { return align_up((size_t)M * VPITCH * sizeof(float), 1024); }

extern "C" size_t hp3d_smpl_pose_blend_workspace_bytes(int M) { return M > 0 ? blend_tc_workspace_bytes(M) : 0; }

// Default: the fused kernel (csrc/smpl_fused.cu: transposed blend GEMM -> tensor-core skinning -> statistics; v_posed never
// in HBM; speed independent of the model's vertex order). HP3D_SMPL=staged selects the round-1 three-kernel path (blend GEMM
// -> v_posed in HBM -> lbs_tile_kernel / generic lbs_kernel -> statistics kernel), kept for comparison and as the path for
// hosts without cuTensorMapEncodeTiled. Measured per 25,600 meshes incl. statistics (profiles/r02r_bench_smpl.jsonl):
// fused 2.25 ms on any vertex order; staged 2.32 ms part-ordered, 3.84 ms shuffled.
static bool use_fused(const hp3d_smpl* h) {
  if (!h->fused) return false;
  const char* e = getenv("HP3D_SMPL");
  return !(e && !strcmp(e, "staged"));
}

extern "C" size_t hp3d_smpl_workspace_bytes(const hp3d_smpl* h, int M, int Mb) {
  if (M <= 0 || Mb <= 0) return 0;
  if (h && use_fused(h)) return smpl_fused_workspace_bytes(M);
  return ws_vshaped(Mb) + ws_J(Mb) + ws_vposed(M) + blend_tc_workspace_bytes(M);
}

extern "C" int hp3d_smpl_layout_info(const hp3d_smpl* h, int* fused, int* permuted, int* tile_joint_sum, int* tile_joint_max) {
  HP3D_ARG(h, "null handle");
  if (fused) *fused = use_fused(h) ? 1 : 0;
  smpl_fused_info(h->fused, permuted, tile_joint_sum, tile_joint_max);
  return 0;
}

extern "C" int hp3d_smpl_shape_blend(const hp3d_smpl* h, const float* betas, int Mb, float* v_shaped, float* J,
                                     void* stream) {
  HP3D_ARG(h && betas && v_shaped && J && Mb > 0, "bad argument");
  dim3 grid(cdiv(VPITCH / 4, 256), Mb);
  shape_blend_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(betas, Mb, h->v_template, h->shapedirs_t, h->J_template,
                                                              h->J_shapedirs, v_shaped, J);
  return launch_status("shape_blend_kernel");
}

// tuning / debugging knobs are read from the environment on every call (no process-global state; getenv is ~100 ns)
static int blend_mode() {            // 0 fp32 CUDA-core; 1 tensor-core
  const char* e = getenv("HP3D_BLEND");
  return (e && !strcmp(e, "fp32")) ? 0 : 1;
}

extern "C" int hp3d_smpl_pose_blend(const hp3d_smpl* h, const float* betas, const float* v_shaped, int Mb,
                                    const float* body_pose, int M, float* v_posed, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  HP3D_ARG(h && betas && v_shaped && body_pose && v_posed && M > 0 && Mb > 0 && M % Mb == 0, "bad argument");
  if (blend_mode() == 1 && h->blend_tc) {
    HP3D_ARG(workspace && workspace_bytes >= blend_tc_workspace_bytes(M), "workspace too small (hp3d_smpl_pose_blend_workspace_bytes)");
    return blend_tc_forward(h->blend_tc, betas, Mb, body_pose, M, v_posed, workspace, (cudaStream_t)stream);
  }
  dim3 grid(cdiv(NV3, PB_BN), cdiv(M, PB_BM));
  pose_blend_fp32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(body_pose, h->posedirs, v_shaped, M, M / Mb, v_posed);
  return launch_status("pose_blend_fp32_kernel");
}

extern "C" int hp3d_smpl_lbs(const hp3d_smpl* h, const float* v_posed, const float* J, int Mb,
                             const float* global_orient, int Mg, const float* body_pose, int M, float* vertices,
                             float* joints, void* stream) {
  HP3D_ARG(h && v_posed && J && global_orient && body_pose && vertices, "null argument");
  HP3D_ARG(M > 0 && Mb > 0 && Mg > 0 && M % Mb == 0 && M % Mg == 0, "M must be a multiple of Mb and Mg");
  int force_generic = 0;
  { const char* e = getenv("HP3D_LBS"); force_generic = (e && !strcmp(e, "generic")) ? 1 : 0; }
  if (h->tile_nq_max > 0 && !force_generic) {
    // HP3D_LBS_MODE (tuning sweeps): 0 = scalar FFMA, 2 meshes ahead; 1 = FFMA2, 2 ahead; 2 = FFMA2, 4 ahead;
    // 3 = scalar FFMA, 4 ahead. Default: LBS_DEFAULT_MODE.
    int mode = LBS_DEFAULT_MODE;
    { const char* e = getenv("HP3D_LBS_MODE"); if (e && *e >= '0' && *e <= '3') mode = *e - '0'; }
    const int grid = cdiv(M, LBS_GMAX);
#define HP3D_LBS_LAUNCH(NQM, F2, PF)                                                                                   \
    lbs_tile_kernel<NQM, F2, PF, LBS_GMAX><<<grid, 256, 0, (cudaStream_t)stream>>>(v_posed, J, Mb, global_orient, Mg, body_pose, M, \
        h->tile_nq, h->tile_joff, h->tile_w, h->tile_ustart, h->tile_uent, h->reg_rowptr, h->reg_slot, h->reg_val,      \
        h->pick_slot, h->tree, vertices, joints)
    if (h->tile_nq_max <= 8 && mode == 1 && grid < 148) {
      // small batch (e.g. the B mode meshes): 2 meshes per CTA so the launch fills the GPU
      lbs_tile_kernel<8, true, 2, 2><<<cdiv(M, 2), 256, 0, (cudaStream_t)stream>>>(v_posed, J, Mb, global_orient, Mg, body_pose, M,
          h->tile_nq, h->tile_joff, h->tile_w, h->tile_ustart, h->tile_uent, h->reg_rowptr, h->reg_slot, h->reg_val,
          h->pick_slot, h->tree, vertices, joints);
    } else if (h->tile_nq_max <= 8) {
      switch (mode) {
        case 0: HP3D_LBS_LAUNCH(8, false, 2); break;
        case 1: HP3D_LBS_LAUNCH(8, true, 2); break;
        case 2: HP3D_LBS_LAUNCH(8, true, 4); break;
        default: HP3D_LBS_LAUNCH(8, false, 4); break;
      }
    } else {
      if (mode == 1 || mode == 2) HP3D_LBS_LAUNCH(NQCAP, true, 2); else HP3D_LBS_LAUNCH(NQCAP, false, 2);
    }
#undef HP3D_LBS_LAUNCH
    return launch_status("lbs_tile_kernel");
  }
  const int grid = std::min(M, 148 * 8);
  lbs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(v_posed, J, Mb, global_orient, Mg, body_pose, M, h->skin_idx,
                                                      h->skin_w, h->skin_k, h->reg_rowptr, h->reg_col, h->reg_val,
                                                      h->pick_ids, h->tree, vertices, joints);
  return launch_status("lbs_kernel");
}

extern "C" int hp3d_smpl_forward(const hp3d_smpl* h, const float* betas, int Mb, const float* global_orient, int Mg,
                                 const float* body_pose, int M, float* vertices, float* joints, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  HP3D_ARG(h && betas && global_orient && body_pose && vertices && workspace, "null argument");
  HP3D_ARG(M > 0 && Mb > 0 && Mg > 0 && M % Mb == 0 && M % Mg == 0, "M must be a multiple of Mb and Mg");
  HP3D_ARG(workspace_bytes >= hp3d_smpl_workspace_bytes(h, M, Mb), "workspace too small");
  if (use_fused(h))
    return smpl_fused_forward(h->fused, betas, Mb, global_orient, Mg, body_pose, M, 0, vertices, joints, nullptr, nullptr, workspace,
                              (cudaStream_t)stream);
  char* ws = (char*)workspace;
  float* v_shaped = (float*)ws; ws += ws_vshaped(Mb);
  float* J = (float*)ws; ws += ws_J(Mb);
  float* v_posed = (float*)ws; ws += ws_vposed(M);
  int rc = hp3d_smpl_shape_blend(h, betas, Mb, v_shaped, J, stream);
  if (rc) return rc;
  rc = hp3d_smpl_pose_blend(h, betas, v_shaped, Mb, body_pose, M, v_posed, ws, blend_tc_workspace_bytes(M), stream);
  if (rc) return rc;
  return hp3d_smpl_lbs(h, v_posed, J, Mb, global_orient, Mg, body_pose, M, vertices, joints, stream);
}

extern "C" int hp3d_smpl_forward_stats(const hp3d_smpl* h, const float* betas, int Mb, const float* global_orient, int Mg,
                                       const float* body_pose, int M, int samples_per_image, float* vertices, float* joints,
                                       float* avg_dist, float* mean_vertices, void* workspace, size_t workspace_bytes, void* stream) {
  HP3D_ARG(h && betas && global_orient && body_pose && vertices && workspace && avg_dist, "null argument");
  HP3D_ARG(M > 0 && Mb > 0 && Mg > 0 && M % Mb == 0 && M % Mg == 0, "M must be a multiple of Mb and Mg");
  HP3D_ARG(samples_per_image > 0 && M % samples_per_image == 0, "M must be a multiple of samples_per_image");
  HP3D_ARG(workspace_bytes >= hp3d_smpl_workspace_bytes(h, M, Mb), "workspace too small");
  if (use_fused(h) && samples_per_image <= 112 && samples_per_image >= 8)      // chunk = image: statistics come out of the same kernel
    return smpl_fused_forward(h->fused, betas, Mb, global_orient, Mg, body_pose, M, samples_per_image, vertices, joints, avg_dist,
                              mean_vertices, workspace, (cudaStream_t)stream);
  int rc = hp3d_smpl_forward(h, betas, Mb, global_orient, Mg, body_pose, M, vertices, joints, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return hp3d_vertex_uncertainty(vertices, M / samples_per_image, samples_per_image, mean_vertices, avg_dist, stream);
}

extern "C" int hp3d_rodrigues(const float* aa, int n, float* R, void* stream) {
  HP3D_ARG(aa && R && n > 0, "bad argument");
  rodrigues_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(aa, n, R);
  return launch_status("rodrigues_kernel");
}

extern "C" int hp3d_rot6d_to_rotmat(const float* x, int n, float* R, void* stream) {
  HP3D_ARG(x && R && n > 0, "bad argument");
  rot6d_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(x, n, R);
  return launch_status("rot6d_kernel");
}

extern "C" int hp3d_vertex_uncertainty(const float* vertices, int B, int N, float* mean_vertices, float* avg_dist,
                                       void* stream) {
  HP3D_ARG(vertices && avg_dist && B > 0 && N > 0, "bad argument");
  int variant = UNC_DEFAULT_VARIANT;   // HP3D_UNC: 1 = 32-vertex tiles, 4-byte loads; 2 = 64-vertex tiles, 8-byte loads; 3 = 64-vertex tiles, cp.async
  { const char* e = getenv("HP3D_UNC"); if (e && *e >= '1' && *e <= '3') variant = *e - '0'; }
  const size_t smem2 = ((size_t)N * UNC2_F + UNC2_F + 256) * sizeof(float);
  if (variant >= 2 && smem2 <= 110 * 1024) {
    dim3 grid(cdiv(NV, UNC2_TV), B);
    if (variant == 2) {
      HP3D_SMEM_OPT_IN(vertex_uncertainty_smem2_kernel<false>, 110 * 1024);      // per device, to the largest size this path uses
      vertex_uncertainty_smem2_kernel<false><<<grid, 256, smem2, (cudaStream_t)stream>>>(vertices, B, N, mean_vertices, avg_dist);
    } else {
      HP3D_SMEM_OPT_IN(vertex_uncertainty_smem2_kernel<true>, 110 * 1024);
      vertex_uncertainty_smem2_kernel<true><<<grid, 256, smem2, (cudaStream_t)stream>>>(vertices, B, N, mean_vertices, avg_dist);
    }
    return launch_status("vertex_uncertainty_smem2_kernel");
  }
  const size_t smem = ((size_t)N * UNC_F + UNC_F + 256) * sizeof(float);
  if (smem <= 220 * 1024) {
    HP3D_SMEM_OPT_IN(vertex_uncertainty_smem_kernel, 220 * 1024);
    dim3 grid(cdiv(NV, UNC_TV), B);
    vertex_uncertainty_smem_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(vertices, B, N, mean_vertices, avg_dist);
    return launch_status("vertex_uncertainty_smem_kernel");
  }
  dim3 grid(cdiv(NV, 256), B);
  vertex_uncertainty_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(vertices, B, N, mean_vertices, avg_dist);
  return launch_status("vertex_uncertainty_kernel");
}
