// Matrix-Fisher sampler for sm_100a: one CTA per image, one warp per joint.
//
// Replaces reference utils/sampling_utils.py:74-143 (+ :10-71, + utils/rigid_transform_utils.py:113-133):
// the reference loops over (image, joint) in Python, launching ~12 tiny kernels and one host sync
// per pair. Here every warp owns one (image, joint) matrix-Fisher M(U S V^T) and, entirely in
// registers/shared memory,
//   1. makes (U,S,V) proper (:104-111), builds the Bingham diagonal A and the ACG envelope
//      Omega = 1 + 2A/b, sigma = Omega^-1/2, M* = exp(-(4-b)/2)(4/b)^2 (:118-125),
//   2. proposes 32 unit quaternions per round, x = normalise(sigma * eps) (:51-53), accepts lane l
//      iff w < exp(-x'Ax) / (M* (x'Omega x)^-2) (:56-61), compacts accepted lanes in index order
//      with a ballot + prefix popcount ("first N accepted", :64-65),
//   3. converts to rotation matrices (quat_to_rotmat) and applies R = U_p R_q V_p^T (:139-141),
//   4. stages a [32 samples][J][9] tile in shared memory so that the (B,N,J,3,3) output is
//      written with fully coalesced stores (the tile is contiguous in HBM).
// Noise source: injected eps/w tensors (bit-faithful replay of the reference's draw order) or an
// in-register Philox4x32-10 stream keyed by (seed, image*J+joint, lane, round).
#include "common.cuh"
#include <algorithm>

using namespace hp3d;

namespace {

constexpr int CHUNK = 32;   // samples staged per store

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) {     // (0,1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float r = sqrtf(-2.0f * logf(u01(a)));
  float s, c;
  sincospif(2.0f * u01(b), &s, &c);
  n0 = r * c; n1 = r * s;
}

// Two CTAs per SM (2 x 23 warps = 72 % of the SM's warp slots): the per-joint frames U_p, V_p are only needed by
// the conversion step, so they live in shared memory and the rejection loop keeps ~20 live registers.
__global__ void __launch_bounds__(768, 2) mf_sample_kernel(const float* __restrict__ U, const float* __restrict__ S,
                                                         const float* __restrict__ V, int B, int J, int N, float b,
                                                         float m_star, uint64_t seed, uint64_t offset, uint64_t image_offset,
                                                         const float* __restrict__ eps_in,
                                                         const float* __restrict__ w_in, int n_cand, int max_rounds, int parts,
                                                         float* __restrict__ R_out, unsigned long long* stats) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = warp;                                  // joint
  const int tile_stride = J * 9;
  float* tile = smem;                                  // [CHUNK][J*9]
  float4* stage = reinterpret_cast<float4*>(smem + CHUNK * tile_stride) + warp * 64;   // [64] quats per warp
  float* frames = smem + CHUNK * tile_stride + J * 64 * 4 + warp * 20;                   // U_p [9] | V_p [9] per warp
  // work item = (image, part): `parts` CTAs share an image's N samples (Philox mode only) so that a batch smaller than the
  // GPU's 2 x #SM CTA slots still fills them -- B = 256 images alone occupy 256 of 296 slots and drain unevenly (54.7 %
  // achieved occupancy, profiles/r01z_ncu_full_summary.csv); the injected-noise mode keeps one CTA per image because the
  // reference's "first N accepted" rule runs over ONE candidate sequence.
  const int per_part = (N + parts - 1) / parts;
  for (int item = blockIdx.x; item < B * parts; item += gridDim.x) {
    const int img = item / parts, part = item - img * parts;
    const int n_begin = part * per_part, n_end = min(N, n_begin + per_part);
    const size_t ij = (size_t)img * J + j;
    // ---- proper SVD factors and envelope parameters (warp-uniform, held by every lane)
    float s0, s1, s2;
    {
      float Up[9], Vp[9];
      const float* u = U + ij * 9; const float* v = V + ij * 9; const float* s = S + ij * 3;
#pragma unroll
      for (int e = 0; e < 9; ++e) { Up[e] = u[e]; Vp[e] = v[e]; }
      const float du = Up[0] * (Up[4] * Up[8] - Up[5] * Up[7]) - Up[1] * (Up[3] * Up[8] - Up[5] * Up[6]) +
                       Up[2] * (Up[3] * Up[7] - Up[4] * Up[6]);
      const float dv = Vp[0] * (Vp[4] * Vp[8] - Vp[5] * Vp[7]) - Vp[1] * (Vp[3] * Vp[8] - Vp[5] * Vp[6]) +
                       Vp[2] * (Vp[3] * Vp[7] - Vp[4] * Vp[6]);
      s0 = s[0]; s1 = s[1]; s2 = s[2] * (du * dv);
      Up[2] *= du; Up[5] *= du; Up[8] *= du;
      Vp[2] *= dv; Vp[5] *= dv; Vp[8] *= dv;
      if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 9; ++e) { frames[e] = Up[e]; frames[9 + e] = Vp[e]; }
      }
      __syncwarp();
    }
    const float A1 = 2.f * (s1 + s2), A2 = 2.f * (s0 + s2), A3 = 2.f * (s0 + s1);      // A0 = 0
    const float O0 = 1.f, O1 = 1.f + 2.f * A1 / b, O2 = 1.f + 2.f * A2 / b, O3 = 1.f + 2.f * A3 / b;
    // torch.pow(x, -0.5) on CPU is 1/sqrt(x) (both IEEE-rounded), reference :124
    const float g0 = 1.f, g1 = 1.0f / sqrtf(O1), g2 = 1.0f / sqrtf(O2), g3 = 1.0f / sqrtf(O3);
    int have = 0, round = 0;
    unsigned n_prop = 0, n_acc = 0, n_fail = 0;
    for (int n0 = n_begin; n0 < n_end; n0 += CHUNK) {
      const int need = min(CHUNK, n_end - n0);
      while (have < need) {
        bool valid, accept = false;
        float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
        float e0, e1, e2, e3, wv;
        if (eps_in) {
          const int c = round * 32 + lane;
          valid = c < n_cand;
          if (__all_sync(0xffffffffu, !valid)) break;          // proposals exhausted
          if (valid) {
            const float4 e = reinterpret_cast<const float4*>(eps_in)[ij * n_cand + c];
            e0 = e.x; e1 = e.y; e2 = e.z; e3 = e.w;
            wv = w_in[ij * n_cand + c];
          }
        } else {
          valid = true;
          if (round >= max_rounds) break;
          // Philox subsequence = GLOBAL (image, joint, lane): a rank that owns images [image_offset, image_offset + B) of a
          // sharded batch draws exactly what a single GPU would draw for those images (results independent of world size)
          const uint64_t sub = (ij + image_offset * (uint64_t)J) * 32 + lane;
          const uint64_t cnt = offset + 2ull * ((uint64_t)round + (uint64_t)part * (uint64_t)max_rounds);
          const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
          const uint4 r0 = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)sub, (uint32_t)(sub >> 32)), key);
          const uint4 r1 = philox4x32_10(make_uint4((uint32_t)(cnt + 1), (uint32_t)((cnt + 1) >> 32), (uint32_t)sub, (uint32_t)(sub >> 32)), key);
          box_muller(r0.x, r0.y, e0, e1);
          box_muller(r0.z, r0.w, e2, e3);
          wv = ((float)(r1.x >> 8)) * (1.0f / 16777216.0f);   // [0,1) like torch.rand
        }
        if (valid) {
          const float y0 = g0 * e0, y1 = g1 * e1, y2 = g2 * e2, y3 = g3 * e3;
          const float nrm = sqrtf(y0 * y0 + y1 * y1 + y2 * y2 + y3 * y3);
          const float x0 = y0 / nrm, x1 = y1 / nrm, x2 = y2 / nrm, x3 = y3 / nrm;
          const float p_b = expf(-(x1 * A1 * x1 + x2 * A2 * x2 + x3 * A3 * x3));
          const float qa = x0 * O0 * x0 + x1 * O1 * x1 + x2 * O2 * x2 + x3 * O3 * x3;
          const float p_a = 1.0f / (qa * qa);
          accept = wv < p_b / (m_star * p_a);
          q = make_float4(x0, x1, x2, x3);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, accept);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (accept) stage[have + __popc(mask & ((1u << lane) - 1u))] = q;
        have += __popc(mask);
        n_prop += __popc(vmask); n_acc += __popc(mask);
        ++round;
        __syncwarp();
      }
      if (have < need) {          // ran out of proposals: fill with the mode (identity quaternion) and flag
        if (lane >= have && lane < need) stage[lane] = make_float4(1.f, 0.f, 0.f, 0.f);
        __syncwarp();
        n_fail += 1;
        have = need;
      }
      if (lane < need) {
        float4 q = stage[lane];
        // quat_to_rotmat (reference :113-133) re-normalises
        const float qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        const float w = q.x / qn, x = q.y / qn, y = q.z / qn, z = q.w / qn;
        const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
        const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
        float Rq[9] = {w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                       2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                       2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2};
        float Up[9], Vp[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) { Up[e] = frames[e]; Vp[e] = frames[9 + e]; }
        float T[9];    // Rq * Vp^T
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            T[r * 3 + c] = Rq[r * 3] * Vp[c * 3] + Rq[r * 3 + 1] * Vp[c * 3 + 1] + Rq[r * 3 + 2] * Vp[c * 3 + 2];
        float* o = tile + lane * tile_stride + j * 9;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            o[r * 3 + c] = Up[r * 3] * T[c] + Up[r * 3 + 1] * T[3 + c] + Up[r * 3 + 2] * T[6 + c];
      }
      // carry surplus accepted quaternions to the next chunk
      float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
      const int extra = have - need;
      if (lane < extra) carry = stage[need + lane];
      __syncwarp();
      if (lane < extra) stage[lane] = carry;
      have = extra;
      __syncthreads();
      // coalesced store of the [need][J*9] tile (contiguous in HBM)
      {
        float* dst = R_out + ((size_t)img * N + n0) * tile_stride;
        const int total = need * tile_stride;
        for (int t = threadIdx.x; t < total; t += blockDim.x) dst[t] = tile[t];
      }
      __syncthreads();
    }
    if (stats && lane == 0) {
      atomicAdd(stats + 0, (unsigned long long)n_prop); atomicAdd(stats + 1, (unsigned long long)n_acc);
      if (n_fail) atomicAdd(stats + 2, (unsigned long long)n_fail);
    }
  }
}

}  // namespace

extern "C" int hp3d_mf_sample(const float* U, const float* S, const float* V, int B, int J, int N, float b,
                              uint64_t seed, uint64_t offset, const float* eps, const float* w, int oversampling,
                              float* R_out, unsigned long long* stats, void* stream) {
  return hp3d_mf_sample_sharded(U, S, V, B, J, N, b, seed, offset, 0ull, eps, w, oversampling, R_out, stats, stream);
}

extern "C" int hp3d_mf_sample_sharded(const float* U, const float* S, const float* V, int B, int J, int N, float b,
                                      uint64_t seed, uint64_t offset, uint64_t image_offset, const float* eps, const float* w,
                                      int oversampling, float* R_out, unsigned long long* stats, void* stream) {
  HP3D_ARG(U && S && V && R_out, "null argument");
  HP3D_ARG(B > 0 && N > 0 && J > 0 && J <= 24, "need B>0, N>0, 0<J<=24");
  HP3D_ARG(b > 0.f && b < 4.f, "envelope parameter b must be in (0,4)");
  HP3D_ARG((eps == nullptr) == (w == nullptr), "eps and w must be given together");
  HP3D_ARG(!eps || oversampling > 0, "oversampling must be > 0 with injected noise");
  const float m_star = (float)(exp(-(4.0 - (double)b) / 2.0) * (4.0 / (double)b) * (4.0 / (double)b));
  const size_t smem = (size_t)CHUNK * J * 9 * sizeof(float) + (size_t)J * 64 * sizeof(float4) + (size_t)J * 20 * sizeof(float);
  HP3D_SMEM_OPT_IN(mf_sample_kernel, 96 * 1024);
  const int n_cand = eps ? oversampling * N : 0;
  const int max_rounds = 64 + 16 * ((N + 31) / 32);     // Philox mode: acceptance >= 0.43 => ~2.3 rounds per 32
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int slots = 2 * sms;
  // Philox mode: split every image's samples over `parts` CTAs of >= 32 samples. `parts` depends on N ONLY -- it decides which
  // Philox counters a sample consumes, and results must not depend on how many images a rank holds
  const int parts = eps ? 1 : std::max(1, std::min(N / 32, 4));
  const int grid = std::min(B * parts, slots);
  mf_sample_kernel<<<grid, J * 32, smem, (cudaStream_t)stream>>>(U, S, V, B, J, N, b, m_star, seed, offset, image_offset, eps, w,
                                                                 n_cand, max_rounds, parts, R_out, stats);
  return launch_status("mf_sample_kernel");
}
