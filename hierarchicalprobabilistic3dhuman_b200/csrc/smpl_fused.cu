// SMPL forward as ONE tensor-core kernel: pose/shape blend GEMM -> skinning -> (per-vertex sample statistics), sm_100a.
//
// Replaces smplx lbs() (blend_shapes + pose-corrective blend + batch_rigid_transform + skinning)
// as called from reference models/smpl_official.py:29, and utils/sampling_utils.py:189-190 (mean mesh, per-vertex mean
// distance). The staged path (gemm_tc.cu -> smpl.cu) writes v_posed (2.06 GB per 25,600 meshes) to HBM, reads it back in the
// LBS kernel and reads the 2.1 GB of vertices a third time for the statistics: 8.5 GB per step against 2.2 GB of
// algorithmic output. Here v_posed never exists in memory:
//
//   * the blend GEMM runs TRANSPOSED: D[vertex coordinate][mesh] = P'[coordinate][K] x F[mesh][K]^T, K = 207 pose
//     features + 10 betas (fp16 hi/lo pairs, three products, fp32 TMEM). The rows of P' are re-ordered at create time into
//     PLANES -- 128 x-coordinates, then the 128 y-, then the 128 z-coordinates of a 128-vertex group -- so three M = 128
//     accumulators side by side give every TMEM lane (= thread of the epilogue) the x, y, z of ONE vertex for every mesh of
//     the chunk: exactly the layout skinning wants (lane = vertex, loop over meshes, joint transforms broadcast from
//     shared memory, per-lane weights in registers) -- no transpose, no shared-memory staging of v_posed;
//   * a work item = one chunk of <= 112 meshes (the N samples of an image) x 6 vertex groups: the chunk's features stay
//     resident in shared memory (112 KB), the posedirs planes stream through a TMA ring, the per-mesh skinning transforms
//     A (24 x 3x4, from smpl_fk_kernel) stream in 16-mesh sub-chunks through a second ring;
//   * 8 epilogue warps: warp = (TMEM lane quarter, mesh half). Per 32-vertex warp tile the distinct joints are listed at
//     create time (vertices are re-ordered by dominant joint when that shortens the lists, so ANY weight layout gets short
//     lists; tiles with more than 12 joints take a rolled loop -- a per-tile, never a global, fallback) and the body is
//     specialised per joint count like lbs_tile_kernel. Skinned vertices leave through a 1.5 KB per-warp staging buffer as
//     full-sector 8-byte stores (one vertex per lane would otherwise write 4-byte pieces at a 12-byte stride);
//   * statistics: each thread sums its vertex over the chunk's meshes while skinning; after the chunk the two mesh halves
//     are combined, and a second pass re-reads the just-written vertices from L2 (not HBM) for the mean distance.
// The 24 posed joints come from smpl_fk_kernel, the 21 picked + 45 regressed joints from a small gather kernel.
#include "common.cuh"
#include "tc_common.cuh"
#include <vector>
#include <algorithm>
#include <numeric>
#include <math.h>

using namespace hp3d;
using namespace hp3d::tc;

namespace {

constexpr int KP = 224;                     // padded K: row pitch 448 B (207 pose features | 10 betas | unused)
constexpr int KBLKS = 4;                    // 64-wide k-blocks; the last one holds 32 valid columns
constexpr int GV = 128;                     // vertices per group = TMEM lanes
constexpr int NGRP = (NV + GV - 1) / GV;    // 54
constexpr int NWT = NGRP * 4;               // 32-vertex warp tiles: 216
constexpr int PROWS = NGRP * 3 * GV;        // 20,736 plane rows of P'
constexpr int NPMAX = 112;                  // meshes per chunk = MMA N (multiple of 16)
constexpr int PTILE = GV * 128;             // posedirs tile [128 rows][64 k] fp16: 16,384 B
constexpr int FTILE = NPMAX * 128;          // feature tile  [112 rows][64 k] fp16: 14,336 B
constexpr int RING = 3;
constexpr int ASUB = 16;                    // meshes per skinning-transform sub-chunk
constexpr int AMESH = NJ * 12;              // floats per mesh: 24 joints x (3 rows x 4)
constexpr int ASUB_BYTES = ASUB * AMESH * 4;   // 18,432
constexpr int GPI = 6;                      // vertex groups per work item
constexpr int NRANGE = NGRP / GPI;          // 9
static_assert(NRANGE * GPI == NGRP, "vertex groups must split evenly into work items");
constexpr int NQF = 12;                     // specialised bodies for 1..12 joints per warp tile
constexpr int NQTAB = NJ;                   // table width (rolled fallback handles up to all 24 joints)
constexpr int EPI_WARPS = 8;
constexpr int STG_MESHES = 2;
constexpr int THREADS = 384;

struct FusedSmem {
  static constexpr int F_OFF = 0;                                              // [hi kb0..3 | lo kb0..3] feature tiles
  static constexpr int RING_OFF = F_OFF + 8 * FTILE;                           // 114,688
  static constexpr int A_OFF = RING_OFF + RING * PTILE;                        // 163,840
  static constexpr int STG_OFF = A_OFF + 2 * ASUB_BYTES;                       // 200,704
  static constexpr int SUM_OFF = STG_OFF + EPI_WARPS * STG_MESHES * 384;       // 212,992
  static constexpr int BAR_OFF = SUM_OFF + 2 * 2 * GV * 16;                    // sums [parity][half][128] float4: 221,184
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};
static_assert(FusedSmem::RING_OFF % 1024 == 0 && FusedSmem::TOTAL <= 232448, "shared memory budget");

struct FusedArgs {
  int M, cs, n_chunks, stats;
  float inv_scale;
  const float* A;            // [M][288] skinning transforms (smpl_fk_kernel)
  const int* tile_nq;        // [NWT]
  const int* tile_joff;      // [NWT][24] float4 offset of the joint inside one mesh's A block (joint * 3)
  const float* tile_w;       // [NWT][24][32]
  const float4* vt;          // [NGRP*128] (v_template xyz of the permuted vertex, original vertex index as int bits; -1 = padding)
  const int* tile_base;      // [NWT] original index of the tile's first vertex if its vertices are consecutive, else -1
  float* vertices;           // [M][6890][3]
  float* unc;                // [n_chunks][6890] or null
  float* mean;               // [n_chunks][6890][3] or null
};

__device__ __forceinline__ void ffma2(float2& acc, float s, float x, float y) {
  unsigned long long a, b, c, d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(x), "f"(y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(d));
}

struct EpiCtx {
  const FusedArgs* args;
  uint32_t taddr;            // TMEM address of (this warp's lane quarter, column 0)
  int lane, quarter, half, tile, chunk_base, cs;
  const float4* a_smem;      // [2][ASUB][72] float4
  float* stg;                // this warp's staging buffer [STG_MESHES][96]
  uint64_t* a_full; uint64_t* a_empty;
  int abuf; uint32_t aphase;
};

// two staged meshes (m_first, m_first + 1; the second only if `two`) -> HBM
__device__ __forceinline__ void flush_pair(const EpiCtx& c, float* vertices, int m_first, bool two, int tbase, int tcnt, int orig, bool valid) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (j == 1 && !two) break;
    const float* s = c.stg + j * 96;
    float* dst = vertices + (size_t)(c.chunk_base + m_first + j) * NV3;
    if (tbase >= 0 && !(tbase & 1)) {          // consecutive vertices, 8-byte aligned run: full-sector float2 stores
      const int n2 = (3 * tcnt) >> 1;          // tcnt is 32 or 10: 3 * tcnt is even
      float2* d2 = reinterpret_cast<float2*>(dst + 3 * tbase);
      const float2* s2 = reinterpret_cast<const float2*>(s);
      if (c.lane < n2) d2[c.lane] = s2[c.lane];
      if (32 + c.lane < n2) d2[32 + c.lane] = s2[32 + c.lane];
    } else if (valid) {
      dst[3 * orig] = s[3 * c.lane]; dst[3 * orig + 1] = s[3 * c.lane + 1]; dst[3 * orig + 2] = s[3 * c.lane + 2];
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void tmem_ld_32x2(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
}

// Pass 1 of one (chunk, 128-vertex group) for this warp: skin its 32 vertices for its half of every 16-mesh sub-chunk.
// NQ > 0: exactly NQ joints, joint loop unrolled; NQ == 0: rolled loop over `nq` joints (tiles with more than NQF joints).
// The mesh loop is ROLLED, two meshes per iteration, with the TMEM loads of the next pair in flight while the current pair
// is skinned (a first version unrolled 8 meshes x 13 joint-count bodies and ptxas unrolled the sub-chunk loop on top:
// 120k instructions, instruction-cache bound at 4.0 ms per 25,600 meshes).
template <int NQ>
__device__ __forceinline__ void fused_tile_pass1(EpiCtx& c, int nq, float& sx, float& sy, float& sz) {
  const FusedArgs& a = *c.args;
  const int lane = c.lane, t = c.tile;
  const float inv_scale = a.inv_scale;
  float* const vertices = a.vertices;
  const int* const tjoff = a.tile_joff + t * NQTAB;
  const float* const tw = a.tile_w + (size_t)t * NQTAB * 32 + lane;
  constexpr int NW = NQ > 0 ? NQ : 1;
  float w[NW]; int joff[NW];
  if constexpr (NQ > 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) { w[q] = tw[q * 32]; joff[q] = tjoff[q]; }
  }
  const float4 vt = a.vt[t * 32 + lane];
  const int orig = __float_as_int(vt.w);
  const bool valid = orig >= 0;
  const int tbase = a.tile_base[t];
  const int tcnt = min(32, NV - t * 32);
  const int nsc = (c.cs + ASUB - 1) / ASUB;
  auto skin = [&](const float4* Ag, uint32_t xr, uint32_t yr, uint32_t zr, float& ox, float& oy, float& oz) {
    float2 r[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) r[e] = make_float2(0.f, 0.f);
    if constexpr (NQ > 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 r0 = Ag[joff[q]], r1 = Ag[joff[q] + 1], r2 = Ag[joff[q] + 2];
        const float u = w[q];
        ffma2(r[0], u, r0.x, r0.y); ffma2(r[1], u, r0.z, r0.w);
        ffma2(r[2], u, r1.x, r1.y); ffma2(r[3], u, r1.z, r1.w);
        ffma2(r[4], u, r2.x, r2.y); ffma2(r[5], u, r2.z, r2.w);
      }
    } else {
#pragma unroll 1
      for (int q = 0; q < nq; ++q) {
        const int jo = tjoff[q];
        const float u = tw[q * 32];
        const float4 r0 = Ag[jo], r1 = Ag[jo + 1], r2 = Ag[jo + 2];
        ffma2(r[0], u, r0.x, r0.y); ffma2(r[1], u, r0.z, r0.w);
        ffma2(r[2], u, r1.x, r1.y); ffma2(r[3], u, r1.z, r1.w);
        ffma2(r[4], u, r2.x, r2.y); ffma2(r[5], u, r2.z, r2.w);
      }
    }
    const float x = fmaf(__uint_as_float(xr), inv_scale, vt.x);
    const float y = fmaf(__uint_as_float(yr), inv_scale, vt.y);
    const float z = fmaf(__uint_as_float(zr), inv_scale, vt.z);
    ox = fmaf(r[1].x, z, fmaf(r[0].y, y, r[0].x * x)) + r[1].y;
    oy = fmaf(r[3].x, z, fmaf(r[2].y, y, r[2].x * x)) + r[3].y;
    oz = fmaf(r[5].x, z, fmaf(r[4].y, y, r[4].x * x)) + r[5].y;
  };
#pragma unroll 1
  for (int sc = 0; sc < nsc; ++sc) {
    mbar_wait(&c.a_full[c.abuf], c.aphase, 31);
    const int m_lo = sc * ASUB + c.half * 8;
    const int cnt = min(8, c.cs - m_lo);
    if (cnt > 0) {
      const float4* Ab = c.a_smem + (size_t)c.abuf * (ASUB * 72) + (size_t)(c.half * 8) * 72;
      uint32_t x0, x1, y0, y1, z0, z1;
      tmem_ld_32x2(c.taddr + (uint32_t)m_lo, x0, x1);
      tmem_ld_32x2(c.taddr + (uint32_t)(NPMAX + m_lo), y0, y1);
      tmem_ld_32x2(c.taddr + (uint32_t)(2 * NPMAX + m_lo), z0, z1);
#pragma unroll 1
      for (int mi = 0; mi < cnt; mi += 2) {
        tmem_ld_wait();
        const uint32_t cx0 = x0, cx1 = x1, cy0 = y0, cy1 = y1, cz0 = z0, cz1 = z1;
        if (mi + 2 < cnt) {                      // next pair's accumulators in flight while this pair is skinned
          tmem_ld_32x2(c.taddr + (uint32_t)(m_lo + mi + 2), x0, x1);
          tmem_ld_32x2(c.taddr + (uint32_t)(NPMAX + m_lo + mi + 2), y0, y1);
          tmem_ld_32x2(c.taddr + (uint32_t)(2 * NPMAX + m_lo + mi + 2), z0, z1);
        }
        const bool two = mi + 1 < cnt;
        float ox0, oy0, oz0, ox1 = 0.f, oy1 = 0.f, oz1 = 0.f;
        skin(Ab + mi * 72, cx0, cy0, cz0, ox0, oy0, oz0);
        if (two) skin(Ab + (mi + 1) * 72, cx1, cy1, cz1, ox1, oy1, oz1);
        sx += ox0 + ox1; sy += oy0 + oy1; sz += oz0 + oz1;
        float* s = c.stg + 3 * lane;
        s[0] = ox0; s[1] = oy0; s[2] = oz0;
        s[96] = ox1; s[97] = oy1; s[98] = oz1;
        flush_pair(c, vertices, m_lo + mi, two, tbase, tcnt, orig, valid);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&c.a_empty[c.abuf]);
    if (++c.abuf == 2) { c.abuf = 0; c.aphase ^= 1; }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
smpl_fused_kernel(const __grid_constant__ CUtensorMap tmFhi, const __grid_constant__ CUtensorMap tmFlo,
                  const __grid_constant__ CUtensorMap tmPhi, const __grid_constant__ CUtensorMap tmPlo,
                  const __grid_constant__ FusedArgs args) {
  using L = FusedSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* f_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* f_empty = f_full + 1;
  uint64_t* ring_full = f_empty + 1;        // [RING]
  uint64_t* ring_empty = ring_full + RING;  // [RING]
  uint64_t* tmem_full = ring_empty + RING;
  uint64_t* tmem_empty = tmem_full + 1;
  uint64_t* a_full = tmem_empty + 1;        // [2]
  uint64_t* a_empty = a_full + 2;           // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(a_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = args.n_chunks * NRANGE;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmFhi); tma_prefetch_desc(&tmFlo); tma_prefetch_desc(&tmPhi); tma_prefetch_desc(&tmPlo); }
  if (warp == 1 && lane == 0) {
    mbar_init(f_full, 1); mbar_init(f_empty, 1);
    for (int s = 0; s < RING; ++s) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_empty, EPI_WARPS);
    for (int b = 0; b < 2; ++b) { mbar_init(&a_full[b], 1); mbar_init(&a_empty[b], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_base_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================================== TMA producer: chunk features (resident per item), posedirs ring
    int stage = 0; uint32_t phase = 0, fphase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / NRANGE, g0 = (item - chunk * NRANGE) * GPI;
      mbar_wait(f_empty, fphase ^ 1, 41);
      fphase ^= 1;
      if (elect_one()) mbar_arrive_expect_tx(f_full, 8 * FTILE);
      for (int kb = 0; kb < KBLKS; ++kb) {
        if (elect_one()) tma_load_2d(smem + L::F_OFF + kb * FTILE, &tmFhi, f_full, kb * 64, chunk * args.cs);
        if (elect_one()) tma_load_2d(smem + L::F_OFF + (KBLKS + kb) * FTILE, &tmFlo, f_full, kb * 64, chunk * args.cs);
      }
      for (int g = g0; g < g0 + GPI; ++g)
        for (int plane = 0; plane < 3; ++plane)
          for (int kb = 0; kb < KBLKS; ++kb)
            for (int part = 0; part < 2; ++part) {
              mbar_wait(&ring_empty[stage], phase ^ 1, 42);
              if (elect_one()) mbar_arrive_expect_tx(&ring_full[stage], PTILE);
              if (elect_one()) tma_load_2d(smem + L::RING_OFF + stage * PTILE, part == 0 ? &tmPhi : &tmPlo, &ring_full[stage], kb * 64,
                                           (g * 3 + plane) * GV);
              if (++stage == RING) { stage = 0; phase ^= 1; }
            }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(NPMAX);
    int stage = 0; uint32_t phase = 0, fphase = 0, tphase = 0;
    const uint32_t f_base = smem_u32(smem + L::F_OFF), r_base = smem_u32(smem + L::RING_OFF);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      mbar_wait(f_full, fphase, 43);
      fphase ^= 1;
      tc_fence_after_sync();
      for (int g = 0; g < GPI; ++g) {
        mbar_wait(tmem_empty, tphase ^ 1, 44);          // the epilogue has read the previous group's accumulators
        tphase ^= 1;
        tc_fence_after_sync();
        for (int plane = 0; plane < 3; ++plane) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(plane * NPMAX);
          for (int kb = 0; kb < KBLKS; ++kb) {
            const bool tail = kb == KBLKS - 1;           // 32 valid k columns: two K = 16 steps
            const uint64_t fh = umma_desc_sw128(f_base + kb * FTILE), fl = umma_desc_sw128(f_base + (KBLKS + kb) * FTILE);
            // P'_hi tile: P_hi F_hi + P_hi F_lo
            mbar_wait(&ring_full[stage], phase, 45);
            tc_fence_after_sync();
            uint64_t pd = umma_desc_sw128(r_base + stage * PTILE);
            if (elect_one()) {
              if (!tail) { umma_f16_x4(d_tmem, pd, fh, idesc, kb != 0 ? 1u : 0u); umma_f16_x4(d_tmem, pd, fl, idesc, 1u); }
              else { umma_f16_x2(d_tmem, pd, fh, idesc, 1u); umma_f16_x2(d_tmem, pd, fl, idesc, 1u); }
            }
            if (elect_one()) umma_commit(&ring_empty[stage]);
            if (++stage == RING) { stage = 0; phase ^= 1; }
            // P'_lo tile: P_lo F_hi
            mbar_wait(&ring_full[stage], phase, 45);
            tc_fence_after_sync();
            pd = umma_desc_sw128(r_base + stage * PTILE);
            if (elect_one()) { if (!tail) umma_f16_x4(d_tmem, pd, fh, idesc, 1u); else umma_f16_x2(d_tmem, pd, fh, idesc, 1u); }
            if (elect_one()) umma_commit(&ring_empty[stage]);
            if (++stage == RING) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit(tmem_full);
      }
      if (elect_one()) umma_commit(f_empty);            // the resident features may be replaced once every MMA of the item retired
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================================================== skinning-transform loader: 16-mesh sub-chunks, 1-D bulk copies
    int abuf = 0; uint32_t aphase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / NRANGE;
      const int m0 = chunk * args.cs, cs = min(args.cs, args.M - m0);
      const int nsc = (cs + ASUB - 1) / ASUB;
      for (int g = 0; g < GPI; ++g)
        for (int sc = 0; sc < nsc; ++sc) {
          mbar_wait(&a_empty[abuf], aphase ^ 1, 46);
          const uint32_t bytes = (uint32_t)min(ASUB, cs - sc * ASUB) * AMESH * 4;
          if (elect_one()) mbar_arrive_expect_tx(&a_full[abuf], bytes);
          if (elect_one()) bulk_load_1d(smem + L::A_OFF + abuf * ASUB_BYTES, args.A + (size_t)(m0 + sc * ASUB) * AMESH, bytes, &a_full[abuf]);
          if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue: 8 warps = 4 TMEM lane quarters x 2 mesh halves
    EpiCtx c;
    c.args = &args; c.lane = lane; c.quarter = warp & 3; c.half = (warp - 4) >> 2;
    c.taddr = tmem_base + ((uint32_t)(c.quarter * 32) << 16);
    c.a_smem = reinterpret_cast<const float4*>(smem + L::A_OFF);
    c.stg = reinterpret_cast<float*>(smem + L::STG_OFF) + (warp - 4) * STG_MESHES * 96;
    c.a_full = a_full; c.a_empty = a_empty; c.abuf = 0; c.aphase = 0;
    float4* sums = reinterpret_cast<float4*>(smem + L::SUM_OFF);          // [parity][half][128]
    uint32_t tphase = 0;
    int gcount = 0;
#pragma unroll 1
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / NRANGE, g0 = (item - chunk * NRANGE) * GPI;
      c.chunk_base = chunk * args.cs;
      c.cs = min(args.cs, args.M - c.chunk_base);
#pragma unroll 1
      for (int g = g0; g < g0 + GPI; ++g, ++gcount) {        // NOT unrolled: the body holds 13 specialised skinning loops
        c.tile = g * 4 + c.quarter;
        const int nq = args.tile_nq[c.tile];
        mbar_wait(tmem_full, tphase, 47);
        tphase ^= 1;
        tc_fence_after_sync();
        float sx = 0.f, sy = 0.f, sz = 0.f;
        switch (nq) {
          case 1: fused_tile_pass1<1>(c, nq, sx, sy, sz); break;
          case 2: fused_tile_pass1<2>(c, nq, sx, sy, sz); break;
          case 3: fused_tile_pass1<3>(c, nq, sx, sy, sz); break;
          case 4: fused_tile_pass1<4>(c, nq, sx, sy, sz); break;
          case 5: fused_tile_pass1<5>(c, nq, sx, sy, sz); break;
          case 6: fused_tile_pass1<6>(c, nq, sx, sy, sz); break;
          case 7: fused_tile_pass1<7>(c, nq, sx, sy, sz); break;
          case 8: fused_tile_pass1<8>(c, nq, sx, sy, sz); break;
          case 9: fused_tile_pass1<9>(c, nq, sx, sy, sz); break;
          case 10: fused_tile_pass1<10>(c, nq, sx, sy, sz); break;
          case 11: fused_tile_pass1<11>(c, nq, sx, sy, sz); break;
          case 12: fused_tile_pass1<12>(c, nq, sx, sy, sz); break;
          default: fused_tile_pass1<0>(c, nq, sx, sy, sz); break;
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty);          // accumulators free: the next group's MMAs overlap the statistics
        if (args.stats) {
          const float4 vt = args.vt[c.tile * 32 + lane];
          const int orig = __float_as_int(vt.w);
          const bool valid = orig >= 0;
          float4* mine = sums + ((gcount & 1) * 2 + c.half) * GV + c.quarter * 32 + lane;
          float4* other = sums + ((gcount & 1) * 2 + (c.half ^ 1)) * GV + c.quarter * 32 + lane;
          *mine = make_float4(sx, sy, sz, 0.f);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float4 o = *other;
          const float inv_n = 1.0f / (float)c.cs;
          const float mx = (sx + o.x) * inv_n, my = (sy + o.y) * inv_n, mz = (sz + o.z) * inv_n;
          // pass 2: mean distance to the mean over this warp's meshes, re-read from L2 (written by this warp a moment ago)
          float dsum = 0.f;
          if (valid) {
            const int nsc = (c.cs + ASUB - 1) / ASUB;
            for (int sc = 0; sc < nsc; ++sc) {
              const int m_lo = sc * ASUB + c.half * 8;
              const int cnt = min(8, c.cs - m_lo);
#pragma unroll 4
              for (int mi = 0; mi < cnt; ++mi) {
                const float* v = args.vertices + (size_t)(c.chunk_base + m_lo + mi) * NV3 + 3 * orig;
                const float dx = __ldcg(v) - mx, dy = __ldcg(v + 1) - my, dz = __ldcg(v + 2) - mz;
                dsum += sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
              }
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");     // everyone has read the position sums
          mine->w = dsum;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (c.half == 0 && valid) {
            args.unc[(size_t)chunk * NV + orig] = (dsum + other->w) * inv_n;
            if (args.mean) { float* mo = args.mean + ((size_t)chunk * NV + orig) * 3; mo[0] = mx; mo[1] = my; mo[2] = mz; }
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<512>(tmem_base); }
}

// ---------------------------------------------------------------- forward kinematics -> skinning transforms + 24 joints
struct FkTree { int8_t parent[NJ]; int8_t depth[NJ]; int max_depth; };

__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* c) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c[i * 3 + j] = fmaf(a[i * 3 + 2], b[6 + j], fmaf(a[i * 3 + 1], b[3 + j], a[i * 3] * b[j]));
}

// warp = mesh, lane = joint: J = J_template + J_shapedirs beta (the regressor folded at create time), the 24-joint chain
// walked level by level through shared memory, A_j = [R_j | t_j - R_j J_j] (smplx batch_rigid_transform).
__global__ void __launch_bounds__(256) smpl_fk_kernel(const float* __restrict__ betas, int Mb, const float* __restrict__ global_orient,
                                                      int Mg, const float* __restrict__ body_pose, int M,
                                                      const float* __restrict__ J_template, const float* __restrict__ J_shapedirs,
                                                      FkTree tree, float* __restrict__ A, float* __restrict__ joints) {
  __shared__ float sG[8][NJ][12];
  const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  const int repb = M / Mb, repg = M / Mg;
  float R[9], Jj[3] = {0.f, 0.f, 0.f}, rel[3] = {0.f, 0.f, 0.f};
  int par = -1, dep = 99;
  if (j < NJ) {
    const float* src = (j == 0) ? (global_orient + (size_t)(m / repg) * 9) : (body_pose + ((size_t)m * NBJ + (j - 1)) * 9);
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = src[e];
    const float* b = betas + (size_t)(m / repb) * NBETA;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      float a = J_template[j * 3 + e];
#pragma unroll
      for (int l = 0; l < NBETA; ++l) a = fmaf(b[l], J_shapedirs[(j * 3 + e) * NBETA + l], a);
      Jj[e] = a;
    }
    par = tree.parent[j]; dep = tree.depth[j];
  }
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const float pj = __shfl_sync(0xffffffffu, Jj[e], par >= 0 ? par : 0);
    rel[e] = (par >= 0) ? (Jj[e] - pj) : Jj[e];
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
    float4* Am = reinterpret_cast<float4*>(A + (size_t)m * AMESH) + j * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float t = G[9 + i] - fmaf(G[i * 3 + 2], Jj[2], fmaf(G[i * 3 + 1], Jj[1], G[i * 3] * Jj[0]));
      Am[i] = make_float4(G[i * 3], G[i * 3 + 1], G[i * 3 + 2], t);
    }
    if (joints) {
      float* jo = joints + ((size_t)m * NOUTJ + j) * 3;
      jo[0] = G[9]; jo[1] = G[10]; jo[2] = G[11];
    }
  }
}

// 21 picked + 45 regressed joints (smplx VertexJointSelector; reference models/smpl_official.py:30-34) gathered from the
// vertices just written; thread = (mesh, joint), CSR order of accumulation as in lbs_tile_kernel's epilogue.
__global__ void __launch_bounds__(256) smpl_extra_joints_kernel(const float* __restrict__ vertices, int M,
                                                                const int* __restrict__ pick_ids, const int* __restrict__ reg_rowptr,
                                                                const int* __restrict__ reg_col, const float* __restrict__ reg_val,
                                                                float* __restrict__ joints) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= M * (NPICK + NREG)) return;
  const int m = i / (NPICK + NREG), r = i - m * (NPICK + NREG);
  const float* v = vertices + (size_t)m * NV3;
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (r < NPICK) {
    const int id = pick_ids[r];
    ax = v[3 * id]; ay = v[3 * id + 1]; az = v[3 * id + 2];
  } else {
    const int rr = r - NPICK;
    for (int p = reg_rowptr[rr]; p < reg_rowptr[rr + 1]; ++p) {
      const int id = reg_col[p];
      const float w = reg_val[p];
      ax = fmaf(w, v[3 * id], ax); ay = fmaf(w, v[3 * id + 1], ay); az = fmaf(w, v[3 * id + 2], az);
    }
  }
  float* jo = joints + ((size_t)m * NOUTJ + NJ + r) * 3;
  jo[0] = ax; jo[1] = ay; jo[2] = az;
}

// pose features + betas as fp16 hi/lo rows [M][KP] (A'[m] = [R[m] - I | beta[m / rep] | 0...])
__global__ void __launch_bounds__(256) fused_feature_split_kernel(const float* __restrict__ body_pose, const float* __restrict__ betas,
                                                                  int rep, int M, __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * KP) return;
  const int m = (int)(i / KP), k = (int)(i - (size_t)m * KP);
  float v = 0.f;
  if (k < NPF) {
    const int e = k % 9;
    v = body_pose[(size_t)m * NPF + k] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
  } else if (k < NPF + NBETA) {
    v = betas[(size_t)(m / rep) * NBETA + (k - NPF)];
  }
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

struct SmplFused {
  __half *p_hi = nullptr, *p_lo = nullptr;     // [PROWS][KP] posedirs/shapedirs planes (scaled by 2^ex)
  CUtensorMap tmPhi, tmPlo;
  float inv_scale = 1.f;
  int* tile_nq = nullptr; int* tile_joff = nullptr; float* tile_w = nullptr; float4* vt = nullptr; int* tile_base = nullptr;
  float *J_template = nullptr, *J_shapedirs = nullptr;
  int *pick_ids = nullptr, *reg_rowptr = nullptr, *reg_col = nullptr;
  float* reg_val = nullptr;
  FkTree tree;
  int num_sms = 148;
  int permuted = 0, nq_sum = 0, nq_max = 0;
};

}  // namespace

namespace hp3d {

void smpl_fused_destroy(void* p) {
  if (!p) return;
  SmplFused* h = (SmplFused*)p;
  cudaFree(h->p_hi); cudaFree(h->p_lo); cudaFree(h->tile_nq); cudaFree(h->tile_joff); cudaFree(h->tile_w); cudaFree(h->vt);
  cudaFree(h->tile_base); cudaFree(h->J_template); cudaFree(h->J_shapedirs); cudaFree(h->pick_ids); cudaFree(h->reg_rowptr);
  cudaFree(h->reg_col); cudaFree(h->reg_val);
  delete h;
}

// sum over 32-vertex tiles of the number of distinct joints with non-zero weight, for a vertex order
static void tile_joint_counts(const double* W, const std::vector<int>& order, int& sum, int& mx) {
  sum = 0; mx = 0;
  for (int t = 0; t < NWT; ++t) {
    bool used[NJ] = {false};
    for (int i = t * 32; i < std::min(NV, (t + 1) * 32); ++i)
      for (int j = 0; j < NJ; ++j) if (W[(size_t)order[i] * NJ + j] != 0.0) used[j] = true;
    int n = 0;
    for (int j = 0; j < NJ; ++j) n += used[j];
    sum += n; mx = std::max(mx, n);
  }
}

int smpl_fused_create(const hp3d_smpl_model* md, const float* Jt, const float* Js, void** out) {
  *out = nullptr;
  if (!encode_fn()) return 0;                // no tensor-map entry point: the staged path is used
  SmplFused* h = new SmplFused();
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
  h->num_sms = persistent_ctas(h->num_sms);
  for (int j = 0; j < NJ; ++j) { h->tree.parent[j] = (int8_t)md->parents[j]; h->tree.depth[j] = (j == 0) ? 0 : (int8_t)(h->tree.depth[md->parents[j]] + 1); }
  h->tree.max_depth = 0;
  for (int j = 0; j < NJ; ++j) h->tree.max_depth = std::max<int>(h->tree.max_depth, h->tree.depth[j]);
  // ---- vertex order: identity, unless clustering by dominant joint shortens the per-tile joint lists markedly (an
  //      arbitrary weight layout then still gets short lists; runs of consecutive indices survive the stable sort)
  std::vector<int> order(NV);
  std::iota(order.begin(), order.end(), 0);
  int sum_id, mx_id;
  tile_joint_counts(md->lbs_weights, order, sum_id, mx_id);
  {
    std::vector<int> dom(NV, 0), sorted = order;
    for (int v = 0; v < NV; ++v) {
      double best = -1.0;
      for (int j = 0; j < NJ; ++j) if (md->lbs_weights[(size_t)v * NJ + j] > best) { best = md->lbs_weights[(size_t)v * NJ + j]; dom[v] = j; }
    }
    std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return dom[a] < dom[b]; });
    int sum_s, mx_s;
    tile_joint_counts(md->lbs_weights, sorted, sum_s, mx_s);
    const char* e = getenv("HP3D_SMPL_ORDER");     // "identity" / "sorted" force the choice (tests, experiments)
    const bool force_sorted = e && !strcmp(e, "sorted"), force_id = e && !strcmp(e, "identity");
    if (!force_id && (force_sorted || sum_s * 100 < sum_id * 85 || (mx_id > NQF && mx_s < mx_id))) {
      order = sorted; h->permuted = 1; h->nq_sum = sum_s; h->nq_max = mx_s;
    } else { h->nq_sum = sum_id; h->nq_max = mx_id; }
  }
  // ---- tile tables
  std::vector<int> tnq(NWT, 0), tjoff((size_t)NWT * NQTAB, 0), tbase(NWT, -1);
  std::vector<float> tw((size_t)NWT * NQTAB * 32, 0.f);
  std::vector<float4> vt((size_t)NGRP * GV);
  for (int p = 0; p < NGRP * GV; ++p) {
    if (p < NV) {
      const int v = order[p];
      vt[p] = make_float4((float)md->v_template[v * 3], (float)md->v_template[v * 3 + 1], (float)md->v_template[v * 3 + 2], 0.f);
      const int bits = v; memcpy(&vt[p].w, &bits, 4);
    } else { vt[p] = make_float4(0.f, 0.f, 0.f, 0.f); const int bits = -1; memcpy(&vt[p].w, &bits, 4); }
  }
  for (int t = 0; t < NWT; ++t) {
    const int p0 = t * 32, p1 = std::min(NV, p0 + 32);
    int slot[NJ];
    for (int j = 0; j < NJ; ++j) slot[j] = -1;
    for (int p = p0; p < p1; ++p)
      for (int j = 0; j < NJ; ++j) if (md->lbs_weights[(size_t)order[p] * NJ + j] != 0.0) slot[j] = 0;
    int q = 0;
    for (int j = 0; j < NJ; ++j) if (slot[j] == 0) { slot[j] = q; tjoff[(size_t)t * NQTAB + q] = j * 3; ++q; }
    tnq[t] = std::max(q, 1);                 // a tile of all-zero weights still runs the 1-joint body with zero weights
    for (int p = p0; p < p1; ++p)
      for (int j = 0; j < NJ; ++j) {
        const double w = md->lbs_weights[(size_t)order[p] * NJ + j];
        if (w != 0.0) tw[((size_t)t * NQTAB + slot[j]) * 32 + (p - p0)] = (float)w;
      }
    bool contig = p1 > p0;
    for (int p = p0 + 1; p < p1; ++p) contig &= (order[p] == order[p0] + (p - p0));
    tbase[t] = contig ? order[p0] : -1;
  }
  // ---- P' planes: row (g*3 + plane)*128 + i <-> coordinate `plane` of vertex order[g*128 + i]; power-of-two pre-scale
  double mx = 0.0;
  for (size_t i = 0; i < (size_t)NPF * NV3; ++i) mx = std::max(mx, fabs(md->posedirs[i]));
  for (size_t i = 0; i < (size_t)NV3 * NBETA; ++i) mx = std::max(mx, fabs(md->shapedirs[i]));
  int ex = 0;
  if (mx > 0.0) ex = (int)floor(log2(16384.0 / mx));
  const double scale = ldexp(1.0, ex);
  h->inv_scale = (float)ldexp(1.0, -ex);
  std::vector<__half> ph((size_t)PROWS * KP, __float2half_rn(0.f)), pl((size_t)PROWS * KP, __float2half_rn(0.f));
  for (int g = 0; g < NGRP; ++g)
    for (int plane = 0; plane < 3; ++plane)
      for (int i = 0; i < GV; ++i) {
        const int p = g * GV + i;
        if (p >= NV) continue;
        const int c = order[p] * 3 + plane;
        const size_t row = (size_t)(g * 3 + plane) * GV + i;
        for (int k = 0; k < NPF + NBETA; ++k) {
          const double d = (k < NPF) ? md->posedirs[(size_t)k * NV3 + c] : md->shapedirs[(size_t)c * NBETA + (k - NPF)];
          const float v = (float)(d * scale);
          const __half hv = __float2half_rn(v);
          ph[row * KP + k] = hv;
          pl[row * KP + k] = __float2half_rn((float)(d * scale - (double)__half2float(hv)));
        }
      }
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
  int rc = upload(&h->p_hi, ph.data(), ph.size());
  rc = rc ? rc : upload(&h->p_lo, pl.data(), pl.size());
  rc = rc ? rc : upload(&h->tile_nq, tnq.data(), tnq.size());
  rc = rc ? rc : upload(&h->tile_joff, tjoff.data(), tjoff.size());
  rc = rc ? rc : upload(&h->tile_w, tw.data(), tw.size());
  rc = rc ? rc : upload(&h->vt, vt.data(), vt.size());
  rc = rc ? rc : upload(&h->tile_base, tbase.data(), tbase.size());
  rc = rc ? rc : upload(&h->J_template, Jt, (size_t)NJ * 3);
  rc = rc ? rc : upload(&h->J_shapedirs, Js, (size_t)NJ * 3 * NBETA);
  rc = rc ? rc : upload(&h->pick_ids, picks.data(), picks.size());
  rc = rc ? rc : upload(&h->reg_rowptr, rp.data(), rp.size());
  rc = rc ? rc : upload(&h->reg_col, rcol.data(), rcol.size());
  rc = rc ? rc : upload(&h->reg_val, rval.data(), rval.size());
  const uint64_t dims[2] = {KP, PROWS};
  const uint64_t st[1] = {KP * 2};
  const uint32_t box[2] = {64, GV};
  rc = rc ? rc : make_tmap_f16(&h->tmPhi, h->p_hi, 2, dims, st, box);
  rc = rc ? rc : make_tmap_f16(&h->tmPlo, h->p_lo, 2, dims, st, box);
  if (rc) { smpl_fused_destroy(h); return rc; }
  *out = h;
  return 0;
}

static size_t fused_feat_bytes(int M) { return align_up((size_t)M * KP * sizeof(__half), 1024); }
size_t smpl_fused_workspace_bytes(int M) { return 2 * fused_feat_bytes(M) + align_up((size_t)M * AMESH * 4, 1024); }

void smpl_fused_info(const void* p, int* permuted, int* nq_sum, int* nq_max) {
  const SmplFused* h = (const SmplFused*)p;
  if (permuted) *permuted = h ? h->permuted : 0;
  if (nq_sum) *nq_sum = h ? h->nq_sum : 0;
  if (nq_max) *nq_max = h ? h->nq_max : 0;
}

// samples_per_image: N in [1, 112] with M % N == 0 -> chunk = image, `unc` [M/N][6890] (and optionally `mean` [M/N][6890][3])
// are written; 0 -> chunks of 112 consecutive meshes, no statistics.
int smpl_fused_forward(void* p, const float* betas, int Mb, const float* global_orient, int Mg, const float* body_pose, int M,
                       int samples_per_image, float* vertices, float* joints, float* unc, float* mean, void* workspace,
                       cudaStream_t stream) {
  SmplFused* h = (SmplFused*)p;
  char* ws = (char*)workspace;
  __half* f_hi = (__half*)ws; ws += fused_feat_bytes(M);
  __half* f_lo = (__half*)ws; ws += fused_feat_bytes(M);
  float* A = (float*)ws;
  fused_feature_split_kernel<<<(unsigned)(((size_t)M * KP + 255) / 256), 256, 0, stream>>>(body_pose, betas, M / Mb, M, f_hi, f_lo);
  int rc = launch_status("fused_feature_split_kernel");
  if (rc) return rc;
  smpl_fk_kernel<<<cdiv(M, 8), 256, 0, stream>>>(betas, Mb, global_orient, Mg, body_pose, M, h->J_template, h->J_shapedirs, h->tree, A, joints);
  rc = launch_status("smpl_fk_kernel");
  if (rc) return rc;
  CUtensorMap tmFhi, tmFlo;
  const uint64_t dims[2] = {KP, (uint64_t)M};      // rows >= M are out of bounds -> zero filled
  const uint64_t st[1] = {KP * 2};
  const uint32_t box[2] = {64, NPMAX};
  rc = make_tmap_f16(&tmFhi, f_hi, 2, dims, st, box);
  rc = rc ? rc : make_tmap_f16(&tmFlo, f_lo, 2, dims, st, box);
  if (rc) return rc;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.stats = (samples_per_image > 0 && unc) ? 1 : 0;
  a.cs = samples_per_image > 0 ? samples_per_image : std::min(M, NPMAX);
  a.n_chunks = cdiv(M, a.cs);
  a.inv_scale = h->inv_scale;
  a.A = A; a.tile_nq = h->tile_nq; a.tile_joff = h->tile_joff; a.tile_w = h->tile_w; a.vt = h->vt; a.tile_base = h->tile_base;
  a.vertices = vertices; a.unc = a.stats ? unc : nullptr; a.mean = a.stats ? mean : nullptr;
  const int grid = std::min(a.n_chunks * NRANGE, h->num_sms);
  HP3D_SMEM_OPT_IN(smpl_fused_kernel, FusedSmem::TOTAL);
  smpl_fused_kernel<<<grid, THREADS, FusedSmem::TOTAL, stream>>>(tmFhi, tmFlo, h->tmPhi, h->tmPlo, a);
  rc = launch_status("smpl_fused_kernel");
  if (rc) return rc;
  if (joints) {
    const int n = M * (NPICK + NREG);
    smpl_extra_joints_kernel<<<cdiv(n, 256), 256, 0, stream>>>(vertices, M, h->pick_ids, h->reg_rowptr, h->reg_col, h->reg_val, joints);
    rc = launch_status("smpl_extra_joints_kernel");
  }
  return rc;
}

}  // namespace hp3d
