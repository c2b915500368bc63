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
//   * SKINNING is a second tensor-core contraction: T[vertex][(mesh, 3x4)] = W[vertex][joint] x A[(mesh, 3x4)][joint]^T
//     (K = 24 joints, fp16 hi/lo pairs, three products) -- what smplx itself computes as a dense matmul. Four meshes at a
//     time (N = 48) it lands in a double-buffered TMEM accumulator NEXT to the blend accumulators, lane = the same vertex,
//     so the epilogue is 12 FMAs per (vertex, mesh): out = T[:, :3] v_posed + T[:, 3]. (A first version blended the joint
//     transforms on the CUDA cores like lbs_tile_kernel: 154 warp instructions per 32 vertices x mesh on 8 epilogue warps,
//     3.8 ms per 25,600 meshes against 2.3 ms staged -- profiles/r02o_fused_ncu.md.) W (per 128-vertex group) and A^T (per 8
//     meshes, written by smpl_fk_kernel) are stored in HBM directly in the UMMA no-swizzle K-major core-matrix order, so
//     plain 1-D bulk copies stage them;
//   * a work item = one chunk of <= 112 meshes (the N samples of an image) x 6 vertex groups: the chunk's features stay
//     resident in shared memory (112 KB), the posedirs planes stream through a TMA ring;
//   * 16 epilogue warps: warp = (TMEM lane quarter, half, mesh pair); a half owns one of the two T accumulators, i.e.
//     every other 4-mesh sub-chunk, and its two warps per quarter take two meshes each. Vertices may be re-ordered at create time (a hook for
//     layouts whose order scatters the stores; identity for part-ordered models). Skinned vertices leave through a small
//     per-warp staging buffer as full-sector 8-byte stores (one vertex per lane would otherwise write 4-byte pieces at a
//     12-byte stride);
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
constexpr int MS = 4;                       // meshes per skinning (T) sub-chunk
constexpr int TN = MS * 12;                 // T-GEMM N: 48 columns = 4 meshes x (3 rows x 4)
constexpr int ASUB = 8;                     // meshes per A^T buffer = two T sub-chunks
constexpr int ATPART = TN * 32 * 2;         // one fp16 part of one T sub-chunk, K padded 24 -> 32: 3,072 B
constexpr int ABUF_BYTES = 2 * 2 * ATPART;  // [tsub][hi|lo]: 12,288 B
constexpr int ARING = 2;
constexpr int WPART = GV * 32 * 2;          // one fp16 part of a group's skinning weights [128][32]: 8,192 B
constexpr int WG_BYTES = 2 * WPART;         // 16,384 B
constexpr int TCOL0 = 3 * NPMAX;            // TMEM column of the first T accumulator (336); the second follows at + TN
constexpr int GPI = 6;                      // vertex groups per work item
constexpr int NRANGE = NGRP / GPI;          // 9
static_assert(NRANGE * GPI == NGRP, "vertex groups must split evenly into work items");
constexpr int EPI_WARPS = 16;                // 4 TMEM lane quarters x 2 T accumulators x 2 mesh pairs of a 4-mesh sub-chunk
constexpr int THREADS = 128 + EPI_WARPS * 32;   // 640

struct FusedSmem {
  static constexpr int F_OFF = 0;                                              // [hi kb0..3 | lo kb0..3] feature tiles
  static constexpr int RING_OFF = F_OFF + 8 * FTILE;                           // 114,688
  static constexpr int AT_OFF = RING_OFF + RING * PTILE;                       // 163,840
  static constexpr int W_OFF = AT_OFF + ARING * ABUF_BYTES;                    // 200,704
  static constexpr int STG_OFF = W_OFF + WG_BYTES;                             // 217,088
  static constexpr int SUM_OFF = STG_OFF + EPI_WARPS * 2 * 384;                // 223,232
  static constexpr int BAR_OFF = SUM_OFF + 4 * GV * 16;                        // sums [half * 2 + pair][128] float4
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};
static_assert(FusedSmem::RING_OFF % 1024 == 0 && FusedSmem::TOTAL <= 232448, "shared memory budget");

struct FusedArgs {
  int M, cs, n_chunks, stats, nsu;   // nsu = A^T buffers (8 meshes) per chunk
  int debug;                         // timing experiments only (HP3D_FUSED_DEBUG): 1 no vertex stores, 2 no statistics pass, 4 no blend
                                     // MMAs, 8 no posedirs loads (+4), 16 no skinning arithmetic, 32 no T MMAs
  float inv_scale;
  const uint8_t* AT;         // [n_chunks][nsu][ABUF_BYTES] skinning transforms, fp16 hi/lo, UMMA core-matrix order (smpl_fk_kernel)
  const uint8_t* Wg;         // [NGRP][WG_BYTES] skinning weights of each 128-vertex group, fp16 hi/lo, UMMA core-matrix order
  const float4* vt;          // [NGRP*128] (v_template xyz of the permuted vertex, original vertex index as int bits; -1 = padding)
  const int* tile_base;      // [NWT] original index of the tile's first vertex if its vertices are consecutive, else -1
  float* vertices;           // [M][6890][3]
  float* unc;                // [n_chunks][6890] or null
  float* mean;               // [n_chunks][6890][3] or null
};

struct EpiCtx {
  int lane, chunk_base;
  float* stg;                // this warp's staging buffer [2][96]
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
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load delivered to the same shared-memory offset of BOTH CTAs of the pair; each CTA's own barrier (same offset)
// receives the bytes
__device__ __forceinline__ void tma_load_2d_mc2(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc2(uint64_t* bar) {      // arrives on the same barrier of both CTAs
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// CL2 (opt-in, HP3D_SMPL_CLUSTER=1): the kernel runs as 2-CTA clusters. The two CTAs work on two different chunks (images) but
// walk the SAME vertex groups in lockstep, and every posedirs tile is fetched ONCE for the pair: the CTAs take turns issuing
// the TMA load, which is multicast into both CTAs' ring stage; a stage is re-filled when BOTH CTAs' MMAs have retired it (their
// commits are multicast to both CTAs' empty barriers). It halves the L2 -> SMEM stream per SM -- and measured no faster,
// because that stream is hidden behind the epilogue.
template <bool CL2>
__device__ __forceinline__ void smpl_fused_body(const CUtensorMap& tmFhi, const CUtensorMap& tmFlo, const CUtensorMap& tmPhi,
                                                const CUtensorMap& tmPlo, const FusedArgs& args) {
  using L = FusedSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* f_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* f_empty = f_full + 1;
  uint64_t* ring_full = f_empty + 1;        // [RING]
  uint64_t* ring_empty = ring_full + RING;  // [RING]
  uint64_t* blend_full = ring_empty + RING;
  uint64_t* blend_empty = blend_full + 1;
  uint64_t* a_full = blend_empty + 1;       // [ARING]
  uint64_t* a_empty = a_full + ARING;       // [ARING]
  uint64_t* w_full = a_empty + ARING;
  uint64_t* w_empty = w_full + 1;
  uint64_t* t_full = w_empty + 1;           // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work items: (chunk, range of 6 vertex groups); CL2: the pair takes chunks (2q, 2q + 1) of the same range (an odd chunk
  // count makes the last pair's second CTA repeat the last chunk: identical values to identical addresses)
  const int crank = CL2 ? (int)cluster_ctarank() : 0;
  const int item0 = CL2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_step = CL2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_items = (CL2 ? (args.n_chunks + 1) / 2 : args.n_chunks) * NRANGE;
  auto chunk_of = [&](int item) { return CL2 ? min(2 * (item / NRANGE) + crank, args.n_chunks - 1) : item / NRANGE; };

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmFhi); tma_prefetch_desc(&tmFlo); tma_prefetch_desc(&tmPhi); tma_prefetch_desc(&tmPlo); }
  if (warp == 1 && lane == 0) {
    mbar_init(f_full, 1); mbar_init(f_empty, 1);
    for (int s = 0; s < RING; ++s) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], CL2 ? 2 : 1); }
    mbar_init(blend_full, 1); mbar_init(blend_empty, EPI_WARPS);
    for (int b = 0; b < ARING; ++b) { mbar_init(&a_full[b], 1); mbar_init(&a_empty[b], 1); }
    mbar_init(w_full, 1); mbar_init(w_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], EPI_WARPS / 2); }   // a T buffer belongs to one half of the warps
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_base_slot);
  tc_fence_before_sync();
  __syncthreads();
  if (CL2) cluster_sync_all();             // both CTAs' barriers are initialised before any multicast load / commit
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================================== TMA producer: chunk features (resident per item), posedirs ring
    int stage = 0; uint32_t phase = 0, fphase = 0;
    unsigned tcount = 0;
    for (int item = item0; item < n_items; item += item_step) {
      const int chunk = chunk_of(item), g0 = (item % NRANGE) * GPI;
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
              if (args.debug & 8) { if (elect_one()) mbar_arrive(&ring_full[stage]); ++tcount; if (++stage == RING) { stage = 0; phase ^= 1; } continue; }
              if (elect_one()) mbar_arrive_expect_tx(&ring_full[stage], PTILE);
              if (!CL2) {
                if (elect_one()) tma_load_2d(smem + L::RING_OFF + stage * PTILE, part == 0 ? &tmPhi : &tmPlo, &ring_full[stage], kb * 64,
                                             (g * 3 + plane) * GV);
              } else if ((int)(tcount & 1u) == crank) {      // the pair's CTAs take turns; the tile lands in both
                if (elect_one()) tma_load_2d_mc2(smem + L::RING_OFF + stage * PTILE, part == 0 ? &tmPhi : &tmPlo, &ring_full[stage], kb * 64,
                                                 (g * 3 + plane) * GV);
              }
              ++tcount;
              if (++stage == RING) { stage = 0; phase ^= 1; }
            }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer: blend GEMM of a group, then its skinning GEMMs
    constexpr uint32_t idesc = umma_idesc_f16(NPMAX), idesc_t = umma_idesc_f16(TN);
    int stage = 0; uint32_t phase = 0, fphase = 0, bphase = 0, wphase = 0;
    int ab = 0; uint32_t aphase = 0;
    int tb = 0; uint32_t tphase = 0;
    const uint32_t f_base = smem_u32(smem + L::F_OFF), r_base = smem_u32(smem + L::RING_OFF);
    const uint32_t at_base = smem_u32(smem + L::AT_OFF), w_base = smem_u32(smem + L::W_OFF);
    for (int item = item0; item < n_items; item += item_step) {
      const int chunk = chunk_of(item);
      const int cs = min(args.cs, args.M - chunk * args.cs);
      mbar_wait(f_full, fphase, 43);
      fphase ^= 1;
      tc_fence_after_sync();
      for (int g = 0; g < GPI; ++g) {
        mbar_wait(blend_empty, bphase ^ 1, 44);          // the epilogue has read the previous group's v_posed accumulators
        bphase ^= 1;
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
            if (!(args.debug & 12) && elect_one()) {
              if (!tail) { umma_f16_x4(d_tmem, pd, fh, idesc, kb != 0 ? 1u : 0u); umma_f16_x4(d_tmem, pd, fl, idesc, 1u); }
              else { umma_f16_x2(d_tmem, pd, fh, idesc, 1u); umma_f16_x2(d_tmem, pd, fl, idesc, 1u); }
            }
            if (elect_one()) { if (CL2) umma_commit_mc2(&ring_empty[stage]); else umma_commit(&ring_empty[stage]); }
            if (++stage == RING) { stage = 0; phase ^= 1; }
            // P'_lo tile: P_lo F_hi
            mbar_wait(&ring_full[stage], phase, 45);
            tc_fence_after_sync();
            pd = umma_desc_sw128(r_base + stage * PTILE);
            if (!(args.debug & 12) && elect_one()) { if (!tail) umma_f16_x4(d_tmem, pd, fh, idesc, 1u); else umma_f16_x2(d_tmem, pd, fh, idesc, 1u); }
            if (elect_one()) { if (CL2) umma_commit_mc2(&ring_empty[stage]); else umma_commit(&ring_empty[stage]); }
            if (++stage == RING) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit(blend_full);
        // ---- skinning GEMMs: T[128 vertices][4 meshes x 12] = W_g [128][32] x A^T [48][32]^T per 4-mesh sub-chunk
        mbar_wait(w_full, wphase, 48);
        wphase ^= 1;
        tc_fence_after_sync();
        for (int s = 0; s < args.nsu; ++s) {
          mbar_wait(&a_full[ab], aphase, 49);
          tc_fence_after_sync();
          for (int ts = 0; ts < 2; ++ts) {
            if (s * ASUB + ts * MS >= cs) break;
            mbar_wait(&t_empty[tb], tphase ^ 1, 50);
            tc_fence_after_sync();
            const uint32_t d_t = tmem_base + (uint32_t)(TCOL0 + tb * TN);
            const uint32_t at = at_base + (uint32_t)(ab * ABUF_BYTES + ts * 2 * ATPART);
            if (!(args.debug & 32) && elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {           // K = 32 joints (24 used): two K = 16 steps = core-matrix pairs
                const uint64_t wh = umma_desc_nosw(w_base + ks * 2 * 2048, 2048, 128), wl = umma_desc_nosw(w_base + WPART + ks * 2 * 2048, 2048, 128);
                const uint64_t ah = umma_desc_nosw(at + ks * 2 * 768, 768, 128), al = umma_desc_nosw(at + ATPART + ks * 2 * 768, 768, 128);
                umma_f16(d_t, wh, ah, idesc_t, ks != 0 ? 1u : 0u);
                umma_f16(d_t, wh, al, idesc_t, 1u);
                umma_f16(d_t, wl, ah, idesc_t, 1u);
              }
            }
            if (elect_one()) umma_commit(&t_full[tb]);
            if (++tb == 2) { tb = 0; tphase ^= 1; }
          }
          if (elect_one()) umma_commit(&a_empty[ab]);
          if (++ab == ARING) { ab = 0; aphase ^= 1; }
        }
        if (elect_one()) umma_commit(w_empty);
      }
      if (elect_one()) umma_commit(f_empty);            // the resident features may be replaced once every MMA of the item retired
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================================================== loader of the skinning operands (1-D bulk copies): W per group,
    // A^T per 8 meshes -- both already in UMMA core-matrix order in HBM
    int ab = 0; uint32_t aphase = 0, wphase = 0;
    for (int item = item0; item < n_items; item += item_step) {
      const int chunk = chunk_of(item), g0 = (item % NRANGE) * GPI;
      for (int g = g0; g < g0 + GPI; ++g) {
        mbar_wait(w_empty, wphase ^ 1, 51);
        wphase ^= 1;
        if (elect_one()) mbar_arrive_expect_tx(w_full, WG_BYTES);
        if (elect_one()) bulk_load_1d(smem + L::W_OFF, args.Wg + (size_t)g * WG_BYTES, WG_BYTES, w_full);
        for (int s = 0; s < args.nsu; ++s) {
          mbar_wait(&a_empty[ab], aphase ^ 1, 46);
          if (elect_one()) mbar_arrive_expect_tx(&a_full[ab], ABUF_BYTES);
          if (elect_one()) bulk_load_1d(smem + L::AT_OFF + ab * ABUF_BYTES, args.AT + ((size_t)chunk * args.nsu + s) * ABUF_BYTES, ABUF_BYTES, &a_full[ab]);
          if (++ab == ARING) { ab = 0; aphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue: 16 warps = 4 TMEM lane quarters x 2 T accumulators ("half")
    // x 2 mesh pairs ("sub") of each 4-mesh sub-chunk. The kernel is bound by the latency chain of these warps (TMEM load ->
    // 12 FMAs -> staging -> stores; profiles/r02v_fused_debug.txt: blend MMAs, posedirs loads and T MMAs are free next to
    // it), so the warp count, not the arithmetic, sets the speed.
    EpiCtx c;
    const int quarter = warp & 3, half = ((warp - 4) >> 2) & 1, sub = (warp - 4) >> 3;
    c.lane = lane;
    c.stg = reinterpret_cast<float*>(smem + L::STG_OFF) + (warp - 4) * 2 * 96;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float4* sums = reinterpret_cast<float4*>(smem + L::SUM_OFF);          // [half][128]
    const float inv_scale = args.inv_scale;
    float* const vertices = args.vertices;
    uint32_t bphase = 0, tphase = 0;
    unsigned kcount = 0;                     // T sub-chunks issued so far: sub-chunk k (global count) lives in T buffer k & 1
#pragma unroll 1
    for (int item = item0; item < n_items; item += item_step) {
      const int chunk = chunk_of(item), g0 = (item % NRANGE) * GPI;
      c.chunk_base = chunk * args.cs;
      const int cs = min(args.cs, args.M - c.chunk_base);
#pragma unroll 1
      for (int g = g0; g < g0 + GPI; ++g) {
        const int tile = g * 4 + quarter;
        const float4 vt = args.vt[tile * 32 + lane];
        const int orig = __float_as_int(vt.w);
        const bool valid = orig >= 0;
        const int tbase = args.tile_base[tile];
        const int tcnt = min(32, NV - tile * 32);
        mbar_wait(blend_full, bphase, 47);
        bphase ^= 1;
        tc_fence_after_sync();
        float sx = 0.f, sy = 0.f, sz = 0.f;
        const int nT = (cs + MS - 1) / MS;
        const int k0 = (int)((kcount ^ (unsigned)half) & 1u);   // this half takes the sub-chunks that land in T buffer `half`
#pragma unroll 1
        for (int k = k0; k < nT; k += 2) {               // one T sub-chunk = 4 meshes, both pairs skinned by this warp
          mbar_wait(&t_full[half], tphase, 52);
          tphase ^= 1;
          tc_fence_after_sync();
          const int m0 = k * MS;
          const int cnt = min(MS, cs - m0);
          uint32_t t[24], vx[2], vy[2], vz[2];
          const uint32_t tcol = taddr + (uint32_t)(TCOL0 + half * TN + sub * 24);
#pragma unroll
          for (int q = 0; q < 3; ++q) tmem_ld_32x8(tcol + q * 8, *reinterpret_cast<uint32_t(*)[8]>(&t[q * 8]));
          tmem_ld_32x2(taddr + (uint32_t)(m0 + 2 * sub), vx[0], vx[1]);
          tmem_ld_32x2(taddr + (uint32_t)(NPMAX + m0 + 2 * sub), vy[0], vy[1]);
          tmem_ld_32x2(taddr + (uint32_t)(2 * NPMAX + m0 + 2 * sub), vz[0], vz[1]);
          tmem_ld_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[half]);    // T is in registers: the MMAs of this half's next sub-chunk may start
          if (2 * sub < cnt && !(args.debug & 16)) {
            const bool two = 2 * sub + 1 < cnt;
            float o[6];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float x = fmaf(__uint_as_float(vx[j]), inv_scale, vt.x);
              const float y = fmaf(__uint_as_float(vy[j]), inv_scale, vt.y);
              const float z = fmaf(__uint_as_float(vz[j]), inv_scale, vt.z);
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float* T = reinterpret_cast<const float*>(&t[j * 12 + i * 4]);
                o[j * 3 + i] = fmaf(T[2], z, fmaf(T[1], y, T[0] * x)) + T[3];
              }
            }
            if (!two) { o[3] = 0.f; o[4] = 0.f; o[5] = 0.f; }
            sx += o[0] + o[3]; sy += o[1] + o[4]; sz += o[2] + o[5];
            float* s = c.stg + 3 * lane;
            s[0] = o[0]; s[1] = o[1]; s[2] = o[2];
            s[96] = o[3]; s[97] = o[4]; s[98] = o[5];
            if (!(args.debug & 1)) flush_pair(c, vertices, m0 + 2 * sub, two, tbase, tcnt, orig, valid);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(blend_empty);         // v_posed accumulators free: the next group's MMAs overlap the statistics
        if (args.stats) {
          const int part = half * 2 + sub;                  // 4 partial sums per vertex
          float4* mine = sums + part * GV + quarter * 32 + lane;
          *mine = make_float4(sx, sy, sz, 0.f);
          asm volatile("bar.sync 1, 512;" ::: "memory");
          float tx = 0.f, ty = 0.f, tz = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) { const float4 o = sums[q * GV + quarter * 32 + lane]; tx += o.x; ty += o.y; tz += o.z; }
          const float inv_n = 1.0f / (float)cs;
          const float mx = tx * inv_n, my = ty * inv_n, mz = tz * inv_n;
          // pass 2: mean distance to the mean over this warp's meshes, re-read from L2 (written by this warp a moment ago)
          float dsum = 0.f;
          if (valid && !(args.debug & 2)) {
#pragma unroll 1
            for (int k = k0; k < nT; k += 8) {           // four of this warp's sub-chunks = 8 meshes per iteration: 24 loads in flight
              float px[8], py[8], pz[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int m = (k + 2 * (u >> 1)) * MS + 2 * sub + (u & 1);
                const float* v = vertices + (size_t)(c.chunk_base + min(m, cs - 1)) * NV3 + 3 * orig;
                px[u] = __ldcg(v); py[u] = __ldcg(v + 1); pz[u] = __ldcg(v + 2);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int m = (k + 2 * (u >> 1)) * MS + 2 * sub + (u & 1);
                const float dx = px[u] - mx, dy = py[u] - my, dz = pz[u] - mz;
                if (m < cs) dsum += sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
              }
            }
          }
          asm volatile("bar.sync 1, 512;" ::: "memory");     // everyone has read the position sums
          mine->w = dsum;
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (part == 0 && valid) {
            float d = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) d += sums[q * GV + quarter * 32 + lane].w;
            args.unc[(size_t)chunk * NV + orig] = d * inv_n;
            if (args.mean) { float* mo = args.mean + ((size_t)chunk * NV + orig) * 3; mo[0] = mx; mo[1] = my; mo[2] = mz; }
          }
          asm volatile("bar.sync 1, 512;" ::: "memory");     // the sums buffer may be rewritten by the next group
        }
        kcount += (unsigned)nT;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CL2) cluster_sync_all();             // the peer may still multicast into this CTA's ring / signal its barriers
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<512>(tmem_base); }
}

__global__ void __launch_bounds__(THREADS, 1)
smpl_fused_kernel(const __grid_constant__ CUtensorMap tmFhi, const __grid_constant__ CUtensorMap tmFlo,
                  const __grid_constant__ CUtensorMap tmPhi, const __grid_constant__ CUtensorMap tmPlo,
                  const __grid_constant__ FusedArgs args) {
  smpl_fused_body<false>(tmFhi, tmFlo, tmPhi, tmPlo, args);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
smpl_fused_pair_kernel(const __grid_constant__ CUtensorMap tmFhi, const __grid_constant__ CUtensorMap tmFlo,
                       const __grid_constant__ CUtensorMap tmPhi, const __grid_constant__ CUtensorMap tmPlo,
                       const __grid_constant__ FusedArgs args) {
  smpl_fused_body<true>(tmFhi, tmFlo, tmPhi, tmPlo, args);
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
// walked level by level through shared memory, A_j = [R_j | t_j - R_j J_j] (smplx batch_rigid_transform). The 12 entries of
// A_j are written as fp16 hi/lo pairs straight into the B operand of the skinning GEMM: chunk-major buffers of 8 meshes,
// [T sub-chunk (4 meshes)][hi|lo][k-group (8 joints)][row group][8 rows][8 joints], row = mesh_in_sub-chunk * 12 + entry
// (UMMA no-swizzle K-major core matrices: 128 contiguous bytes = 8 rows x 8 joints). The buffer is zeroed beforehand
// (joints 24..31 and the meshes past the end of a chunk stay zero).
__global__ void __launch_bounds__(256) smpl_fk_kernel(const float* __restrict__ betas, int Mb, const float* __restrict__ global_orient,
                                                      int Mg, const float* __restrict__ body_pose, int M,
                                                      const float* __restrict__ J_template, const float* __restrict__ J_shapedirs,
                                                      FkTree tree, int cs, int nsu, uint8_t* __restrict__ AT, float* __restrict__ joints) {
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
    const int chunk = m / cs, ml = m - chunk * cs;
    const int su = ml / ASUB, ts = (ml % ASUB) / MS, mm = ml % MS;
    uint8_t* buf = AT + ((size_t)chunk * nsu + su) * ABUF_BYTES + (size_t)ts * 2 * ATPART;
    const int koff = (j >> 3) * 768 + (j & 7) * 2;                 // k-group stride 6 row groups x 128 B; 2 B per joint
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float t = G[9 + i] - fmaf(G[i * 3 + 2], Jj[2], fmaf(G[i * 3 + 1], Jj[1], G[i * 3] * Jj[0]));
      const float a4[4] = {G[i * 3], G[i * 3 + 1], G[i * 3 + 2], t};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = mm * 12 + i * 4 + e;
        const int off = koff + (r >> 3) * 128 + (r & 7) * 16;
        const __half h = __float2half_rn(a4[e]);
        *reinterpret_cast<__half*>(buf + off) = h;
        *reinterpret_cast<__half*>(buf + ATPART + off) = __float2half_rn(a4[e] - __half2float(h));
      }
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
  uint8_t* Wg = nullptr;                        // [NGRP][WG_BYTES] skinning weights, fp16 hi/lo, UMMA core-matrix order
  float4* vt = nullptr; int* tile_base = nullptr;
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
  cudaFree(h->p_hi); cudaFree(h->p_lo); cudaFree(h->Wg); cudaFree(h->vt);
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
  // ---- vertex order. Skinning is a dense K = 24 contraction on the tensor cores, so its cost does NOT depend on which
  //      joints the vertices of a tile use: ANY weight layout (part-ordered or not) runs at the same speed in the identity
  //      order, which also keeps every warp's stores consecutive. HP3D_SMPL_ORDER=sorted forces a re-ordering by dominant
  //      joint (tests of the scattered-store path); the per-tile joint counts are reported for information only.
  std::vector<int> order(NV);
  std::iota(order.begin(), order.end(), 0);
  tile_joint_counts(md->lbs_weights, order, h->nq_sum, h->nq_max);
  {
    const char* e = getenv("HP3D_SMPL_ORDER");
    if (e && !strcmp(e, "sorted")) {
      std::vector<int> dom(NV, 0);
      for (int v = 0; v < NV; ++v) {
        double best = -1.0;
        for (int j = 0; j < NJ; ++j) if (md->lbs_weights[(size_t)v * NJ + j] > best) { best = md->lbs_weights[(size_t)v * NJ + j]; dom[v] = j; }
      }
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dom[a] < dom[b]; });
      h->permuted = 1;
      tile_joint_counts(md->lbs_weights, order, h->nq_sum, h->nq_max);
    }
  }
  // ---- per-vertex constants and per-tile store layout
  std::vector<int> tbase(NWT, -1);
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
    bool contig = p1 > p0;
    for (int p = p0 + 1; p < p1; ++p) contig &= (order[p] == order[p0] + (p - p0));
    tbase[t] = contig ? order[p0] : -1;
  }
  // ---- skinning weights of every 128-vertex group as the A operand of the T-GEMM: [group][hi|lo][k-group (8 joints)]
  //      [row group (8 vertices)][8 vertices][8 joints] fp16 (no-swizzle K-major core matrices), K padded 24 -> 32
  std::vector<__half> wg((size_t)NGRP * WG_BYTES / 2, __float2half_rn(0.f));
  for (int g = 0; g < NGRP; ++g)
    for (int i = 0; i < GV; ++i) {
      const int p = g * GV + i;
      if (p >= NV) continue;
      for (int j = 0; j < NJ; ++j) {
        const double w = md->lbs_weights[(size_t)order[p] * NJ + j];
        if (w == 0.0) continue;
        const size_t off = (size_t)g * (WG_BYTES / 2) + (size_t)(j >> 3) * 1024 + (size_t)(i >> 3) * 64 + (size_t)(i & 7) * 8 + (j & 7);   // in halfs
        const __half hv = __float2half_rn((float)w);
        wg[off] = hv;
        wg[off + WPART / 2] = __float2half_rn((float)(w - (double)__half2float(hv)));
      }
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
  rc = rc ? rc : upload((__half**)&h->Wg, wg.data(), wg.size());
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
// A^T buffers: chunks of cs meshes in units of 8 -> at most M/8 + (number of chunks) units; chunks are >= 8 meshes in the
// statistics mode (smaller sample counts take the chunk-of-112 path + the separate statistics kernel)
static size_t fused_at_bytes(int M) { return align_up(((size_t)M / 8 + (size_t)M / 8 + 2) * ABUF_BYTES, 1024); }
size_t smpl_fused_workspace_bytes(int M) { return 2 * fused_feat_bytes(M) + fused_at_bytes(M); }
int smpl_fused_min_samples() { return 8; }

void smpl_fused_info(const void* p, int* permuted, int* nq_sum, int* nq_max) {
  const SmplFused* h = (const SmplFused*)p;
  if (permuted) *permuted = h ? h->permuted : 0;
  if (nq_sum) *nq_sum = h ? h->nq_sum : 0;
  if (nq_max) *nq_max = h ? h->nq_max : 0;
}

// samples_per_image: N in [8, 112] with M % N == 0 -> chunk = image, `unc` [M/N][6890] (and optionally `mean` [M/N][6890][3])
// are written; 0 -> chunks of 112 consecutive meshes, no statistics.
int smpl_fused_forward(void* p, const float* betas, int Mb, const float* global_orient, int Mg, const float* body_pose, int M,
                       int samples_per_image, float* vertices, float* joints, float* unc, float* mean, void* workspace,
                       cudaStream_t stream) {
  SmplFused* h = (SmplFused*)p;
  char* ws = (char*)workspace;
  __half* f_hi = (__half*)ws; ws += fused_feat_bytes(M);
  __half* f_lo = (__half*)ws; ws += fused_feat_bytes(M);
  uint8_t* AT = (uint8_t*)ws;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.stats = (samples_per_image > 0 && unc) ? 1 : 0;
  a.cs = samples_per_image > 0 ? samples_per_image : std::min(M, NPMAX);
  a.n_chunks = cdiv(M, a.cs);
  a.nsu = cdiv(a.cs, ASUB);
  if ((size_t)a.n_chunks * a.nsu * ABUF_BYTES > fused_at_bytes(M)) { set_error("smpl_fused_forward: chunk size %d too small for the workspace", a.cs); return -1; }
  fused_feature_split_kernel<<<(unsigned)(((size_t)M * KP + 255) / 256), 256, 0, stream>>>(body_pose, betas, M / Mb, M, f_hi, f_lo);
  int rc = launch_status("fused_feature_split_kernel");
  if (rc) return rc;
  HP3D_CUDA(cudaMemsetAsync(AT, 0, (size_t)a.n_chunks * a.nsu * ABUF_BYTES, stream));
  smpl_fk_kernel<<<cdiv(M, 8), 256, 0, stream>>>(betas, Mb, global_orient, Mg, body_pose, M, h->J_template, h->J_shapedirs, h->tree, a.cs, a.nsu,
                                                  AT, joints);
  rc = launch_status("smpl_fk_kernel");
  if (rc) return rc;
  CUtensorMap tmFhi, tmFlo;
  const uint64_t dims[2] = {KP, (uint64_t)M};      // rows >= M are out of bounds -> zero filled
  const uint64_t st[1] = {KP * 2};
  const uint32_t box[2] = {64, NPMAX};
  rc = make_tmap_f16(&tmFhi, f_hi, 2, dims, st, box);
  rc = rc ? rc : make_tmap_f16(&tmFlo, f_lo, 2, dims, st, box);
  if (rc) return rc;
  a.inv_scale = h->inv_scale;
  { const char* de = getenv("HP3D_FUSED_DEBUG"); a.debug = de ? atoi(de) : 0; }
  a.AT = AT; a.Wg = h->Wg; a.vt = h->vt; a.tile_base = h->tile_base;
  a.vertices = vertices; a.unc = a.stats ? unc : nullptr; a.mean = a.stats ? mean : nullptr;
  // HP3D_SMPL_CLUSTER=1: 2-CTA clusters with the posedirs tiles multicast to both CTAs (smpl_fused_pair_kernel). Parity-green
  // and measured: 2.243 vs 2.253 ms per 25,600 meshes (profiles/r02u_bench_smpl.jsonl) -- no gain, because the posedirs ring is
  // NOT what bounds the kernel (profiles/r02v_fused_debug.txt: removing the blend MMAs and their loads changes nothing while
  // the epilogue runs); kept opt-in as the building block for a deeper pipeline.
  const char* ce = getenv("HP3D_SMPL_CLUSTER");
  const bool pair = a.n_chunks >= 2 && ce && !strcmp(ce, "1");
  if (pair) {
    const int grid = 2 * std::max(1, std::min(((a.n_chunks + 1) / 2) * NRANGE, h->num_sms / 2));
    HP3D_SMEM_OPT_IN(smpl_fused_pair_kernel, FusedSmem::TOTAL);
    smpl_fused_pair_kernel<<<grid, THREADS, FusedSmem::TOTAL, stream>>>(tmFhi, tmFlo, h->tmPhi, h->tmPlo, a);
  } else {
    const int grid = std::min(a.n_chunks * NRANGE, h->num_sms);
    HP3D_SMEM_OPT_IN(smpl_fused_kernel, FusedSmem::TOTAL);
    smpl_fused_kernel<<<grid, THREADS, FusedSmem::TOTAL, stream>>>(tmFhi, tmFlo, h->tmPhi, h->tmPlo, a);
  }
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
