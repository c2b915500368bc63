// SMPL pose-corrective blend on the tensor cores (tcgen05 + TMA), sm_100a.
//
//   v_posed[m][c] = v_template[c] + sum_l beta[m/rep][l] shapedirs[c][l] + sum_k (R[m][k] - I[k]) posedirs[k][c]
//   M = B*N meshes (25,600 at B=256, N=100), C = 20,670                                 
// i.e. ONE GEMM with K = 207 pose features + 10 betas + 1 (template) = 218 (padded to 224):
//   A'[m] = [R[m]-I | beta[m/rep] | 1],  B'[c] = [posedirs[:,c] | shapedirs[c,:] | v_template[c]]
// so the shape blend rides along and the epilogue is a pure scaled store.
//
// smplx computes this as one fp32 matmul (lbs(): pose_feature @ posedirs). To hold the 1e-4 parity
// contract on fp16 tensor cores both operands are split into fp16 hi + lo parts and three products
// are accumulated in fp32 (TMEM):  A_hi B_hi + A_lo B_hi + A_hi B_lo   (error ~2^-22 relative);
// posedirs are pre-scaled by a power of two so the lo parts stay in fp16's normal range.
// HP3D_BLEND_PASSES=1 selects the single-product variant (error ~2^-11 of the *offset*, still far
// below 1e-4 of the vertex scale for SMPL-sized correctives).
//
// Kernel anatomy (persistent, warp specialised, one CTA per SM):
//   work unit = (128-mesh M tile) x (chunk of n tiles); the A tile (pose features, all of K, hi+lo =
//   128 KB) stays resident in shared memory for the whole unit while 16 KB B tiles (posedirs) stream
//   through a TMA ring; accumulators are double buffered in TMEM so the epilogue of n-tile i
//   (TMEM -> registers -> smem transpose -> coalesced fp32 stores, + v_shaped) overlaps the MMAs of i+1.
#include "common.cuh"
#include "tc_common.cuh"
#include <vector>
#include <algorithm>
#include <math.h>

using namespace hp3d;
using namespace hp3d::tc;

namespace {

constexpr int KP = 224;              // padded K = 207 + 10 + 1 -> 224 (row pitch of the fp16 operands, 448 B)
constexpr int KBLKS = 4;             // 64-wide k-blocks; the last one holds 32 valid columns
constexpr int BM = 128, BN = 128;          // (N = 256 tiles with 2 stages measured no faster: 1.00 vs 0.97 ms)
constexpr int TILE_BYTES = BM * 64 * 2;    // A k-block tile: 16 KB
constexpr int BTILE_BYTES = BN * 64 * 2;   // B k-block tile: 32 KB
constexpr int STAGES = 4;
constexpr int NPAD = 20736;          // 162 * 128

struct BlendArgs {
  int M, rep;
  int m_tiles, n_tiles, n_chunk, n_chunks;
  float inv_scale;
  float* v_posed;          // [M][VPITCH] (written through tmOut)
};

template <int PASSES>
struct BlendSmem {
  static constexpr int A_TILES = (PASSES == 3) ? 2 * KBLKS : KBLKS;
  static constexpr int A_OFFSET = 0;
  static constexpr int B_OFFSET = A_TILES * TILE_BYTES;
  static constexpr int STG_OFFSET = B_OFFSET + STAGES * BTILE_BYTES;    // epilogue transpose staging
  static constexpr int STG_BYTES = 4 * 2 * 4096;                         // per epilogue warp: two 32x32 fp32 TMA-store tiles
  static_assert(STG_OFFSET % 1024 == 0, "TMA store staging must be 1024-byte aligned (128B swizzle)");
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

__global__ void __launch_bounds__(256) pose_feature_split_kernel(const float* __restrict__ body_pose,
                                                                 const float* __restrict__ betas, int rep, int M,
                                                                 __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * KP) return;
  const int m = (int)(i / KP), k = (int)(i - (size_t)m * KP);
  float v = 0.f;
  if (k < NPF) {
    const int e = k % 9;
    v = body_pose[(size_t)m * NPF + k] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
  } else if (k < NPF + NBETA) {
    v = betas[(size_t)(m / rep) * NBETA + (k - NPF)];
  } else if (k == NPF + NBETA) {
    v = 1.f;
  }
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

template <int PASSES>
__global__ void __launch_bounds__(256, 1)
blend_tc_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ BlendArgs args) {
  using L = BlendSmem<PASSES>;
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the extern array (not an integer round trip) keeps the shared address space visible to ptxas
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint64_t* a_full = tmem_empty + 2;          // [1]
  uint64_t* a_empty = a_full + 1;             // [1]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_units = args.m_tiles * args.n_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAhi); tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(tmem_base_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    {
      int stage = 0; uint32_t phase = 0, a_phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int mt = u / args.n_chunks, chunk = u - mt * args.n_chunks;
        const int nt0 = chunk * args.n_chunk, nt1 = min(nt0 + args.n_chunk, args.n_tiles);
        mbar_wait(a_empty, a_phase ^ 1, 5);
        if (elect_one()) mbar_arrive_expect_tx(a_full, L::A_TILES * TILE_BYTES);
        for (int kb = 0; kb < KBLKS; ++kb) {
          if (elect_one()) tma_load_2d(smem + L::A_OFFSET + kb * TILE_BYTES, &tmAhi, a_full, kb * 64, mt * BM);
          if (PASSES == 3) if (elect_one()) tma_load_2d(smem + L::A_OFFSET + (KBLKS + kb) * TILE_BYTES, &tmAlo, a_full, kb * 64, mt * BM);
        }
        a_phase ^= 1;
        for (int nt = nt0; nt < nt1; ++nt)
          for (int kb = 0; kb < KBLKS; ++kb)
            for (int part = 0; part < (PASSES == 3 ? 2 : 1); ++part) {
              mbar_wait(&empty_bar[stage], phase ^ 1, 1);
              if (elect_one()) mbar_arrive_expect_tx(&full_bar[stage], BTILE_BYTES);
              if (elect_one()) tma_load_2d(smem + L::B_OFFSET + stage * BTILE_BYTES, part == 0 ? &tmBhi : &tmBlo, &full_bar[stage], kb * 64, nt * BN);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      int stage = 0; uint32_t phase = 0, a_phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t a_base = smem_u32(smem + L::A_OFFSET);
      const uint32_t b_base = smem_u32(smem + L::B_OFFSET);
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int mt = u / args.n_chunks, chunk = u - mt * args.n_chunks;
        const int nt0 = chunk * args.n_chunk, nt1 = min(nt0 + args.n_chunk, args.n_tiles);
        mbar_wait(a_full, a_phase, 6);
        a_phase ^= 1;
        tc_fence_after_sync();
        for (int nt = nt0; nt < nt1; ++nt) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < KBLKS; ++kb) {
            const int nmma = (kb < KBLKS - 1) ? 4 : (KP - 64 * (KBLKS - 1)) / 16;   // 4,4,4,2 (K = 224)
            static_assert((KP - 64 * (KBLKS - 1)) / 16 == 2, "tail k-block must hold two K=16 steps");
            const uint64_t ahi = umma_desc_sw128(a_base + kb * TILE_BYTES);
            // B_hi[kb]:  A_hi B_hi (+ A_lo B_hi)
            mbar_wait(&full_bar[stage], phase, 3);
            tc_fence_after_sync();
            uint64_t bd = umma_desc_sw128(b_base + stage * BTILE_BYTES);
            if (elect_one()) { if (nmma == 4) umma_f16_x4(d_tmem, ahi, bd, idesc, kb != 0 ? 1u : 0u); else umma_f16_x2(d_tmem, ahi, bd, idesc, 1u); }
            if (PASSES == 3) {
              const uint64_t alo = umma_desc_sw128(a_base + (KBLKS + kb) * TILE_BYTES);
              if (elect_one()) { if (nmma == 4) umma_f16_x4(d_tmem, alo, bd, idesc, 1u); else umma_f16_x2(d_tmem, alo, bd, idesc, 1u); }
            }
            if (elect_one()) umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            if (PASSES == 3) {   // B_lo[kb]:  A_hi B_lo
              mbar_wait(&full_bar[stage], phase, 3);
              tc_fence_after_sync();
              bd = umma_desc_sw128(b_base + stage * BTILE_BYTES);
              if (elect_one()) { if (nmma == 4) umma_f16_x4(d_tmem, ahi, bd, idesc, 1u); else umma_f16_x2(d_tmem, ahi, bd, idesc, 1u); }
              if (elect_one()) umma_commit(&empty_bar[stage]);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          if (elect_one()) umma_commit(&tmem_full[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (elect_one()) umma_commit(a_empty);     // resident A tile may be overwritten once every MMA of this unit retired
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue: TMEM -> registers -> (x inv_scale) -> 128B-swizzled
    // 32x32 fp32 tile in shared memory -> one TMA store per tile chunk (fully coalesced, asynchronous, no LSU stores)
    const int ew = warp - 4;
    uint8_t* stg_base = smem + L::STG_OFFSET + ew * 2 * 4096;
    int acc = 0; uint32_t acc_phase = 0;
    int sbuf = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int mt = u / args.n_chunks, chunk = u - mt * args.n_chunks;
      const int nt0 = chunk * args.n_chunk, nt1 = min(nt0 + args.n_chunk, args.n_tiles);
      const int m_base = mt * BM + ew * 32;
      for (int nt = nt0; nt < nt1; ++nt) {
        mbar_wait(&tmem_full[acc], acc_phase, 4);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ch * 32, r);
          tmem_ld_wait();
          // the staging buffer used two chunks ago must have been read by its TMA store
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          float4* row = reinterpret_cast<float4*>(stg_base + sbuf * 4096 + lane * 128);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v;
            v.x = __uint_as_float(r[4 * q]) * args.inv_scale; v.y = __uint_as_float(r[4 * q + 1]) * args.inv_scale;
            v.z = __uint_as_float(r[4 * q + 2]) * args.inv_scale; v.w = __uint_as_float(r[4 * q + 3]) * args.inv_scale;
            row[q ^ (lane & 7)] = v;                       // 128B swizzle: 16-byte chunk index XOR (row & 7)
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, stg_base + sbuf * 4096, nt * BN + ch * 32, m_base);   // rows >= M / cols >= pitch are clipped
            tma_store_commit();
          }
          sbuf ^= 1;
        }
        tc_fence_before_sync();
        mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<2 * BN>(tmem_base); }
}

// ---------------------------------------------------------------- CTA-pair variant (opt-in: HP3D_BLEND=pair)
// STATUS: written at the end of round 1 from the profile of blend_tc_kernel (its posedirs ring is latency-bound: 64 KB
// in flight per SM against ~90 KB needed); compiles for sm_100a, NOT yet run on hardware, never selected
// unless HP3D_BLEND=pair is set.
//
// Two CTAs of a cluster (one TPC) own 256 meshes: each keeps the A tile of ITS 128 meshes resident, each loads HALF of
// every B tile (64 of the 128 posedirs columns, 8 KB -> the same 64 KB ring is 8 stages deep), and the leader CTA issues
// `tcgen05.mma.cta_group::2` (M = 256, N = 128): the pair reads each operand byte once from L2 and a stage covers twice the
// MMA time. Protocol (CUTLASS sm100 2-SM conventions): TMA loads of both CTAs signal the LEADER's full barriers
// (.cta_group::2, barrier address with the peer bit cleared), only the leader posts arrive.expect_tx for the pair's bytes;
// tcgen05.commit multicasts to the empty / accumulator-full barriers of BOTH CTAs; the epilogue warps of both CTAs arrive
// on the leader's accumulator-empty barrier (mapa + remote arrive); TMEM is allocated with cta_group::2 by the same warp
// of both CTAs.
constexpr int PAIR_STAGES = 8;
constexpr int BHALF_BYTES = (BN / 2) * 64 * 2;       // 8 KB: this CTA's half of one B k-block tile
struct PairSmem {
  static constexpr int A_TILES = 2 * KBLKS;
  static constexpr int B_OFFSET = A_TILES * TILE_BYTES;
  static constexpr int STG_OFFSET = B_OFFSET + PAIR_STAGES * BHALF_BYTES;
  static constexpr int STG_BYTES = 4 * 2 * 4096;
  static_assert(STG_OFFSET % 1024 == 0, "TMA store staging must be 1024-byte aligned (128B swizzle)");
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a CTA pair: data lands in THIS CTA's shared memory, the bytes are credited to the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on the same barrier of both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 remote;\n\t"
      "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remote];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// `steps` K=16 steps of one 64-wide k-block (descriptors advance by 32 B = +2 per step)
__device__ __forceinline__ void umma_f16_pair_kblock(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc0,
                                                     int steps) {
  for (int k = 0; k < steps; ++k) umma_f16_pair(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, k == 0 ? acc0 : 1u);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_m256(int bn) {
  return (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
blend_pair_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ BlendArgs args) {
  using L = PairSmem;
  constexpr int STAGES_ = PAIR_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);   // used on the leader only
  uint64_t* empty_bar = full_bar + STAGES_;                                 // per CTA (multicast commit)
  uint64_t* tmem_full = empty_bar + STAGES_;     // [2] per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;          // [2] leader's: 2 x 128 epilogue threads
  uint64_t* a_full = tmem_empty + 2;             // leader's: both CTAs' A tiles
  uint64_t* a_empty = a_full + 1;                // per CTA (multicast commit)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int m_pairs = (args.m_tiles + 1) / 2;
  const int num_units = m_pairs * args.n_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAhi); tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES_; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 256); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    fence_barrier_init();
  }
  cluster_sync_all();                              // barrier inits of both CTAs visible before any remote arrive / TMA
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    int stage = 0; uint32_t phase = 0, a_phase = 0;
    for (int u = cid; u < num_units; u += nclusters) {
      const int mp = u / args.n_chunks, chunk = u - mp * args.n_chunks;
      const int nt0 = chunk * args.n_chunk, nt1 = min(nt0 + args.n_chunk, args.n_tiles);
      const int m_row = (mp * 2 + (int)rank) * BM;                    // rows past M are zero-filled by TMA
      mbar_wait(a_empty, a_phase ^ 1, 5);
      if (leader) { if (elect_one()) mbar_arrive_expect_tx(a_full, 2 * L::A_TILES * TILE_BYTES); }
      for (int kb = 0; kb < KBLKS; ++kb) {
        if (elect_one()) tma_load_2d_pair(smem + kb * TILE_BYTES, &tmAhi, a_full, kb * 64, m_row);
        if (elect_one()) tma_load_2d_pair(smem + (KBLKS + kb) * TILE_BYTES, &tmAlo, a_full, kb * 64, m_row);
      }
      a_phase ^= 1;
      for (int nt = nt0; nt < nt1; ++nt)
        for (int kb = 0; kb < KBLKS; ++kb)
          for (int part = 0; part < 2; ++part) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 1);
            if (leader) { if (elect_one()) mbar_arrive_expect_tx(&full_bar[stage], 2 * BHALF_BYTES); }
            if (elect_one()) tma_load_2d_pair(smem + L::B_OFFSET + stage * BHALF_BYTES, part == 0 ? &tmBhi : &tmBlo, &full_bar[stage],
                                              kb * 64, nt * BN + (int)rank * (BN / 2));
            if (++stage == STAGES_) { stage = 0; phase ^= 1; }
          }
    }
    __syncwarp();
  } else if (warp == 1 && leader) {
    // ===================================================== MMA issuer (leader CTA only)
    constexpr uint32_t idesc = umma_idesc_f16_m256(BN);
    int stage = 0; uint32_t phase = 0, a_phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = smem_u32(smem + L::B_OFFSET);
    for (int u = cid; u < num_units; u += nclusters) {
      const int mp = u / args.n_chunks, chunk = u - mp * args.n_chunks;
      const int nt0 = chunk * args.n_chunk, nt1 = min(nt0 + args.n_chunk, args.n_tiles);
      mbar_wait(a_full, a_phase, 6);
      a_phase ^= 1;
      tc_fence_after_sync();
      for (int nt = nt0; nt < nt1; ++nt) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < KBLKS; ++kb) {
          const int steps = (kb < KBLKS - 1) ? 4 : (KP - 64 * (KBLKS - 1)) / 16;
          const uint64_t ahi = umma_desc_sw128(a_base + kb * TILE_BYTES);
          const uint64_t alo = umma_desc_sw128(a_base + (KBLKS + kb) * TILE_BYTES);
          mbar_wait(&full_bar[stage], phase, 3);                       // B_hi halves of both CTAs
          tc_fence_after_sync();
          uint64_t bd = umma_desc_sw128(b_base + stage * BHALF_BYTES);
          if (elect_one()) { umma_f16_pair_kblock(d_tmem, ahi, bd, idesc, kb != 0 ? 1u : 0u, steps); umma_f16_pair_kblock(d_tmem, alo, bd, idesc, 1u, steps); }
          if (elect_one()) umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES_) { stage = 0; phase ^= 1; }
          mbar_wait(&full_bar[stage], phase, 3);                       // B_lo halves
          tc_fence_after_sync();
          bd = umma_desc_sw128(b_base + stage * BHALF_BYTES);
          if (elect_one()) umma_f16_pair_kblock(d_tmem, ahi, bd, idesc, 1u, steps);
          if (elect_one()) umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES_) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (elect_one()) umma_commit_pair(a_empty);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs, own 128 rows of the accumulator)
    const int ew = warp - 4;
    uint8_t* stg_base = smem + L::STG_OFFSET + ew * 2 * 4096;
    int acc = 0; uint32_t acc_phase = 0;
    int sbuf = 0;
    for (int u = cid; u < num_units; u += nclusters) {
      const int mp = u / args.n_chunks, chunk = u - mp * args.n_chunks;
      const int nt0 = chunk * args.n_chunk, nt1 = min(nt0 + args.n_chunk, args.n_tiles);
      const int m_base = (mp * 2 + (int)rank) * BM + ew * 32;
      for (int nt = nt0; nt < nt1; ++nt) {
        mbar_wait(&tmem_full[acc], acc_phase, 4);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ch * 32, r);
          tmem_ld_wait();
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          float4* row = reinterpret_cast<float4*>(stg_base + sbuf * 4096 + lane * 128);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v;
            v.x = __uint_as_float(r[4 * q]) * args.inv_scale; v.y = __uint_as_float(r[4 * q + 1]) * args.inv_scale;
            v.z = __uint_as_float(r[4 * q + 2]) * args.inv_scale; v.w = __uint_as_float(r[4 * q + 3]) * args.inv_scale;
            row[q ^ (lane & 7)] = v;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && m_base < args.M) {
            tma_store_2d(&tmOut, stg_base + sbuf * 4096, nt * BN + ch * 32, m_base);
            tma_store_commit();
          }
          sbuf ^= 1;
        }
        tc_fence_before_sync();
        mbar_arrive_on_cta(&tmem_empty[acc], 0);                       // the leader's MMA warp waits for both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                              // neither CTA may free TMEM / exit while the peer still uses the pair
  if (warp == 2) {
    tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

struct BlendTc {
  __half *b_hi = nullptr, *b_lo = nullptr;   // [NPAD][KP]
  CUtensorMap tmBhi, tmBlo;
  CUtensorMap tmBhiHalf, tmBloHalf;          // box {64, BN/2}: one CTA's half of a B tile (blend_pair_kernel)
  bool pair = false;                          // HP3D_BLEND=pair (experimental)
  float inv_scale = 1.f;
  int num_sms = 148;
  int passes = 3;
};

}  // namespace

namespace hp3d {

void blend_tc_destroy(void* p);

int blend_tc_create(const double* posedirs, const double* shapedirs, const double* v_template, void** out) {
  *out = nullptr;
  if (!encode_fn()) return 0;   // no tensor-map entry point: the fp32 CUDA-core blend is used instead
  BlendTc* h = new BlendTc();
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
  h->num_sms = persistent_ctas(h->num_sms);
  const char* e = getenv("HP3D_BLEND_PASSES");
  h->passes = (e && atoi(e) == 1) ? 1 : 3;
  double mx = 0.0;
  for (size_t i = 0; i < (size_t)NPF * NV3; ++i) mx = std::max(mx, fabs(posedirs[i]));
  for (size_t i = 0; i < (size_t)NV3 * NBETA; ++i) mx = std::max(mx, fabs(shapedirs[i]));
  for (size_t i = 0; i < (size_t)NV3; ++i) mx = std::max(mx, fabs(v_template[i]));
  int ex = 0;
  if (mx > 0.0) ex = (int)floor(log2(16384.0 / mx));
  const double scale = ldexp(1.0, ex);
  h->inv_scale = (float)ldexp(1.0, -ex);
  std::vector<__half> bh((size_t)NPAD * KP, __float2half_rn(0.f)), bl((size_t)NPAD * KP, __float2half_rn(0.f));
  for (int c = 0; c < NV3; ++c)
    for (int k = 0; k < NPF + NBETA + 1; ++k) {
      const double d = (k < NPF) ? posedirs[(size_t)k * NV3 + c]
                     : (k < NPF + NBETA) ? shapedirs[(size_t)c * NBETA + (k - NPF)] : v_template[c];
      const float v = (float)(d * scale);
      const __half hv = __float2half_rn(v);
      bh[(size_t)c * KP + k] = hv;
      bl[(size_t)c * KP + k] = __float2half_rn((float)(d * scale - (double)__half2float(hv)));
    }
  int rc = upload(&h->b_hi, bh.data(), bh.size());
  rc = rc ? rc : upload(&h->b_lo, bl.data(), bl.size());
  const uint64_t dims[2] = {KP, NPAD};
  const uint64_t st[1] = {KP * 2};
  const uint32_t box[2] = {64, BN};
  rc = rc ? rc : make_tmap_f16(&h->tmBhi, h->b_hi, 2, dims, st, box);
  rc = rc ? rc : make_tmap_f16(&h->tmBlo, h->b_lo, 2, dims, st, box);
  const uint32_t box_half[2] = {64, BN / 2};
  rc = rc ? rc : make_tmap_f16(&h->tmBhiHalf, h->b_hi, 2, dims, st, box_half);
  rc = rc ? rc : make_tmap_f16(&h->tmBloHalf, h->b_lo, 2, dims, st, box_half);
  { const char* eb = getenv("HP3D_BLEND"); h->pair = eb && !strcmp(eb, "pair") && h->passes == 3; }
  if (rc) { blend_tc_destroy(h); return rc; }
  *out = h;
  return 0;
}

void blend_tc_destroy(void* p) {
  if (!p) return;
  BlendTc* h = (BlendTc*)p;
  cudaFree(h->b_hi); cudaFree(h->b_lo);
  delete h;
}

size_t blend_tc_workspace_bytes(int M) {
  const size_t Mp = (size_t)cdiv(M, BM) * BM;
  return 2 * align_up(Mp * KP * sizeof(__half), 1024);
}

int blend_tc_forward(void* p, const float* betas, int Mb, const float* body_pose, int M, float* v_posed,
                     void* workspace, cudaStream_t stream) {
  BlendTc* h = (BlendTc*)p;
  const size_t Mp = (size_t)cdiv(M, BM) * BM;
  __half* a_hi = (__half*)workspace;
  __half* a_lo = (__half*)((char*)workspace + align_up(Mp * KP * sizeof(__half), 1024));
  pose_feature_split_kernel<<<(unsigned)(((size_t)M * KP + 255) / 256), 256, 0, stream>>>(body_pose, betas, M / Mb, M, a_hi, a_lo);
  int rc = launch_status("pose_feature_split_kernel");
  if (rc) return rc;
  CUtensorMap tmAhi, tmAlo;
  const uint64_t dims[2] = {KP, (uint64_t)M};      // rows >= M are out of bounds -> zero filled
  const uint64_t st[1] = {KP * 2};
  const uint32_t box[2] = {64, BM};
  rc = make_tmap_f16(&tmAhi, a_hi, 2, dims, st, box);
  rc = rc ? rc : make_tmap_f16(&tmAlo, a_lo, 2, dims, st, box);
  if (rc) return rc;
  BlendArgs a;
  a.M = M; a.rep = M / Mb;
  a.m_tiles = cdiv(M, BM); a.n_tiles = NPAD / BN;
  // n-tiles per work unit (the A tile is re-fetched per unit): 18 amortises it for big M; small M (the mode meshes) gets
  // short chunks so that every SM has a unit
  a.n_chunk = std::max(1, std::min(18, cdiv(a.m_tiles * a.n_tiles, h->num_sms)));
  a.n_chunks = cdiv(a.n_tiles, a.n_chunk);
  a.inv_scale = h->inv_scale; a.v_posed = v_posed;
  CUtensorMap tmOut;
  rc = make_tmap_f32_2d(&tmOut, v_posed, VPITCH, (uint64_t)M, (uint64_t)VPITCH * 4, 32, 32);
  if (rc) return rc;
  const int grid = std::min(a.m_tiles * a.n_chunks, h->num_sms);
  if (h->pair) {       // experimental CTA-pair kernel: clusters of 2, work units of 256 meshes
    const int pairs = (a.m_tiles + 1) / 2;
    a.n_chunk = std::max(1, std::min(18, cdiv(pairs * a.n_tiles, std::max(1, h->num_sms / 2))));
    a.n_chunks = cdiv(a.n_tiles, a.n_chunk);
    const int pgrid = 2 * std::max(1, std::min(pairs * a.n_chunks, h->num_sms / 2));
    HP3D_SMEM_OPT_IN(blend_pair_kernel, PairSmem::TOTAL);
    blend_pair_kernel<<<pgrid, 256, PairSmem::TOTAL, stream>>>(tmAhi, tmAlo, h->tmBhiHalf, h->tmBloHalf, tmOut, a);
    return launch_status("blend_pair_kernel");
  }
  if (h->passes == 3) {
    HP3D_SMEM_OPT_IN(blend_tc_kernel<3>, BlendSmem<3>::TOTAL);
    blend_tc_kernel<3><<<grid, 256, BlendSmem<3>::TOTAL, stream>>>(tmAhi, tmAlo, h->tmBhi, h->tmBlo, tmOut, a);
  } else {
    HP3D_SMEM_OPT_IN(blend_tc_kernel<1>, BlendSmem<1>::TOTAL);
    blend_tc_kernel<1><<<grid, 256, BlendSmem<1>::TOTAL, stream>>>(tmAhi, tmAlo, h->tmBhi, h->tmBlo, tmOut, a);
  }
  return launch_status("blend_tc_kernel");
}

}  // namespace hp3d
