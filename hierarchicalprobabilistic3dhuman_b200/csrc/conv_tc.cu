// ResNet-18 encoder on the 5th-gen tensor cores: every convolution is an im2col-free implicit GEMM
// (tcgen05.mma, fp16 operands, fp32 accumulation in TMEM), operands staged by TMA.
//
// Two arithmetic modes share every kernel through a SPLIT template parameter:
//   fast   (HP3D_ENC_FAST)  one product per k-block, activations fp16 NHWC [N][H][W][C]  -- 3e-4 on the features;
//   split  (HP3D_ENC_SPLIT) the mode that meets the 1e-4 contract (reference models/resnet.py:202-217 is fp32):
//          every fp32 value v is carried as an fp16 pair hi = rn(v), lo = rn(v - hi) and the three products
//          A_hi W_hi + A_hi W_lo + A_lo W_hi are accumulated in the SAME fp32 TMEM accumulator (the dropped A_lo W_lo
//          term is ~2^-22 relative). Weights are BN-folded in fp64, scaled per output channel by a power of two so that
//          max|w| lands in [256, 512) (keeps W_lo in fp16's normal range; undone exactly in the epilogue) and split.
//          MEASURED (profiles/r02c_diag.jsonl): tcgen05.mma adds into its fp32 TMEM accumulator with TRUNCATION (round toward
//          zero) -- every inexact add shrinks the accumulator by half an ulp on average, a systematic relative bias of
//          ~2^-25 per MMA step that reached -5e-5 after layer 4 (1,728 dependent steps per block). So the TMEM chains are
//          kept short: the k loop is cut into CHUNKS of <= 36 MMA steps, each chunk starts a fresh accumulator (the two TMEM
//          buffers alternate per chunk, not per tile), and the epilogue warps add the chunks in fp32 REGISTERS with
//          round-to-nearest while the next chunk's MMAs run.
//          Activations are "split-NHWC": [N][H][W][2C] with channels [hi(C) | lo(C)] in one pixel record, so one
//          tensor map serves both parts (channel coordinate c or C + c) and the epilogue writes both halves of a record.
//
//   D[128 output pixels][BN output channels] += A[128 pixels][64 k] * W[BN][64 k]^T      per k-block
//
// * activations live in HBM as NHWC fp16; a 4-D tensor map {C, W, H, N} with box {64, BW, BH, BIMG}
//   (BW*BH*BIMG = 128) delivers, for filter tap (kh,kw) and channel block c, exactly the A tile of that
//   tap: the box is shifted by (kw-pad, kh-pad) and TMA zero-fills everything outside the image, which
//   *is* the convolution's zero padding. No im2col buffer ever exists.
// * stride-2 convolutions read through "phase" tensor maps (one per input row/column parity: base
//   pointer offset + doubled strides), which turns them into stride-1 problems.
// * the 7x7/2 stem on 18 channels: input channels are padded to 32 and two adjacent pixels form one
//   64-channel "pixel pair", so each filter row is 4 k-blocks (taps -1..6, tap -1 has zero weights).
// * weights are repacked [Cout][K] (K-major, BN folded into them), 2-D tensor map, box {64, BN}.
// * both operand tiles use the 128-byte swizzle (TMA writes it, the UMMA descriptor reads it).
// * warp-specialised persistent kernel (canonical Blackwell GEMM anatomy): warp 0 = TMA producer,
//   warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator, warps 4-7 = epilogue
//   (tcgen05.ld -> +bias (+residual) -> ReLU -> fp16 NHWC store). Two TMEM accumulators let the
//   epilogue of tile i overlap the main loop of tile i+1; a STAGES-deep smem ring feeds the MMAs.
#include "encoder.cuh"
#include "tc_common.cuh"
#include <vector>
#include <algorithm>

using namespace hp3d;
using namespace hp3d::tc;

namespace {

constexpr int MAX_KB = 72;
constexpr int BLOCK_M = 128, BLOCK_K = 64;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB

struct KBlock { int8_t map, dx, dy, pad; int16_t c, bk; };     // A-tile source of one k-block + weight block index

struct ConvTcArgs {
  int num_kb;
  int tiles_m, tiles_n;
  int tiles_x, tiles_y;        // spatial tiles per image
  int BW, BH, BIMG;            // output-pixel box of one M tile (BW*BH*BIMG == 128)
  int N, Ho, Wo, Cout;         // N = number of images actually present
  int Cin;                     // SPLIT: channel offset of the lo half inside an input pixel record
  const float* bias;
  const float* scale;          // SPLIT: 2^-s per output channel (undoes the weight pre-scaling)
  const __half* residual;
  __half* out;
  int relu;
  KBlock kb[MAX_KB];
};

// Epilogue of one 128 x BN accumulator tile, executed by the 4 epilogue warps (thread = TMEM lane = output pixel).
// fast: the residual (skip connection) is prefetched into registers BEFORE waiting for the accumulator so its HBM/L2
// latency hides behind the MMAs of this tile; bias comes from a CTA-resident shared-memory copy.
// (SPLIT mode uses split_chunk_add / split_store_tile below.)
template <int BN, bool SPLIT>
__device__ __forceinline__ void conv_epilogue_tile(uint32_t taddr, const float* __restrict__ bias_s, const float* __restrict__ scale_s,
                                                   const __half* __restrict__ residual, __half* __restrict__ out, size_t off, int cout,
                                                   bool store, int relu, uint64_t* tmem_full_bar, uint32_t parity) {
  if constexpr (!SPLIT) {
    uint4 rv[BN / 8];
    if (residual && store) {
#pragma unroll
      for (int q = 0; q < BN / 8; ++q) rv[q] = reinterpret_cast<const uint4*>(residual + off)[q];
    }
    mbar_wait(tmem_full_bar, parity, 4);
    tc_fence_after_sync();
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + ch * 32, r);
      tmem_ld_wait();
      if (store) {
        uint4* dst = reinterpret_cast<uint4*>(out + off + ch * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[q * 8 + e]) + bias_s[ch * 32 + q * 8 + e];
          if (residual) {
            const __half2* h = reinterpret_cast<const __half2*>(&rv[ch * 4 + q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
          }
          if (relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          uint4 o;
          __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
          dst[q] = o;
        }
      }
    }
  } else {
    static_assert(!SPLIT, "split mode: split_chunk_add + split_store_tile");
  }
}

// SPLIT epilogue, part 0: pull this thread's residual record (hi and lo halves, BN fp16 each) into L2 at the START of the tile.
// The residual is read after the last chunk has been summed; measured (profiles/r02t_ncu_full_summary.csv) the convolutions
// with a skip connection ran 50 % (layer 1) / 20 % (layers 2-3) slower than their twins without one -- the block input had
// long left L2 and the epilogue's tail sat on HBM latency while the MMA warp ran out of free TMEM chunks.
template <int BN>
__device__ __forceinline__ void split_prefetch_residual(const __half* __restrict__ residual, size_t off, int cout, bool store) {
  if (!residual || !store) return;
#pragma unroll
  for (int b = 0; b < BN * 2; b += 128) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(residual + off) + b));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(residual + off + cout) + b));
  }
}

// SPLIT epilogue, part 1: add one chunk's 128 x BN accumulator (TMEM) into the thread's fp32 registers (round to nearest).
template <int BN>
__device__ __forceinline__ void split_chunk_add(uint32_t taddr, float (&accr)[BN], bool first, uint64_t* tmem_full_bar, uint32_t parity) {
  mbar_wait(tmem_full_bar, parity, 4);
  tc_fence_after_sync();
#pragma unroll
  for (int ch = 0; ch < BN / 32; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch * 32, r);
    tmem_ld_wait();
    if (first) {
#pragma unroll
      for (int e = 0; e < 32; ++e) accr[ch * 32 + e] = __uint_as_float(r[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) accr[ch * 32 + e] += __uint_as_float(r[e]);
    }
  }
}
// SPLIT epilogue, part 2: out = acc * 2^-s[c] + bias[c] (+ residual hi + lo), ReLU, hi = rn_f16(v), lo = rn_f16(v - hi)
// stored at channel c and Cout + c of the pixel record.
template <int BN, int TOT, int OFF>
__device__ __forceinline__ void split_store_tile(const float (&accr)[TOT], const float* __restrict__ bias_s, const float* __restrict__ scale_s,
                                                 const __half* __restrict__ residual, __half* __restrict__ out, size_t off, int cout,
                                                 bool store, int relu) {
  if (!store) return;
#pragma unroll
  for (int ch = 0; ch < BN / 32; ++ch) {
    uint4 rh[4], rl[4];
    if (residual) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        rh[q] = reinterpret_cast<const uint4*>(residual + off + ch * 32)[q];
        rl[q] = reinterpret_cast<const uint4*>(residual + off + cout + ch * 32)[q];
      }
    }
    uint4* dst_hi = reinterpret_cast<uint4*>(out + off + ch * 32);
    uint4* dst_lo = reinterpret_cast<uint4*>(out + off + cout + ch * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(accr[OFF + ch * 32 + q * 8 + e], scale_s[ch * 32 + q * 8 + e], bias_s[ch * 32 + q * 8 + e]);
      if (residual) {
        const __half2* h = reinterpret_cast<const __half2*>(&rh[q]);
        const __half2* l = reinterpret_cast<const __half2*>(&rl[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fh = __half22float2(h[e]), fl = __half22float2(l[e]);
          v[2 * e] += fh.x + fl.x; v[2 * e + 1] += fh.y + fl.y;
        }
      }
      if (relu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
      }
      uint4 oh4, ol4;
      __half2* oh = reinterpret_cast<__half2*>(&oh4);
      __half2* ol = reinterpret_cast<__half2*>(&ol4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 f = __half22float2(oh[e]);
        ol[e] = __floats2half2_rn(v[2 * e] - f.x, v[2 * e + 1] - f.y);
      }
      dst_hi[q] = oh4;
      dst_lo[q] = ol4;
    }
  }
}

constexpr int SPLIT_CHUNK_KB = 3;     // conv_tc_kernel, split mode: k-blocks (12 MMA steps each) per TMEM accumulation chunk

template <int BN, int STAGES, bool SPLIT>
struct SmemLayout {
  static constexpr int A_STAGE = SPLIT ? 2 * A_BYTES : A_BYTES;           // A_hi [, A_lo]
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int B_STAGE = SPLIT ? 2 * B_BYTES : B_BYTES;           // W_hi [, W_lo]: ONE TMA box of two planes
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
  static constexpr int TOTAL = BIAS_OFFSET + 2 * 512 * 4 + 1024;  // bias + scale (Cout <= 512) + slack for 1024-byte alignment
};

template <int BN, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvTcArgs args) {
  using L = SmemLayout<BN, STAGES, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the extern array (not an integer round trip) keeps the shared address space visible to ptxas
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);   // [Cout]
  float* scale_s = bias_s + 512;                                     // [Cout] (SPLIT: 2^-s per output channel)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = args.tiles_m * args.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmA3);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
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
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % args.tiles_n, mt = tile / args.tiles_n;
        const int tx = mt % args.tiles_x, ty = (mt / args.tiles_x) % args.tiles_y;
        const int n0 = (mt / (args.tiles_x * args.tiles_y)) * args.BIMG;
        const int ox0 = tx * args.BW, oy0 = ty * args.BH;
        for (int kb = 0; kb < args.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
          uint8_t* b_dst = a_dst + L::A_STAGE;
          if (elect_one()) mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          const KBlock k = args.kb[kb];
          const CUtensorMap* ma = (k.map == 0) ? &tmA0 : (k.map == 1) ? &tmA1 : (k.map == 2) ? &tmA2 : &tmA3;
          if (elect_one()) tma_load_4d(a_dst, ma, &full_bar[stage], k.c, ox0 + k.dx, oy0 + k.dy, n0);
          if (SPLIT) { if (elect_one()) tma_load_4d(a_dst + A_BYTES, ma, &full_bar[stage], k.c + args.Cin, ox0 + k.dx, oy0 + k.dy, n0); }
          if (elect_one()) tma_load_3d(b_dst, &tmB, &full_bar[stage], 0, nt * BN, SPLIT ? 2 * k.bk : k.bk);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single thread)
    {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < args.num_kb; ++kb) {
          // fast: one accumulator per tile; SPLIT: a fresh accumulator every SPLIT_CHUNK_KB k-blocks (short TMEM chains)
          const bool chunk_start = SPLIT ? (kb % SPLIT_CHUNK_KB == 0) : (kb == 0);
          const bool chunk_end = (kb == args.num_kb - 1) || (SPLIT && (kb % SPLIT_CHUNK_KB == SPLIT_CHUNK_KB - 1));
          if (chunk_start) { mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2); tc_fence_after_sync(); }
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(a_addr);
          const uint64_t b_desc = umma_desc_sw128(a_addr + L::A_STAGE);
          if (elect_one()) umma_f16_x4(d_tmem, a_desc, b_desc, idesc, chunk_start ? 0u : 1u);
          if (SPLIT) {
            const uint64_t al_desc = umma_desc_sw128(a_addr + A_BYTES);
            const uint64_t bl_desc = umma_desc_sw128(a_addr + L::A_STAGE + L::B_BYTES);
            if (elect_one()) umma_f16_x4(d_tmem, a_desc, bl_desc, idesc, 1u);
            if (elect_one()) umma_f16_x4(d_tmem, al_desc, b_desc, idesc, 1u);
          }
          if (elect_one()) umma_commit(&empty_bar[stage]);                       // frees the smem slot when these MMAs retire
          if (chunk_end) {
            if (elect_one()) umma_commit(&tmem_full[acc]);                       // accumulator (chunk) complete -> epilogue
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue (128 threads = 128 TMEM lanes)
    const int ew = warp - 4;                    // == warp % 4: the TMEM lane quarter this warp may access
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;
    int acc = 0; uint32_t acc_phase = 0;
    const int img_px = args.BW * args.BH;
    const int rec = SPLIT ? 2 * args.Cout : args.Cout;                           // fp16 elements per output pixel record
    for (int c = et; c < args.Cout; c += 128) { bias_s[c] = args.bias[c]; scale_s[c] = SPLIT ? args.scale[c] : 1.f; }   // once per CTA
    asm volatile("bar.sync 1, 128;" ::: "memory");                               // epilogue-only named barrier
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % args.tiles_n, mt = tile / args.tiles_n;
      const int tx = mt % args.tiles_x, ty = (mt / args.tiles_x) % args.tiles_y;
      const int n0 = (mt / (args.tiles_x * args.tiles_y)) * args.BIMG;
      const int il = row / img_px, rem = row - il * img_px;
      const int yl = rem / args.BW, xl = rem - yl * args.BW;
      const int n = n0 + il;
      const size_t pix = ((size_t)n * args.Ho + (ty * args.BH + yl)) * args.Wo + (tx * args.BW + xl);
      const size_t off = pix * rec + (size_t)nt * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (!SPLIT) {
        conv_epilogue_tile<BN, false>(taddr, bias_s + nt * BN, scale_s + nt * BN, args.residual, args.out, off, args.Cout, n < args.N,
                                      args.relu, &tmem_full[acc], acc_phase);
        tc_fence_before_sync();
        mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      } else {
        float accr[BN];
        split_prefetch_residual<BN>(args.residual, off, args.Cout, n < args.N);
        const int n_chunks = (args.num_kb + SPLIT_CHUNK_KB - 1) / SPLIT_CHUNK_KB;
        for (int c = 0; c < n_chunks; ++c) {
          split_chunk_add<BN>(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN), accr, c == 0, &tmem_full[acc], acc_phase);
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        split_store_tile<BN, BN, 0>(accr, bias_s + nt * BN, scale_s + nt * BN, args.residual, args.out, off, args.Cout, n < args.N, args.relu);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<2 * BN>(tmem_base); }
}

// ---------------------------------------------------------------- halo-patch variant (operand reuse across taps)
// conv_tc_kernel re-fetches the activation tile once per filter tap (9x for 3x3, 28 k-blocks for the stem): ncu
// showed those kernels bound by L2->SMEM operand traffic (10.6 TB/s), not by the tensor pipe. Here the input
// HALO PATCH of an 8(x) x 16(y) output tile is fetched ONCE per 64-channel block with a single TMA box
// {64 ch, 16 px, PH rows} (128-byte pixel records, 128B swizzle), and every filter tap is a *shifted view* of that
// patch: with 8 output pixels per row an 8-row swizzle group is one output row, so group g of the A operand lives at
// start + g * 2048 B (SBO = patch pitch of 16 pixels) and the tap offset (dy*16 + dx) * 128 B only moves the start
// address. The 128B swizzle is a function of the absolute shared-memory address bits (the same reason the usual
// +32 B K-advance inside a swizzle atom works), so TMA-written data and the shifted UMMA view agree.
// (A first version used eight 16-byte-wide un-swizzled planes per patch: correct, but TMA moves 16-byte box rows
// far too slowly -- the stem stayed at 2.1 ms.)  Weights stay 128B-swizzled tiles; they are resident in shared memory when they fit (layer1, 72 KB) and
// streamed through a ring otherwise. The 7x7/2 stem uses two patches (one per input-row parity).
//
// SPLIT mode: a tile's k loop runs over PHASES = 2 x channel blocks; phase (cb, hi) loads the hi patch of block cb and
// multiplies every tap with W_hi AND W_lo (18 weight tiles), phase (cb, lo) loads the lo patch and multiplies with W_hi
// (9 tiles) -- patch-ring stages keep their size, and the weight ring sees one tile per MMA group exactly as in fast
// mode. Streamed weights are laid out in consumption order (27 tiles per channel block: t0.hi t0.lo t1.hi ... t8.lo, then
// t0.hi ... t8.hi); resident weights (layer 1) are the 18 tiles [W_hi t0..t8 | W_lo t0..t8].
constexpr int MAX_TAPS = 28;
struct PatchArgs {
  int tiles_m, tiles_n, tiles_x, tiles_y;
  int N, Ho, Wo, Cout;
  int Cin;                                 // SPLIT: channel offset of the lo half inside an input pixel record
  const float* bias;
  const float* scale;                      // SPLIT: 2^-s per output channel
  const __half* residual;
  __half* out;
  int relu;
  int n_cblk, n_taps, n_patch, patch_tx;   // patch_tx: bytes TMA delivers per patch stage
  int debug;                               // timing experiments only (HP3D_CONV_DEBUG): 1 no stores, 2 no MMAs, 4 no TMA
  int p_pw[2], p_ph[2], p_ox[2], p_oy[2], p_base[2];
  uint16_t t_off[MAX_TAPS];   // view offset of the tap inside its patch, in pixels (128-byte records)
  uint8_t t_patch[MAX_TAPS];
};

template <int BN, bool RESIDENT, int PS, int BS, int PATCH_STAGE_BYTES, int NKB_RES, int TPS>
struct PatchSmem {
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int BSTAGE_BYTES = TPS * B_BYTES;          // one ring stage = TPS consecutive weight tiles, ONE TMA box
  static constexpr int B_OFFSET = (PS * PATCH_STAGE_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_REGION = RESIDENT ? NKB_RES * B_BYTES : BS * BSTAGE_BYTES;
  static constexpr int BAR_OFFSET = B_OFFSET + B_REGION;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 512;
  static constexpr int TOTAL = BIAS_OFFSET + 2 * 512 * 4 + 1024;   // bias + scale
};

// FIXED3: the 3x3/1 geometry (9 taps, one patch of pitch PW) is baked in at compile time, so the MMA issuer's tap
// loop is fully unrolled with immediate view offsets instead of reading the tap table per iteration (the unrolled stem
// kernel below issues an MMA every ~60 cycles, this loop needed ~100).
// One phase of the FIXED3 issue loop: NT weight tiles against one patch stage; tile j multiplies tap j / TDIV.
// SPLIT: a new TMEM accumulation chunk starts every 9 tiles (36 MMA steps): the issuer switches accumulator buffers there.
struct AccState { int acc; uint32_t phase; };
template <int BN, bool RESIDENT, int BS, int TPS, int B_BYTES, int BSTAGE_BYTES, int PW, int NT, int TDIV, bool SPLIT>
__device__ __forceinline__ void patch3_issue_phase(uint32_t tmem_base, uint32_t stage, uint32_t b_base, uint32_t b_res_first, bool res_split,
                                                   uint64_t* bfull, uint64_t* bempty, int& bs, uint32_t& bphase, bool first_phase,
                                                   uint64_t* tmem_full, uint64_t* tmem_empty, AccState& st, int debug) {
  constexpr uint32_t idesc = umma_idesc_f16(BN);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int t = j / TDIV;
    if (SPLIT && j % 9 == 0) { mbar_wait(&tmem_empty[st.acc], st.phase ^ 1, 2); tc_fence_after_sync(); }
    const uint32_t d_tmem = tmem_base + (uint32_t)(st.acc * BN);
    uint32_t b_addr;
    if (RESIDENT) {
      // resident layout: fast = [cb][tap]; split = [W_hi taps | W_lo taps]; b_res_first = first tile of this phase
      const int idx = (TDIV == 2 && res_split) ? ((j & 1) * 9 + t) : t;
      b_addr = b_base + (b_res_first + (uint32_t)idx) * (uint32_t)B_BYTES;
    } else {
      if (j % TPS == 0) { mbar_wait(&bfull[bs], bphase, 15); tc_fence_after_sync(); }
      b_addr = b_base + (uint32_t)(bs * BSTAGE_BYTES + (j % TPS) * B_BYTES);
    }
    const uint32_t a_view = stage + (uint32_t)(((t / 3) * PW + (t % 3)) * 128);
    const uint64_t a_desc = umma_desc_sw128_sbo(a_view, PW * 128u);
    const uint64_t b_desc = umma_desc_sw128(b_addr);
    if (!(debug & 2)) {
      const bool overwrite = SPLIT ? (j % 9 == 0) : (first_phase && j == 0);
      if (elect_one()) umma_f16_x4(d_tmem, a_desc, b_desc, idesc, overwrite ? 0u : 1u);
    }
    if (!RESIDENT && (j % TPS) == TPS - 1) { if (elect_one()) umma_commit(&bempty[bs]); if (++bs == BS) { bs = 0; bphase ^= 1; } }
    if (SPLIT && j % 9 == 8) {
      if (elect_one()) umma_commit(&tmem_full[st.acc]);
      if (++st.acc == 2) { st.acc = 0; st.phase ^= 1; }
    }
  }
}

template <int BN, bool RESIDENT, int PS, int BS, int PATCH_STAGE_BYTES, int NKB_RES, int TPS, bool FIXED3, bool SPLIT, int PW>
__global__ void __launch_bounds__(256, 1)
conv_patch_kernel(const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmP1,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ PatchArgs args) {
  static_assert(!SPLIT || FIXED3, "the split mode of the patch kernel covers the 3x3/1 geometry only");
  using L = PatchSmem<BN, RESIDENT, PS, BS, PATCH_STAGE_BYTES, NKB_RES, TPS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* pempty = pfull + PS;
  uint64_t* bfull = pempty + PS;          // [BS] (streamed) / [1] (resident)
  uint64_t* bempty = bfull + BS;
  uint64_t* tmem_full = bempty + BS;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
  float* scale_s = bias_s + 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = args.tiles_m * args.tiles_n;
  constexpr int BW = 8, BH = 16;
  const int n_phase = SPLIT ? 2 * args.n_cblk : args.n_cblk;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmP0); tma_prefetch_desc(&tmP1); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PS; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
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
      if (RESIDENT) {
        const int nkb = SPLIT ? 18 : args.n_taps * args.n_cblk;
        if (elect_one()) mbar_arrive_expect_tx(&bfull[0], nkb * L::B_BYTES);
        if (elect_one()) tma_load_3d(smem + L::B_OFFSET, &tmB, &bfull[0], 0, 0, 0);     // the whole weight tensor: one box
      }
      int ps = 0, bs = 0; uint32_t pphase = 0, bphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % args.tiles_n, mt = tile / args.tiles_n;
        const int tx = mt % args.tiles_x, ty = (mt / args.tiles_x) % args.tiles_y;
        const int n = mt / (args.tiles_x * args.tiles_y);
        const int ox0 = tx * BW, oy0 = ty * BH;
        for (int ph = 0; ph < n_phase; ++ph) {
          const int cb = SPLIT ? (ph >> 1) : ph;
          const bool lo = SPLIT && (ph & 1);
          mbar_wait(&pempty[ps], pphase ^ 1, 11);
          uint8_t* stage = smem + ps * PATCH_STAGE_BYTES;
          if (args.debug & 4) { if (elect_one()) mbar_arrive(&pfull[ps]); }
          else {
            if (elect_one()) mbar_arrive_expect_tx(&pfull[ps], args.patch_tx);
            for (int p = 0; p < args.n_patch; ++p)
              if (elect_one()) tma_load_4d(stage + args.p_base[p], p == 0 ? &tmP0 : &tmP1, &pfull[ps], cb * 64 + (lo ? args.Cin : 0), ox0 + args.p_ox[p], oy0 + args.p_oy[p], n);
          }
          if (++ps == PS) { ps = 0; pphase ^= 1; }
          if (!RESIDENT) {
            // weight tiles of this phase, in consumption order (fast: n_taps per channel block; split: 18 then 9 of 27)
            const int ntile = SPLIT ? (lo ? 9 : 18) : args.n_taps;
            const int first = SPLIT ? cb * 27 + (lo ? 18 : 0) : cb * args.n_taps;
            for (int t = 0; t < ntile; t += TPS) {
              mbar_wait(&bempty[bs], bphase ^ 1, 12);
              if (args.debug & 4) { if (elect_one()) mbar_arrive(&bfull[bs]); }
              else {
                if (elect_one()) mbar_arrive_expect_tx(&bfull[bs], L::BSTAGE_BYTES);
                if (elect_one()) tma_load_3d(smem + L::B_OFFSET + bs * L::BSTAGE_BYTES, &tmB, &bfull[bs], 0, nt * BN, first + t);
              }
              if (++bs == BS) { bs = 0; bphase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      int ps = 0, bs = 0; uint32_t pphase = 0, bphase = 0;
      AccState st = {0, 0u};
      const uint32_t b_base = smem_u32(smem + L::B_OFFSET);
      if (RESIDENT) { mbar_wait(&bfull[0], 0, 13); tc_fence_after_sync(); }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        if (!SPLIT) { mbar_wait(&tmem_empty[st.acc], st.phase ^ 1, 2); tc_fence_after_sync(); }
        const uint32_t d_tmem = tmem_base + (uint32_t)(st.acc * BN);
        for (int ph = 0; ph < n_phase; ++ph) {
          mbar_wait(&pfull[ps], pphase, 14);
          tc_fence_after_sync();
          const uint32_t stage = smem_u32(smem + ps * PATCH_STAGE_BYTES);
          if constexpr (FIXED3) {
            if (SPLIT && !(ph & 1))
              patch3_issue_phase<BN, RESIDENT, BS, TPS, L::B_BYTES, L::BSTAGE_BYTES, PW, 18, 2, SPLIT>(tmem_base, stage, b_base, 0u, true, bfull, bempty, bs, bphase, ph == 0, tmem_full, tmem_empty, st, args.debug);
            else
              patch3_issue_phase<BN, RESIDENT, BS, TPS, L::B_BYTES, L::BSTAGE_BYTES, PW, 9, 1, SPLIT>(tmem_base, stage, b_base, SPLIT ? 0u : (uint32_t)(ph * 9), false, bfull, bempty, bs, bphase, ph == 0, tmem_full, tmem_empty, st, args.debug);
          } else {
            for (int t = 0; t < args.n_taps; ++t) {
              uint32_t b_addr;
              if (RESIDENT) b_addr = b_base + (uint32_t)((ph * args.n_taps + t) * L::B_BYTES);
              else {
                if (t % TPS == 0) { mbar_wait(&bfull[bs], bphase, 15); tc_fence_after_sync(); }
                b_addr = b_base + (uint32_t)(bs * L::BSTAGE_BYTES + (t % TPS) * L::B_BYTES);
              }
              const int p = args.t_patch[t];
              const uint32_t a_view = stage + (uint32_t)args.p_base[p] + (uint32_t)args.t_off[t] * 128u;
              const uint64_t a_desc = umma_desc_sw128_sbo(a_view, (uint32_t)args.p_pw[p] * 128u);   // 8-row groups are one patch row apart
              const uint64_t b_desc = umma_desc_sw128(b_addr);
              if (!(args.debug & 2)) {
                if (elect_one()) umma_f16_x4(d_tmem, a_desc, b_desc, idesc, (ph | t) != 0 ? 1u : 0u);
              }
              if (!RESIDENT && (t % TPS) == TPS - 1) { if (elect_one()) umma_commit(&bempty[bs]); if (++bs == BS) { bs = 0; bphase ^= 1; } }
            }
          }
          if (elect_one()) umma_commit(&pempty[ps]);
          if (++ps == PS) { ps = 0; pphase ^= 1; }
        }
        if (!SPLIT) {
          if (elect_one()) umma_commit(&tmem_full[st.acc]);
          if (++st.acc == 2) { st.acc = 0; st.phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue (identical to conv_tc_kernel, BW=8 BH=16)
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;
    int acc = 0; uint32_t acc_phase = 0;
    const int rec = SPLIT ? 2 * args.Cout : args.Cout;
    for (int c = et; c < args.Cout; c += 128) { bias_s[c] = args.bias[c]; scale_s[c] = SPLIT ? args.scale[c] : 1.f; }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % args.tiles_n, mt = tile / args.tiles_n;
      const int tx = mt % args.tiles_x, ty = (mt / args.tiles_x) % args.tiles_y;
      const int n = mt / (args.tiles_x * args.tiles_y);
      const int yl = row / BW, xl = row - yl * BW;
      const size_t pix = ((size_t)n * args.Ho + (ty * BH + yl)) * args.Wo + (tx * BW + xl);
      const size_t off = pix * rec + (size_t)nt * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (!SPLIT) {
        conv_epilogue_tile<BN, false>(taddr, bias_s + nt * BN, scale_s + nt * BN, args.residual, args.out, off, args.Cout, !(args.debug & 1),
                                      args.relu, &tmem_full[acc], acc_phase);
        tc_fence_before_sync();
        mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      } else {
        float accr[BN];
        split_prefetch_residual<BN>(args.residual, off, args.Cout, !(args.debug & 1));
        const int n_chunks = 3 * args.n_cblk;          // per channel block: hi phase = 2 chunks of 9 tiles, lo phase = 1
        for (int c = 0; c < n_chunks; ++c) {
          split_chunk_add<BN>(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN), accr, c == 0, &tmem_full[acc], acc_phase);
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        split_store_tile<BN, BN, 0>(accr, bias_s + nt * BN, scale_s + nt * BN, args.residual, args.out, off, args.Cout, !(args.debug & 1), args.relu);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<2 * BN>(tmem_base); }
}

// ---------------------------------------------------------------- stem, two M-tiles per weight pass
// Timing experiments on the generic patch kernel (profiles/README.md, r01l) showed the 7x7/2 stem is bound by the
// round-trip latency of its weight ring, not by bytes or MMAs: 224 KB of weights (28 taps x 8 KB) cannot stay resident
// next to the patches, so every 128-pixel tile streams all of them through a 96 KB ring that covers only ~2300 cycles
// of tensor work. This variant makes one weight stage (the 4 pair-taps of one filter row) feed TWO vertically adjacent
// M-tiles (an 8 x 32 output block, two 64-column TMEM accumulators): weight traffic, barrier round trips and commits
// per output halve, and a ring stage covers twice the MMA time. To keep the patch ring double-buffered in the same
// shared-memory budget, a patch stage holds ONE row parity of the block (35 or 34 rows x 12 pixel-pair records), and
// the filter rows are visited parity-major: kh = 0,2,4,6 (odd input rows), then kh = 1,3,5 (even input rows).
//
// SPLIT mode (the 1e-4 path). The three products A_hi W_hi + A_hi W_lo + A_lo W_hi are ONE contraction over a tripled
// K, so the 18 input channels of a pixel become a 54 (-> 64) channel record [A_hi | A_hi | A_lo] against weights
// [W_hi | W_lo | W_hi]: one 128-byte record per PIXEL instead of per pixel pair, K = 49 taps x 64 = 3,136 per output
// (84 % useful) instead of 3 x 28 pair-taps x 64 = 5,376. Stride 2 in x is then handled like stride 2 in y: a patch stage
// holds one (row parity, column parity) plane of the block, so a block runs 4 phases -- (odd rows: kh = 0,2,4,6 | even
// rows: kh = 1,3,5) x (odd columns: kw = 0,2,4,6 | even columns: kw = 1,3,5) -- of 4 or 3 filter rows x 4 or 3 taps; a
// weight stage is still one filter row of one phase (4 tile slots, the 4th unused for the 3-tap phases).
constexpr int STEM2_PW = 12;                                         // patch pitch in 128-byte records
constexpr int STEM2_STAGE_BYTES = (35 * STEM2_PW * 128 + 1023) / 1024 * 1024;   // 54,272
constexpr int STEM2_PS = 2, STEM2_BS = 3, STEM2_TPS = 4;
struct Stem2Smem {
  static constexpr int B_BYTES = 64 * BLOCK_K * 2;                   // one tap: 8 KB
  static constexpr int BSTAGE_BYTES = STEM2_TPS * B_BYTES;           // one filter row: 32 KB
  static constexpr int B_OFFSET = STEM2_PS * STEM2_STAGE_BYTES;
  static constexpr int BAR_OFFSET = B_OFFSET + STEM2_BS * BSTAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 512;
  static constexpr int TOTAL = BIAS_OFFSET + 2 * 64 * 4 + 1024;
};
struct Stem2Args {
  int num_blocks, tiles_x, tiles_y;      // 8 x 32 output blocks
  int Ho, Wo;
  const float* bias;
  const float* scale;
  __half* out;
  int relu, debug;
};

// tensor maps: index = row parity (fast) / row parity * 2 + column parity (SPLIT); phase ph reads map NPH-1-ph
template <bool SPLIT>
__global__ void __launch_bounds__(256, 1)
stem2_kernel(const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmP1,
             const __grid_constant__ CUtensorMap tmP2, const __grid_constant__ CUtensorMap tmP3,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ Stem2Args args) {
  using L = Stem2Smem;
  constexpr int BN = 64, PS = STEM2_PS, BS = STEM2_BS;
  constexpr int NPH = SPLIT ? 4 : 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* pempty = pfull + PS;
  uint64_t* bfull = pempty + PS;
  uint64_t* bempty = bfull + BS;
  uint64_t* tmem_full = bempty + BS;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
  float* scale_s = bias_s + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP0); tma_prefetch_desc(&tmP1); tma_prefetch_desc(&tmP2); tma_prefetch_desc(&tmP3); tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PS; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<4 * BN>(tmem_base_slot);     // 2 accumulator sets x 2 tiles x 64 columns
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;
  // phase ph: row parity rp = 1 first (kh = 0,2,4,6 -> dy = -2..1, 35 patch rows from oy0-2), then rp = 0 (kh = 1,3,5 ->
  // dy = -1..1, 34 rows from oy0-1); SPLIT: inside each, column parity 1 (kw = 0,2,4,6 -> dx = -2..1, 4 taps) then
  // column parity 0 (kw = 1,3,5 -> dx = -1..1, 3 taps, views start one record further right); patches start at ox0-2.
  if (warp == 0) {
    // ===================================================== TMA producer
    int ps = 0, bs = 0; uint32_t pphase = 0, bphase = 0;
    for (int blk = blockIdx.x; blk < args.num_blocks; blk += gridDim.x) {
      const int tx = blk % args.tiles_x, ty = (blk / args.tiles_x) % args.tiles_y;
      const int n = blk / (args.tiles_x * args.tiles_y);
      const int ox0 = tx * 8, oy0 = ty * 32;
      int wrow = 0;
#pragma unroll
      for (int ph = 0; ph < NPH; ++ph) {
        const bool rp1 = SPLIT ? (ph < 2) : (ph == 0);
        const CUtensorMap* mp = (NPH - 1 - ph) == 0 ? &tmP0 : (NPH - 1 - ph) == 1 ? &tmP1 : (NPH - 1 - ph) == 2 ? &tmP2 : &tmP3;
        mbar_wait(&pempty[ps], pphase ^ 1, 21);
        uint8_t* stage = smem + ps * STEM2_STAGE_BYTES;
        if (args.debug & 4) { if (elect_one()) mbar_arrive(&pfull[ps]); }
        else {
          if (elect_one()) mbar_arrive_expect_tx(&pfull[ps], (rp1 ? 35 : 34) * STEM2_PW * 128);
          if (elect_one()) tma_load_4d(stage, mp, &pfull[ps], 0, ox0 - 2, oy0 + (rp1 ? -2 : -1), n);
        }
        if (++ps == PS) { ps = 0; pphase ^= 1; }
        const int nrows = rp1 ? 4 : 3;
        for (int i = 0; i < nrows; ++i, ++wrow) {
          // weight rows are stored in consumption order (fast: kh = 0,2,4,6,1,3,5 is NOT the storage order -> index by kh)
          const int row = SPLIT ? wrow : (2 * i + (rp1 ? 0 : 1));
          mbar_wait(&bempty[bs], bphase ^ 1, 22);
          if (args.debug & 4) { if (elect_one()) mbar_arrive(&bfull[bs]); }
          else {
            if (elect_one()) mbar_arrive_expect_tx(&bfull[bs], L::BSTAGE_BYTES);
            if (elect_one()) tma_load_3d(smem + L::B_OFFSET + bs * L::BSTAGE_BYTES, &tmB, &bfull[bs], 0, 0, row * 4);
          }
          if (++bs == BS) { bs = 0; bphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BN);
    int ps = 0, bs = 0; uint32_t pphase = 0, bphase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t b_base = smem_u32(smem + L::B_OFFSET);
    for (int blk = blockIdx.x; blk < args.num_blocks; blk += gridDim.x) {
      if (!SPLIT) { mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 23); tc_fence_after_sync(); }
#pragma unroll
      for (int ph = 0; ph < NPH; ++ph) {
        if (SPLIT) { mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 23); tc_fence_after_sync(); }   // SPLIT: one accumulation chunk per phase
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * BN);
        const bool rp1 = SPLIT ? (ph < 2) : (ph == 0);
        const int ntap = (SPLIT && (ph & 1)) ? 3 : 4;
        const int cshift = (SPLIT && (ph & 1)) ? 1 : 0;
        mbar_wait(&pfull[ps], pphase, 24);
        tc_fence_after_sync();
        const uint32_t stage = smem_u32(smem + ps * STEM2_STAGE_BYTES);
        const int nrows = rp1 ? 4 : 3;
        for (int i = 0; i < nrows; ++i) {            // patch row offset of filter row kh = 2i + (1 - rp) is i
          mbar_wait(&bfull[bs], bphase, 25);
          tc_fence_after_sync();
          const uint32_t b_stage = b_base + (uint32_t)(bs * L::BSTAGE_BYTES);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
              if (dp < ntap) {
                const uint32_t a_view = stage + (uint32_t)(((i + 16 * u) * STEM2_PW + dp + cshift) * 128);
                const uint64_t a_desc = umma_desc_sw128_sbo(a_view, STEM2_PW * 128u);
                const uint64_t b_desc = umma_desc_sw128(b_stage + (uint32_t)(dp * L::B_BYTES));
                if (!(args.debug & 2)) {
                  if (elect_one()) umma_f16_x4(d_tmem + (uint32_t)(u * BN), a_desc, b_desc, idesc, ((SPLIT ? 0 : ph) | i | dp) != 0 ? 1u : 0u);
                }
              }
            }
          }
          if (elect_one()) umma_commit(&bempty[bs]);
          if (++bs == BS) { bs = 0; bphase ^= 1; }
        }
        if (elect_one()) umma_commit(&pempty[ps]);
        if (++ps == PS) { ps = 0; pphase ^= 1; }
        if (SPLIT) {
          if (elect_one()) umma_commit(&tmem_full[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (!SPLIT) {
        if (elect_one()) umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue: two 128 x 64 accumulators per block
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;
    int acc = 0; uint32_t acc_phase = 0;
    if (et < BN) { bias_s[et] = args.bias[et]; scale_s[et] = SPLIT ? args.scale[et] : 1.f; }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int yl = row >> 3, xl = row & 7;
    constexpr int REC = SPLIT ? 2 * BN : BN;
    for (int blk = blockIdx.x; blk < args.num_blocks; blk += gridDim.x) {
      const int tx = blk % args.tiles_x, ty = (blk / args.tiles_x) % args.tiles_y;
      const int n = blk / (args.tiles_x * args.tiles_y);
      if constexpr (!SPLIT) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const size_t pix = ((size_t)n * args.Ho + (ty * 32 + u * 16 + yl)) * args.Wo + (tx * 8 + xl);
          const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * 2 * BN + u * BN);
          conv_epilogue_tile<BN, false>(taddr, bias_s, scale_s, nullptr, args.out, pix * REC, BN, !(args.debug & 1), args.relu, &tmem_full[acc], acc_phase);
        }
        tc_fence_before_sync();
        mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      } else {
        float accr[2 * BN];                 // both M-tiles of the block: columns [u * 64, u * 64 + 64) of the accumulator set
        for (int c = 0; c < NPH; ++c) {
          split_chunk_add<2 * BN>(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * 2 * BN), accr, c == 0, &tmem_full[acc], acc_phase);
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        const size_t pix0 = ((size_t)n * args.Ho + (ty * 32 + yl)) * args.Wo + (tx * 8 + xl);
        const size_t pix1 = pix0 + (size_t)16 * args.Wo;
        split_store_tile<BN, 2 * BN, 0>(accr, bias_s, scale_s, nullptr, args.out, pix0 * REC, BN, !(args.debug & 1), args.relu);
        split_store_tile<BN, 2 * BN, BN>(accr, bias_s, scale_s, nullptr, args.out, pix1 * REC, BN, !(args.debug & 1), args.relu);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<4 * BN>(tmem_base); }
}

// ---------------------------------------------------------------- elementwise helpers (fp16 NHWC)
// ARGMAX: additionally track, per image and heat-map channel 1..17, the first maximum above `eps` as a packed key
// (float bits << 32 | ~pixel index; atomicMax keeps the largest value and, on ties, the lowest index = torch.max's
// arg-max): the sample-ranking step (utils/label_conversions.py:127-155) then never re-reads the 1.1 GB of heat-maps.
// SPLIT: writes the stem's 64-channel split records instead (128 B per pixel, see stem2_kernel):
//   [A_hi c0..15 | A_hi c0..15 | A_lo c0..15 | A_hi c16,17 | A_hi c16,17 | A_lo c16,17 | 0 x 10]
__device__ __forceinline__ float ld_in(const float* p) { return *p; }
__device__ __forceinline__ float ld_in(const __half* p) { return __half2float(*p); }

// TIN = float (the reference's input type) or __half (opt-in: a host that already holds the proxy representation in fp16
// halves the bytes over PCIe; the values are then what they are -- their lo halves are zero)
template <bool ARGMAX, bool SPLIT, typename TIN = float>
__global__ void __launch_bounds__(256) nchw_f32_to_nhwc32_f16_kernel(const TIN* __restrict__ x, int C, int HW,
                                                                     __half* __restrict__ y, float eps,
                                                                     unsigned long long* __restrict__ keys, float in_scale) {
  // thread = (pixel, 16-channel half): 16 (or C-16) coalesced plane reads -> one full 32-byte sector of the
  // NHWC record (two 16-byte stores); channels >= C are written as zero padding.
  const int n = blockIdx.y;
  const int p = blockIdx.x * 128 + (threadIdx.x & 127);
  const int half_id = threadIdx.x >> 7;               // 0: channels 0..15, 1: channels 16..31 (warp-uniform)
  const bool valid = p < HW;
  const TIN* src = x + (size_t)n * C * HW + (valid ? p : 0);
  __half2 h[8];
  float v[16];
  unsigned candmask = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c0 = half_id * 16 + 2 * k;
    v[2 * k] = (valid && c0 < C) ? ld_in(src + (size_t)c0 * HW) : 0.f;
    v[2 * k + 1] = (valid && c0 + 1 < C) ? ld_in(src + (size_t)(c0 + 1) * HW) : 0.f;
    if (ARGMAX) {
      candmask |= (c0 >= 1 && v[2 * k] > eps) ? (1u << (2 * k)) : 0u;       // channels >= C were loaded as 0
      candmask |= (v[2 * k + 1] > eps) ? (2u << (2 * k)) : 0u;
    }
    h[k] = __floats2half2_rn(v[2 * k] * in_scale, v[2 * k + 1] * in_scale);
  }
  if (valid) {
    if constexpr (!SPLIT) {
      uint4* dst = reinterpret_cast<uint4*>(y + ((size_t)n * HW + p) * 32 + half_id * 16);
      dst[0] = *reinterpret_cast<const uint4*>(&h[0]);
      dst[1] = *reinterpret_cast<const uint4*>(&h[4]);
    } else {
      __half2 l[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 f = __half22float2(h[k]);
        l[k] = __floats2half2_rn(v[2 * k] * in_scale - f.x, v[2 * k + 1] * in_scale - f.y);
      }
      uint4* rec = reinterpret_cast<uint4*>(y + ((size_t)n * HW + p) * 64);
      if (half_id == 0) {
        rec[0] = *reinterpret_cast<const uint4*>(&h[0]); rec[1] = *reinterpret_cast<const uint4*>(&h[4]);
        rec[2] = *reinterpret_cast<const uint4*>(&h[0]); rec[3] = *reinterpret_cast<const uint4*>(&h[4]);
        rec[4] = *reinterpret_cast<const uint4*>(&l[0]); rec[5] = *reinterpret_cast<const uint4*>(&l[4]);
      } else {
        __half2 t[4] = {h[0], h[0], l[0], __float2half2_rn(0.f)};
        rec[6] = *reinterpret_cast<const uint4*>(&t[0]);
        rec[7] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  if (ARGMAX) {
    // heat-map values above eps are rare (one Gaussian blob per map): most warps leave after a single vote; the others
    // reduce (value, lowest pixel) over their lanes per candidate channel and post ONE global atomicMax each
    unsigned any = __reduce_or_sync(0xffffffffu, candmask);
    while (any) {
      const int e = __ffs(any) - 1;
      any &= any - 1;
      const bool cand = (candmask >> e) & 1u;
      unsigned hi = cand ? __float_as_uint(v[e]) : 0u, lo = cand ? (0xFFFFFFFFu - (unsigned)p) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned ohi = __shfl_xor_sync(0xffffffffu, hi, o), olo = __shfl_xor_sync(0xffffffffu, lo, o);
        if (ohi > hi || (ohi == hi && olo > lo)) { hi = ohi; lo = olo; }
      }
      const int c = half_id * 16 + e;
      if ((threadIdx.x & 31) == 0) atomicMax(&keys[(size_t)n * (C - 1) + (c - 1)], ((unsigned long long)hi << 32) | lo);
    }
  }
}

// fp32 NCHW proxy representation -> the split stem's 64-channel fp16 records, staged through shared memory so that every
// global store instruction writes 512 contiguous bytes (the first version wrote 16-byte pieces at a 128-byte stride from two
// of its eight warps: 0.93 ms at 47 % of the DRAM peak for 1.2 GB in + 2.1 GB out). CTA = 128 pixels x 256 threads:
// thread (pixel p, half hf) converts channel pairs {0..3, 8} (hf = 0) or {4..7} (hf = 1) -- pair k = channels 2k, 2k+1 --
// into the record words  hi -> k, dup -> 8 + k, lo -> 16 + k  (pair 8: 24 / 25 / 26; 27..31 = 0), the 16-byte chunks of a
// record XOR-swizzled by the pixel index against bank conflicts; then all 256 threads copy the 16 KB tile out.
template <bool ARGMAX, typename TIN = float>
__global__ void __launch_bounds__(256) nchw_f32_to_split_records_kernel(const TIN* __restrict__ x, int C, int HW,
                                                                        __half* __restrict__ y, float eps,
                                                                        unsigned long long* __restrict__ keys, float in_scale) {
  __shared__ __align__(16) uint32_t tile[128 * 32];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * 128;
  const int pl = threadIdx.x & 127, p = p0 + pl;
  const int hf = threadIdx.x >> 7;                    // warp-uniform
  const bool valid = p < HW;
  const TIN* src = x + (size_t)n * C * HW + (valid ? p : 0);
  const int npair = hf == 0 ? 5 : 4;
  float v[10];
  unsigned candmask = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if (i < npair) {
      const int k = hf == 0 ? (i < 4 ? i : 8) : 4 + i;          // channel pair index
      const int c0 = 2 * k;
      const float a = (valid && c0 < C) ? ld_in(src + (size_t)c0 * HW) : 0.f;
      const float b = (valid && c0 + 1 < C) ? ld_in(src + (size_t)(c0 + 1) * HW) : 0.f;
      v[2 * i] = a; v[2 * i + 1] = b;
      if (ARGMAX) {
        candmask |= (c0 >= 1 && a > eps) ? (1u << (2 * i)) : 0u;      // channel 0 is the edge map
        candmask |= (b > eps) ? (2u << (2 * i)) : 0u;
      }
      const __half2 h = __floats2half2_rn(a * in_scale, b * in_scale);
      const float2 f = __half22float2(h);
      const __half2 l = __floats2half2_rn(a * in_scale - f.x, b * in_scale - f.y);
      const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h), lw = *reinterpret_cast<const uint32_t*>(&l);
      const int w_hi = k < 8 ? k : 24, w_dup = k < 8 ? 8 + k : 25, w_lo = k < 8 ? 16 + k : 26;
      uint32_t* rec = tile + pl * 32;
      rec[(((w_hi >> 2) ^ (pl & 7)) << 2) | (w_hi & 3)] = hw;
      rec[(((w_dup >> 2) ^ (pl & 7)) << 2) | (w_dup & 3)] = hw;
      rec[(((w_lo >> 2) ^ (pl & 7)) << 2) | (w_lo & 3)] = lw;
    }
  }
  if (hf == 1) {                                       // zero padding: words 27..31
    uint32_t* rec = tile + pl * 32;
#pragma unroll
    for (int w = 27; w < 32; ++w) rec[(((w >> 2) ^ (pl & 7)) << 2) | (w & 3)] = 0u;
  }
  __syncthreads();
  {
    // 128 records x 8 chunks of 16 B = 1024 chunks, 4 per thread; consecutive threads -> consecutive 16-byte chunks in HBM
    uint4* dst = reinterpret_cast<uint4*>(y + ((size_t)n * HW + p0) * 64);
    const uint4* t4 = reinterpret_cast<const uint4*>(tile);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = i * 256 + threadIdx.x;             // global chunk index inside the tile
      const int r = g >> 3, ck = g & 7;
      if (p0 + r < HW) dst[g] = t4[r * 8 + (ck ^ (r & 7))];
    }
  }
  if (ARGMAX) {
    // heat-map values above eps are rare (one Gaussian blob per map): most warps leave after a single vote; the others
    // reduce (value, lowest pixel) over their lanes per candidate channel and post ONE global atomicMax each
    unsigned any = __reduce_or_sync(0xffffffffu, candmask);
    while (any) {
      const int e = __ffs(any) - 1;
      any &= any - 1;
      const bool cand = (candmask >> e) & 1u;
      float ve = 0.f;
#pragma unroll
      for (int q = 0; q < 10; ++q) if (q == e) ve = v[q];
      unsigned hi = cand ? __float_as_uint(ve) : 0u, lo = cand ? (0xFFFFFFFFu - (unsigned)p) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned ohi = __shfl_xor_sync(0xffffffffu, hi, o), olo = __shfl_xor_sync(0xffffffffu, lo, o);
        if (ohi > hi || (ohi == hi && olo > lo)) { hi = ohi; lo = olo; }
      }
      const int i = e >> 1;
      const int k = hf == 0 ? (i < 4 ? i : 8) : 4 + i;
      const int c = 2 * k + (e & 1);
      if ((threadIdx.x & 31) == 0) atomicMax(&keys[(size_t)n * (C - 1) + (c - 1)], ((unsigned long long)hi << 32) | lo);
    }
  }
}

// packed arg-max keys -> (x, y) pixel of the maximum (or -1, -1) and visibility, like utils/label_conversions.py:142-153
__global__ void argmax_decode_kernel(const unsigned long long* __restrict__ keys, int n, int W, float* __restrict__ j2d,
                                     int* __restrict__ vis) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  const bool v = k != 0ull;
  const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
  j2d[2 * i] = v ? (float)(idx % (unsigned)W) : -1.f;
  j2d[2 * i + 1] = v ? floorf((float)idx / (float)W) : -1.f;
  vis[i] = v ? 1 : 0;
}

__global__ void __launch_bounds__(256) maxpool3x3s2_f16_kernel(const __half* __restrict__ in, int H, int W, int C, int Ho,
                                                               int Wo, __half* __restrict__ out, size_t total8) {
  // one thread = 8 channels of one output pixel
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int c = (int)(i % c8) * 8;
  size_t r = i / c8;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int n = (int)(r / Ho);
  __half2 m[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) m[e] = __float2half2_rn(-65504.f);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= W) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + iy) * W + ix) * C + c);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], h[e]);
    }
  }
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) oh[e] = m[e];
  *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * C + c) = o;
}

__global__ void __launch_bounds__(256) avgpool_f16_kernel(const __half* __restrict__ in, int HW, int C,
                                                          float* __restrict__ out) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += __half2float(in[((size_t)n * HW + p) * C + c]);
    out[(size_t)n * C + c] = s / (float)HW;
  }
}

// ---------------------------------------------------------------- split-NHWC helpers ([..][2C] = [hi(C) | lo(C)])
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi4, uint4& lo4) {
  __half2* oh = reinterpret_cast<__half2*>(&hi4);
  __half2* ol = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    oh[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    const float2 f = __half22float2(oh[e]);
    ol[e] = __floats2half2_rn(v[2 * e] - f.x, v[2 * e + 1] - f.y);
  }
}
__device__ __forceinline__ void join8(const uint4& hi4, const uint4& lo4, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&hi4);
  const __half2* l = reinterpret_cast<const __half2*>(&lo4);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fh = __half22float2(h[e]), fl = __half22float2(l[e]);
    v[2 * e] = fh.x + fl.x; v[2 * e + 1] = fh.y + fl.y;
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_split_kernel(const __half* __restrict__ in, int H, int W, int C, int Ho,
                                                                 int Wo, __half* __restrict__ out, size_t total8) {
  // one thread = 8 channels (hi + lo) of one output pixel; the maximum is taken on the joined fp32 values
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int c = (int)(i % c8) * 8;
  size_t r = i / c8;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int n = (int)(r / Ho);
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= W) continue;
      const __half* rec = in + (((size_t)n * H + iy) * W + ix) * 2 * C + c;
      float v[8];
      join8(*reinterpret_cast<const uint4*>(rec), *reinterpret_cast<const uint4*>(rec + C), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    }
  }
  uint4 oh, ol;
  split8(m, oh, ol);
  __half* dst = out + (((size_t)n * Ho + oy) * Wo + ox) * 2 * C + c;
  *reinterpret_cast<uint4*>(dst) = oh;
  *reinterpret_cast<uint4*>(dst + C) = ol;
}

__global__ void __launch_bounds__(256) avgpool_split_kernel(const __half* __restrict__ in, int HW, int C,
                                                            float* __restrict__ out, float out_scale) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) {
      const __half* rec = in + ((size_t)n * HW + p) * 2 * C;
      s += __half2float(rec[c]) + __half2float(rec[C + c]);
    }
    out[(size_t)n * C + c] = s / (float)HW * out_scale;
  }
}

// debug taps: split-NHWC -> fp32 NHWC
__global__ void tap_copy_split_kernel(const __half* __restrict__ src, size_t npix, int C, float* __restrict__ dst, float out_scale) {
  const size_t total = npix * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t px = i / C; const int c = (int)(i - px * C);
    dst[i] = (__half2float(src[px * 2 * C + c]) + __half2float(src[px * 2 * C + C + c])) * out_scale;
  }
}

// ---------------------------------------------------------------- host-side plan
inline int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
// diagnostic (HP3D_SPLIT_ASHIFT=a): carry all split-mode activations scaled by 2^a (input cast x 2^a, biases x 2^a, features
// x 2^-a) -- moves the lo halves away from fp16's subnormal range; used once to tell operand flushing from accumulator rounding
inline float split_ascale() { return ldexpf(1.f, env_int("HP3D_SPLIT_ASHIFT", 0)); }

struct TcConv {
  __half* w = nullptr;        // conv_tc_kernel layout: [k-block][Cout][64] (fast) / [k-block][hi|lo][Cout][64] (split)
  __half* w_patch = nullptr;  // split only: patch-kernel consumption order (streamed 27 tiles / channel block, or resident 18)
  float* bias = nullptr;      // [Cout]
  float* scale = nullptr;     // [Cout] split only: 2^-s
  int cin, cout, k, stride, pad, ktot;
  bool stem = false;
  bool split = false;
  int bn;
};

struct EncoderTc {
  TcConv stem, conv[4][2][2], down[4];
  bool has_down[4] = {false, false, false, false};
  int num_sms = 148;
  bool split = false;
};

// 3-D tensor map over blocked weights {64, Cout, n_blocks}; one box = `blocks_per_box` blocks of `bn` channels.
int make_weight_tmap(const __half* w, int cout, int n_blocks, int bn, int blocks_per_box, CUtensorMap* out) {
  const uint64_t dims[3] = {64, (uint64_t)cout, (uint64_t)n_blocks};
  const uint64_t strides[2] = {128, (uint64_t)cout * 128};
  const uint32_t box[3] = {64, (uint32_t)bn, (uint32_t)blocks_per_box};
  return make_tmap_f16(out, w, 3, dims, strides, box);
}

// fp16 hi/lo parts of w * 2^s
inline void split_half(float w, __half& hi, __half& lo) {
  hi = __float2half_rn(w);
  lo = __float2half_rn(w - __half2float(hi));
}

int make_tc_conv(const hp3d_conv_bn& c, float eps, bool stem, bool split, TcConv& L) {
  std::vector<float> w_khwc, bias;
  // fold BN in fp64, layout [kh][kw][cin][cout]
  int rc = fold_conv_bn(c, eps, c.cin, w_khwc, bias);
  if (rc) return rc;
  L.cin = c.cin; L.cout = c.cout; L.k = c.k; L.stride = c.stride; L.pad = c.pad; L.stem = stem; L.split = split;
  L.bn = (c.cout == 64) ? 64 : 128;
  const __half hz = __float2half_rn(0.f);
  std::vector<float> inv_scale(c.cout, 1.f);
  if (split) {
    // per-output-channel power-of-two pre-scale: max|w| -> [256, 512)
    const int ntap = c.k * c.k;
    for (int o = 0; o < c.cout; ++o) {
      float m = 0.f;
      for (int t = 0; t < ntap; ++t) for (int i = 0; i < c.cin; ++i) m = std::max(m, fabsf(w_khwc[((size_t)t * c.cin + i) * c.cout + o]));
      int e = 0;
      if (m > 0.f) { frexpf(m, &e); }            // m = f * 2^e, f in [0.5, 1)
      const int sh = m > 0.f ? 9 - e : 0;        // m * 2^sh in [256, 512)
      const float sc = ldexpf(1.f, sh);
      inv_scale[o] = ldexpf(1.f, -sh);
      for (int t = 0; t < ntap; ++t) for (int i = 0; i < c.cin; ++i) w_khwc[((size_t)t * c.cin + i) * c.cout + o] *= sc;
    }
  }
  auto W = [&](int t, int i, int o) { return w_khwc[((size_t)t * c.cin + i) * c.cout + o]; };
  std::vector<__half> wk, wp;
  if (!stem) {
    L.ktot = c.k * c.k * c.cin;
    const int ntaps = c.k * c.k, ncb = c.cin / 64;
    if (!split) {
      // blocked layout [kbm][Cout][64]: one 128-byte row per (k-block, output channel); k-blocks are ordered
      // channel-block major, tap minor (kbm = cb * ntaps + tap) so the taps of one channel block are contiguous and a
      // group of them is a single 3-D TMA box.
      wk.assign((size_t)c.cout * L.ktot, hz);
      for (int cb = 0; cb < ncb; ++cb)
        for (int t = 0; t < ntaps; ++t)
          for (int o = 0; o < c.cout; ++o)
            for (int i = 0; i < 64; ++i)
              wk[(((size_t)cb * ntaps + t) * c.cout + o) * 64 + i] = __float2half_rn(W(t, cb * 64 + i, o));
    } else {
      // conv_tc_kernel: [kbm][part][Cout][64] -- the hi and lo planes of a k-block are adjacent: one {64, BN, 2} box
      wk.assign((size_t)2 * c.cout * L.ktot, hz);
      for (int cb = 0; cb < ncb; ++cb)
        for (int t = 0; t < ntaps; ++t)
          for (int o = 0; o < c.cout; ++o)
            for (int i = 0; i < 64; ++i) {
              __half hi, lo;
              split_half(W(t, cb * 64 + i, o), hi, lo);
              const size_t kbm = (size_t)cb * ntaps + t;
              wk[((kbm * 2 + 0) * c.cout + o) * 64 + i] = hi;
              wk[((kbm * 2 + 1) * c.cout + o) * 64 + i] = lo;
            }
      if (c.k == 3 && c.stride == 1 && c.pad == 1) {
        const bool resident = (L.bn == 64 && ncb == 1);
        const int nblk = resident ? 18 : 27 * ncb;
        wp.assign((size_t)nblk * c.cout * 64, hz);
        for (int cb = 0; cb < ncb; ++cb)
          for (int t = 0; t < 9; ++t)
            for (int o = 0; o < c.cout; ++o)
              for (int i = 0; i < 64; ++i) {
                __half hi, lo;
                split_half(W(t, cb * 64 + i, o), hi, lo);
                if (resident) {
                  wp[((size_t)(t)*c.cout + o) * 64 + i] = hi;
                  wp[((size_t)(9 + t) * c.cout + o) * 64 + i] = lo;
                } else {
                  wp[((size_t)(cb * 27 + 2 * t) * c.cout + o) * 64 + i] = hi;        // hi phase: t.hi, t.lo interleaved
                  wp[((size_t)(cb * 27 + 2 * t + 1) * c.cout + o) * 64 + i] = lo;
                  wp[((size_t)(cb * 27 + 18 + t) * c.cout + o) * 64 + i] = hi;       // lo phase: W_hi again
                }
              }
      }
    }
  } else if (!split) {
    // k-block = kh*4 + (dp+2); inside it e*32 + ci with input column 2(x+dp)+e = 2x + kw - 3  =>  kw = 2dp + e + 3
    L.ktot = 7 * 256;
    wk.assign((size_t)c.cout * L.ktot, hz);
    for (int o = 0; o < c.cout; ++o)
      for (int kh = 0; kh < 7; ++kh)
        for (int dp = -2; dp <= 1; ++dp)
          for (int e = 0; e < 2; ++e) {
            const int kw = 2 * dp + e + 3;
            if (kw < 0 || kw > 6) continue;
            for (int i = 0; i < c.cin; ++i)
              wk[(((size_t)kh * 4 + (dp + 2)) * c.cout + o) * 64 + e * 32 + i] = __float2half_rn(W(kh * 7 + kw, i, o));
          }
  } else {
    // split stem (stem2_kernel<true>): 14 weight rows in consumption order, 4 tile slots each. Row order: phases
    // (rp,cp) = (1,1), (1,0), (0,1), (0,0); phase rows i -> kh = 2i (rp=1) / 2i+1 (rp=0); slot dp -> kw = 2(dp-2)+4 = 2dp
    // (cp=1) / 2(dp-1)+3 = 2dp+1 (cp=0, 3 slots used). Inside a tile the k index follows the input record
    // [A_hi c0..15 | A_hi c0..15 | A_lo c0..15 | A_hi c16,17 | A_hi c16,17 | A_lo c16,17 | 0]:
    //  weights [W_hi | W_lo | W_hi | W_hi | W_lo | W_hi | 0].
    if (c.cin != 18) { set_error("split stem expects 18 input channels"); return -1; }
    L.ktot = 14 * 256;
    wk.assign((size_t)c.cout * L.ktot, hz);
    int row = 0;
    for (int ph = 0; ph < 4; ++ph) {
      const int rp = ph < 2 ? 1 : 0, cp = (ph & 1) ? 0 : 1;
      const int nrows = rp ? 4 : 3, ntap = cp ? 4 : 3;
      for (int i = 0; i < nrows; ++i, ++row) {
        const int kh = rp ? 2 * i : 2 * i + 1;
        for (int dp = 0; dp < ntap; ++dp) {
          const int kw = cp ? 2 * dp : 2 * dp + 1;
          for (int o = 0; o < c.cout; ++o) {
            __half* blk = &wk[(((size_t)row * 4 + dp) * c.cout + o) * 64];
            for (int i2 = 0; i2 < 18; ++i2) {
              __half hi, lo;
              split_half(W(kh * 7 + kw, i2, o), hi, lo);
              if (i2 < 16) { blk[i2] = hi; blk[16 + i2] = lo; blk[32 + i2] = hi; }
              else { blk[48 + (i2 - 16)] = hi; blk[50 + (i2 - 16)] = lo; blk[52 + (i2 - 16)] = hi; }
            }
          }
        }
      }
    }
  }
  rc = upload(&L.w, wk.data(), wk.size());
  if (!rc && !wp.empty()) rc = upload(&L.w_patch, wp.data(), wp.size());
  if (split) { const float as = split_ascale(); for (auto& b : bias) b *= as; }
  rc = rc ? rc : upload(&L.bias, bias.data(), bias.size());
  if (!rc && split) rc = upload(&L.scale, inv_scale.data(), inv_scale.size());
  return rc;
}

// Launch one convolution: in [N][H][W][Cin_mem] fp16 (fast: Cin_mem = Cin, 32 for the stem input; split: 2 Cin, 64 for
// the stem input), out [N][Ho][Wo][Cout] (split: 2 Cout).
int run_patch_conv(const EncoderTc* E, const TcConv& L, const __half* in, int N, int H, int W, const __half* residual,
                   int relu, __half* out, cudaStream_t s);

template <int BN, int STAGES, bool SPLIT>
int launch_conv_tc(const CUtensorMap* tmA, const CUtensorMap& tmB, const ConvTcArgs& a, int grid, cudaStream_t s) {
  using SL = SmemLayout<BN, STAGES, SPLIT>;
  static_assert(SL::TOTAL <= 232448, "shared memory budget");
  HP3D_SMEM_OPT_IN((conv_tc_kernel<BN, STAGES, SPLIT>), SL::TOTAL);
  conv_tc_kernel<BN, STAGES, SPLIT><<<grid, 256, SL::TOTAL, s>>>(tmA[0], tmA[1], tmA[2], tmA[3], tmB, a);
  return launch_status("conv_tc_kernel");
}

int run_tc_conv(const EncoderTc* E, const TcConv& L, const __half* in, int N, int H, int W, const __half* residual,
                int relu, __half* out, cudaStream_t s) {
  {
    const int prc = run_patch_conv(E, L, in, N, H, W, residual, relu, out, s);
    if (prc != -1000) return prc;  // handled (0) or failed; -1000 = not covered by the patch variant
  }
  if (L.stem && L.split) { set_error("conv_tc: the split stem needs stem2_kernel (output height must be a multiple of 32)"); return -1; }
  const int Ho = H / L.stride, Wo = W / L.stride;
  const uint64_t Np = (uint64_t)((N + 1) & ~1);   // buffers hold an even number of images (layer4 tiles span two)
  const uint64_t rec = (uint64_t)(L.split ? 2 * L.cin : L.cin);   // fp16 elements per input pixel record
  ConvTcArgs a;
  memset(&a, 0, sizeof(a));
  if (Wo >= 16 && Ho >= 8 && Wo % 16 == 0 && Ho % 8 == 0) { a.BW = 16; a.BH = 8; a.BIMG = 1; }
  else if (Wo == 8 && Ho == 8) { a.BW = 8; a.BH = 8; a.BIMG = 2; }
  else { set_error("conv_tc: unsupported output size %dx%d", Ho, Wo); return -1; }
  a.tiles_x = Wo / a.BW; a.tiles_y = Ho / a.BH;
  a.tiles_m = a.tiles_x * a.tiles_y * cdiv(N, a.BIMG);
  a.tiles_n = L.cout / L.bn;
  a.N = N; a.Ho = Ho; a.Wo = Wo; a.Cout = L.cout; a.Cin = L.cin;
  a.bias = L.bias; a.scale = L.scale; a.residual = residual; a.out = out; a.relu = relu;
  CUtensorMap tmA[4];
  const uint32_t box[4] = {64, (uint32_t)a.BW, (uint32_t)a.BH, (uint32_t)a.BIMG};
  int nmaps = 0;
  int nk = 0;
  if (L.stem) {
    // pixel-pair view of [N][H][W][32]: {64, W/2, H/2 (row parity ph), N}
    for (int ph = 0; ph < 2; ++ph) {
      const uint64_t dims[4] = {64, (uint64_t)W / 2, (uint64_t)H / 2, Np};
      const uint64_t st[3] = {128, (uint64_t)2 * W * 64, (uint64_t)H * W * 64};
      int rc = make_tmap_f16(&tmA[ph], in + (size_t)ph * W * 32, 4, dims, st, box);
      if (rc) return rc;
    }
    nmaps = 2;
    for (int kh = 0; kh < 7; ++kh) {
      const int o = kh - 3;
      const int ph = ((o % 2) + 2) % 2;
      const int dy = (o - ph) / 2;
      for (int dp = -2; dp <= 1; ++dp) {
        KBlock& k = a.kb[nk++];
        k.map = (int8_t)ph; k.dx = (int8_t)dp; k.dy = (int8_t)dy; k.c = 0; k.bk = (int16_t)(kh * 4 + (dp + 2));
      }
    }
  } else if (L.stride == 1) {
    const uint64_t dims[4] = {rec, (uint64_t)W, (uint64_t)H, Np};
    const uint64_t st[3] = {rec * 2, (uint64_t)W * rec * 2, (uint64_t)H * W * rec * 2};
    int rc = make_tmap_f16(&tmA[0], in, 4, dims, st, box);
    if (rc) return rc;
    nmaps = 1;
    for (int kh = 0; kh < L.k; ++kh)
      for (int kw = 0; kw < L.k; ++kw)
        for (int c = 0; c < L.cin; c += 64) {
          KBlock& k = a.kb[nk++];
          k.map = 0; k.dx = (int8_t)(kw - L.pad); k.dy = (int8_t)(kh - L.pad); k.c = (int16_t)c;
          k.bk = (int16_t)((c / 64) * (L.k * L.k) + kh * L.k + kw);
        }
  } else {   // stride 2: parity maps, map index = py*2 + px
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const uint64_t dims[4] = {rec, (uint64_t)W / 2, (uint64_t)H / 2, Np};
        const uint64_t st[3] = {(uint64_t)2 * rec * 2, (uint64_t)2 * W * rec * 2, (uint64_t)H * W * rec * 2};
        int rc = make_tmap_f16(&tmA[py * 2 + px], in + ((size_t)py * W + px) * rec, 4, dims, st, box);
        if (rc) return rc;
      }
    nmaps = 4;
    for (int kh = 0; kh < L.k; ++kh)
      for (int kw = 0; kw < L.k; ++kw) {
        const int oy = kh - L.pad, ox = kw - L.pad;
        const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
        for (int c = 0; c < L.cin; c += 64) {
          KBlock& k = a.kb[nk++];
          k.map = (int8_t)(py * 2 + px); k.dx = (int8_t)((ox - px) / 2); k.dy = (int8_t)((oy - py) / 2); k.c = (int16_t)c;
          k.bk = (int16_t)((c / 64) * (L.k * L.k) + kh * L.k + kw);
        }
      }
  }
  for (int i = nmaps; i < 4; ++i) tmA[i] = tmA[0];
  if (nk > MAX_KB || nk * 64 != L.ktot) { set_error("conv_tc: k-block table mismatch (%d blocks, K=%d)", nk, L.ktot); return -1; }
  a.num_kb = nk;
  const int grid = std::min(a.tiles_m * a.tiles_n, E->num_sms);
  CUtensorMap tmB;
  int rc = L.split ? make_weight_tmap(L.w, L.cout, 2 * nk, L.bn, 2, &tmB) : make_weight_tmap(L.w, L.cout, nk, L.bn, 1, &tmB);
  if (rc) return rc;
  if (L.split) return L.bn == 64 ? launch_conv_tc<64, 4, true>(tmA, tmB, a, grid, s) : launch_conv_tc<128, 3, true>(tmA, tmB, a, grid, s);
  return L.bn == 64 ? launch_conv_tc<64, 6, false>(tmA, tmB, a, grid, s) : launch_conv_tc<128, 5, false>(tmA, tmB, a, grid, s);
}

// Patch-variant launch. Returns PATCH_NOT_COVERED if this layer/geometry is not handled (caller falls back to conv_tc_kernel).
constexpr int PATCH_NOT_COVERED = -1000;
template <int BN, bool RESIDENT, int PS, int BS, int PSB, int NKB, int TPS, bool FIXED3, bool SPLIT, int PW>
int launch_patch(const CUtensorMap& p0, const CUtensorMap& p1, const CUtensorMap& b, const PatchArgs& a, int grid, cudaStream_t s) {
  using SL = PatchSmem<BN, RESIDENT, PS, BS, PSB, NKB, TPS>;
  static_assert(SL::TOTAL <= 232448, "shared memory budget");
  HP3D_SMEM_OPT_IN((conv_patch_kernel<BN, RESIDENT, PS, BS, PSB, NKB, TPS, FIXED3, SPLIT, PW>), SL::TOTAL);
  conv_patch_kernel<BN, RESIDENT, PS, BS, PSB, NKB, TPS, FIXED3, SPLIT, PW><<<grid, 256, SL::TOTAL, s>>>(p0, p1, b, a);
  return launch_status("conv_patch_kernel");
}

// 3x3/1 halo patch of an 8x16 output tile, 64 channels: 18 rows x PW pixel records. PW = 16 (36 KB) is the round-1 pitch;
// PW = 10 (8 outputs + 2 halo columns, 22.5 KB) is what the split mode needs so that layer 1's 144 KB of resident hi + lo
// weights leave room for a 3-deep patch ring (the UMMA view and the TMA box are both plain sequences of 128-byte records,
// the 128B swizzle is a function of the shared-memory address bits, so any pitch works -- the stem has always used 12).
constexpr int patch3_bytes(int pw) { return 18 * pw * 128; }
constexpr int patch3_stage(int pw) { return (patch3_bytes(pw) + 1023) / 1024 * 1024; }
constexpr int STEM_PW = 12;                                           // stem patch pitch: 8 outputs + 3 halo pairs, padded to 12
constexpr int STEM_PATCH_TX = (18 + 19) * STEM_PW * 128;              // two row-parity patches of pixel pairs: 55.5 KB
constexpr int STEM_PATCH_BYTES = (STEM_PATCH_TX + 1023) / 1024 * 1024;  // ring stage pitch (1024-byte aligned)


int run_patch_conv(const EncoderTc* E, const TcConv& L, const __half* in, int N, int H, int W, const __half* residual,
                   int relu, __half* out, cudaStream_t s) {
  if (env_int("HP3D_CONV_PATCH", 1) == 0 && !(L.stem && L.split)) return PATCH_NOT_COVERED;
  const int Ho = H / L.stride, Wo = W / L.stride;
  if (Wo % 8 || Ho % 16) return PATCH_NOT_COVERED;
  const bool generic = !L.stem && L.stride == 1 && L.k == 3 && L.pad == 1 && L.cin % 64 == 0;
  if (!generic && !L.stem) return PATCH_NOT_COVERED;
  const uint64_t Np = (uint64_t)((N + 1) & ~1);
  const int debug = env_int("HP3D_CONV_DEBUG", 0);
  if (L.stem && (L.split || env_int("HP3D_STEM2", 1)) && Ho % 32 == 0 && L.cout == 64) {
    // two M-tiles per weight pass (stem2_kernel): one row-parity (split: row x column parity) patch of an 8 x 32 output
    // block per stage
    Stem2Args sa;
    memset(&sa, 0, sizeof(sa));
    sa.tiles_x = Wo / 8; sa.tiles_y = Ho / 32; sa.num_blocks = sa.tiles_x * sa.tiles_y * N;
    sa.Ho = Ho; sa.Wo = Wo; sa.bias = L.bias; sa.scale = L.scale; sa.out = out; sa.relu = relu; sa.debug = debug;
    CUtensorMap tp[4], tmBg;
    const int grid = std::min(sa.num_blocks, E->num_sms);
    if (!L.split) {
      for (int ph = 0; ph < 2; ++ph) {   // pixel-pair view of [N][H][W][32]
        const uint64_t dims[4] = {64, (uint64_t)W / 2, (uint64_t)H / 2, Np};
        const uint64_t st[3] = {128, (uint64_t)2 * W * 64, (uint64_t)H * W * 64};
        const uint32_t box[4] = {64, STEM2_PW, (uint32_t)(ph == 0 ? 34 : 35), 1};
        int rc = make_tmap_f16(&tp[ph], in + (size_t)ph * W * 32, 4, dims, st, box, true);
        if (rc) return rc;
      }
      tp[2] = tp[0]; tp[3] = tp[1];
      int rc = make_weight_tmap(L.w, L.cout, 28, 64, STEM2_TPS, &tmBg); if (rc) return rc;
      static_assert(Stem2Smem::TOTAL <= 232448, "shared memory budget");
      HP3D_SMEM_OPT_IN(stem2_kernel<false>, Stem2Smem::TOTAL);
      stem2_kernel<false><<<grid, 256, Stem2Smem::TOTAL, s>>>(tp[0], tp[1], tp[2], tp[3], tmBg, sa);
    } else {
      for (int rp = 0; rp < 2; ++rp)
        for (int cp = 0; cp < 2; ++cp) {   // (row parity, column parity) plane of [N][H][W][64]
          const uint64_t dims[4] = {64, (uint64_t)W / 2, (uint64_t)H / 2, Np};
          const uint64_t st[3] = {256, (uint64_t)2 * W * 128, (uint64_t)H * W * 128};
          const uint32_t box[4] = {64, STEM2_PW, (uint32_t)(rp == 0 ? 34 : 35), 1};
          int rc = make_tmap_f16(&tp[rp * 2 + cp], in + ((size_t)rp * W + cp) * 64, 4, dims, st, box, true);
          if (rc) return rc;
        }
      int rc = make_weight_tmap(L.w, L.cout, 56, 64, STEM2_TPS, &tmBg); if (rc) return rc;
      HP3D_SMEM_OPT_IN(stem2_kernel<true>, Stem2Smem::TOTAL);
      stem2_kernel<true><<<grid, 256, Stem2Smem::TOTAL, s>>>(tp[0], tp[1], tp[2], tp[3], tmBg, sa);
    }
    return launch_status("stem2_kernel");
  }
  if (L.stem && L.split) { set_error("split stem: unsupported geometry %dx%d", Ho, Wo); return -1; }
  PatchArgs a;
  memset(&a, 0, sizeof(a));
  a.tiles_x = Wo / 8; a.tiles_y = Ho / 16; a.tiles_m = a.tiles_x * a.tiles_y * N; a.tiles_n = L.cout / L.bn;
  a.N = N; a.Ho = Ho; a.Wo = Wo; a.Cout = L.cout; a.Cin = L.cin;
  a.bias = L.bias; a.scale = L.scale; a.residual = residual; a.out = out; a.relu = relu;
  a.debug = debug;
  CUtensorMap tmP[2];
  const int pw = L.split ? env_int("HP3D_PATCH_PW", 10) : 16;
  if (generic) {
    const uint64_t rec = (uint64_t)(L.split ? 2 * L.cin : L.cin);
    a.n_cblk = L.cin / 64; a.n_taps = 9; a.n_patch = 1; a.patch_tx = patch3_bytes(pw);
    a.p_pw[0] = pw; a.p_ph[0] = 18; a.p_ox[0] = -1; a.p_oy[0] = -1; a.p_base[0] = 0;
    for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) { a.t_off[kh * 3 + kw] = (uint16_t)(kh * pw + kw); a.t_patch[kh * 3 + kw] = 0; }
    const uint64_t dims[4] = {rec, (uint64_t)W, (uint64_t)H, Np};
    const uint64_t st[3] = {rec * 2, (uint64_t)W * rec * 2, (uint64_t)H * W * rec * 2};
    const uint32_t box[4] = {64, (uint32_t)pw, 18, 1};
    int rc = make_tmap_f16(&tmP[0], in, 4, dims, st, box, true);
    if (rc) return rc;
    tmP[1] = tmP[0];
  } else {   // stem: pixel pairs, two row-parity patches; tap index == weight k-block index = kh*4 + (dp+2)
    a.n_cblk = 1; a.n_taps = 28; a.n_patch = 2; a.patch_tx = STEM_PATCH_TX;
    // parity 0 rows: kh = 1,3,5 -> dy = -1,0,1 (18 rows); parity 1 rows: kh = 0,2,4,6 -> dy = -2..1 (19 rows)
    a.p_pw[0] = STEM_PW; a.p_ph[0] = 18; a.p_ox[0] = -2; a.p_oy[0] = -1; a.p_base[0] = 0;
    a.p_pw[1] = STEM_PW; a.p_ph[1] = 19; a.p_ox[1] = -2; a.p_oy[1] = -2; a.p_base[1] = 18 * STEM_PW * 128;
    for (int kh = 0; kh < 7; ++kh) {
      const int o = kh - 3, ph = ((o % 2) + 2) % 2, dy = (o - ph) / 2;
      for (int dp = -2; dp <= 1; ++dp) {
        const int t = kh * 4 + (dp + 2);
        a.t_patch[t] = (uint8_t)ph;
        a.t_off[t] = (uint16_t)((dy - a.p_oy[ph]) * STEM_PW + (dp + 2));
      }
    }
    for (int ph = 0; ph < 2; ++ph) {
      const uint64_t dims[4] = {64, (uint64_t)W / 2, (uint64_t)H / 2, Np};
      const uint64_t st[3] = {128, (uint64_t)2 * W * 64, (uint64_t)H * W * 64};
      const uint32_t box[4] = {64, STEM_PW, (uint32_t)a.p_ph[ph], 1};
      int rc = make_tmap_f16(&tmP[ph], in + (size_t)ph * W * 32, 4, dims, st, box, true);
      if (rc) return rc;
    }
  }
  const int grid = std::min(a.tiles_m * a.tiles_n, E->num_sms);
  CUtensorMap tmBg;     // weights grouped: TPS blocks per box
  if (L.stem) {
    int rc = make_weight_tmap(L.w, L.cout, 28, 64, 4, &tmBg); if (rc) return rc;
    return launch_patch<64, false, 2, 3, STEM_PATCH_BYTES, 1, 4, false, false, 16>(tmP[0], tmP[1], tmBg, a, grid, s);
  }
  if (!L.split) {
    if (L.bn == 64 && a.n_cblk == 1) {
      int rc = make_weight_tmap(L.w, L.cout, 9, 64, 9, &tmBg); if (rc) return rc;
      return launch_patch<64, true, 3, 1, patch3_stage(16), 9, 1, true, false, 16>(tmP[0], tmP[1], tmBg, a, grid, s);
    }
    if (L.bn == 128) {
      int rc = make_weight_tmap(L.w, L.cout, 9 * a.n_cblk, 128, 3, &tmBg); if (rc) return rc;
      return launch_patch<128, false, 3, 2, patch3_stage(16), 1, 3, true, false, 16>(tmP[0], tmP[1], tmBg, a, grid, s);
    }
    return PATCH_NOT_COVERED;
  }
  if (!L.w_patch) return PATCH_NOT_COVERED;
  if (L.bn == 64 && a.n_cblk == 1) {      // layer 1: 18 resident tiles (144 KB) + patch ring
    int rc = make_weight_tmap(L.w_patch, L.cout, 18, 64, 18, &tmBg); if (rc) return rc;
    if (pw == 10) return launch_patch<64, true, 3, 1, patch3_stage(10), 18, 1, true, true, 10>(tmP[0], tmP[1], tmBg, a, grid, s);
    return launch_patch<64, true, 2, 1, patch3_stage(16), 18, 1, true, true, 16>(tmP[0], tmP[1], tmBg, a, grid, s);
  }
  if (L.bn == 128) {
    int rc = make_weight_tmap(L.w_patch, L.cout, 27 * a.n_cblk, 128, 3, &tmBg); if (rc) return rc;
    if (pw == 10) return launch_patch<128, false, 4, 2, patch3_stage(10), 1, 3, true, true, 10>(tmP[0], tmP[1], tmBg, a, grid, s);
    return launch_patch<128, false, 3, 2, patch3_stage(16), 1, 3, true, true, 16>(tmP[0], tmP[1], tmBg, a, grid, s);
  }
  return PATCH_NOT_COVERED;
}

size_t act_bytes(int N, int H, int W, int C) { return align_up((size_t)N * H * W * C * 2, 1024); }

int tap_copy_act(const EncoderTc* E, const __half* src, size_t npix, int C, float** taps, cudaStream_t s) {
  if (!E->split) return tap_copy_f16(src, npix * C, taps, s);
  if (!*taps) return 0;
  tap_copy_split_kernel<<<1184, 256, 0, s>>>(src, npix, C, *taps, 1.f / split_ascale());
  *taps += npix * C;
  return launch_status("tap_copy_split_kernel");
}

}  // namespace

namespace hp3d {

int encoder_tc_create(const hp3d_encoder_weights* w, bool split, void** out) {
  if (!encode_fn()) { set_error("the tensor-core encoder needs cuTensorMapEncodeTiled (driver >= 12.0)"); return -3; }
  EncoderTc* E = new EncoderTc();
  E->split = split;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&E->num_sms, cudaDevAttrMultiProcessorCount, dev);
  E->num_sms = persistent_ctas(E->num_sms);
  int rc = make_tc_conv(w->stem, w->bn_eps, true, split, E->stem);
  const int planes[4] = {64, 128, 256, 512};
  int inpl = 64;
  for (int l = 0; l < 4 && !rc; ++l)
    for (int b = 0; b < 2 && !rc; ++b) {
      const hp3d_conv_bn& c1 = w->conv[l][b][0];
      const hp3d_conv_bn& c2 = w->conv[l][b][1];
      const int stride = (l > 0 && b == 0) ? 2 : 1;
      if (c1.cin != inpl || c1.cout != planes[l] || c1.k != 3 || c1.stride != stride || c1.pad != 1 || c2.cin != planes[l] ||
          c2.cout != planes[l] || c2.k != 3 || c2.stride != 1 || c2.pad != 1) {
        set_error("encoder_tc_create: layer%d.%d is not a ResNet-18 BasicBlock", l + 1, b); rc = -1; break;
      }
      rc = make_tc_conv(c1, w->bn_eps, false, split, E->conv[l][b][0]);
      rc = rc ? rc : make_tc_conv(c2, w->bn_eps, false, split, E->conv[l][b][1]);
      if (b == 0 && l > 0 && !rc) {
        const hp3d_conv_bn& d = w->down[l];
        if (!d.w || d.cin != inpl || d.cout != planes[l] || d.k != 1 || d.stride != 2 || d.pad != 0) {
          set_error("encoder_tc_create: layer%d downsample must be 1x1/2", l + 1); rc = -1; break;
        }
        rc = make_tc_conv(d, w->bn_eps, false, split, E->down[l]);
        E->has_down[l] = true;
      }
      inpl = planes[l];
    }
  if (rc) { encoder_tc_destroy(E); return rc; }
  *out = E;
  return 0;
}

void encoder_tc_destroy(void* p) {
  if (!p) return;
  EncoderTc* E = (EncoderTc*)p;
  auto fr = [](TcConv& L) { cudaFree(L.w); cudaFree(L.w_patch); cudaFree(L.bias); cudaFree(L.scale); };
  fr(E->stem);
  for (int l = 0; l < 4; ++l) { for (int b = 0; b < 2; ++b) { fr(E->conv[l][b][0]); fr(E->conv[l][b][1]); } fr(E->down[l]); }
  delete E;
}

size_t encoder_tc_workspace_bytes(const void* p, int B, int H, int W) {
  const int Bp = (B + 1) & ~1;   // layer4 tiles span two images
  const int m = ((const EncoderTc*)p)->split ? 2 : 1;   // split-NHWC records are twice as wide
  return act_bytes(Bp, H, W, 32 * m) + act_bytes(Bp, H / 2, W / 2, 64 * m) + 4 * act_bytes(Bp, H / 4, W / 4, 64 * m) +
         align_up((size_t)B * 17 * 8, 1024);    // heat-map arg-max keys (encoder_tc_forward with an ArgmaxOut)
}

int encoder_tc_forward(const void* p, const float* x, int B, int H, int W, float* feats, void* workspace,
                       size_t workspace_bytes, float* taps, cudaStream_t s, const ImageInput* image, const ArgmaxOut* amax, bool x_half) {
  const EncoderTc* E = (const EncoderTc*)p;
  if (H != 256 || W != 256) { set_error("the tensor-core encoder supports 256x256 proxy representations (DATA.PROXY_REP_SIZE)"); return -1; }
  const int Bp = (B + 1) & ~1;
  const int m = E->split ? 2 : 1;
  char* ws = (char*)workspace;
  __half* xin = (__half*)ws; ws += act_bytes(Bp, H, W, 32 * m);
  __half* stem = (__half*)ws; ws += act_bytes(Bp, H / 2, W / 2, 64 * m);
  __half* buf[4];
  for (int i = 0; i < 4; ++i) { buf[i] = (__half*)ws; ws += act_bytes(Bp, H / 4, W / 4, 64 * m); }
  int rc;
  const dim3 cgrid(cdiv(H * W, 128), B);
  const float in_scale = E->split ? split_ascale() : 1.f;
  if (image) {    // Canny edges + joint heat-maps written straight into the stem's fp16 input records (proxy.cu)
    rc = proxy_rep_nhwc_f16(image->rgb, image->joints2d, image->visibility, B, H, image->gaussian_std, image->gaussian_size,
                            image->threshold, image->nms, image->heat_std, xin, E->split ? 1 : 0, s);
  } else {
    // input cast (+ heat-map arg-max as a by-product); x is fp32 NCHW, or fp16 NCHW when x_half
    unsigned long long* keys = nullptr;
    const float eps = amax ? amax->eps : 0.f;
    if (amax) {
      keys = (unsigned long long*)((char*)workspace + encoder_tc_workspace_bytes(p, B, H, W) - align_up((size_t)B * 17 * 8, 1024));
      HP3D_CUDA(cudaMemsetAsync(keys, 0, (size_t)B * 17 * 8, s));
    }
    const bool staged = E->split && env_int("HP3D_CAST", 2) == 2;
    const __half* xh = (const __half*)x;
#define HP3D_CAST_LAUNCH(AM)                                                                                                        \
    do {                                                                                                                            \
      if (staged) { if (x_half) nchw_f32_to_split_records_kernel<AM, __half><<<cgrid, 256, 0, s>>>(xh, 18, H * W, xin, eps, keys, in_scale);          \
                    else nchw_f32_to_split_records_kernel<AM, float><<<cgrid, 256, 0, s>>>(x, 18, H * W, xin, eps, keys, in_scale); }                  \
      else if (E->split) { if (x_half) nchw_f32_to_nhwc32_f16_kernel<AM, true, __half><<<cgrid, 256, 0, s>>>(xh, 18, H * W, xin, eps, keys, in_scale); \
                           else nchw_f32_to_nhwc32_f16_kernel<AM, true, float><<<cgrid, 256, 0, s>>>(x, 18, H * W, xin, eps, keys, in_scale); }        \
      else { if (x_half) nchw_f32_to_nhwc32_f16_kernel<AM, false, __half><<<cgrid, 256, 0, s>>>(xh, 18, H * W, xin, eps, keys, in_scale);             \
             else nchw_f32_to_nhwc32_f16_kernel<AM, false, float><<<cgrid, 256, 0, s>>>(x, 18, H * W, xin, eps, keys, in_scale); }                     \
    } while (0)
    if (amax) HP3D_CAST_LAUNCH(true); else HP3D_CAST_LAUNCH(false);
#undef HP3D_CAST_LAUNCH
    rc = launch_status("proxy representation cast kernel");
    if (!rc && amax) {
      argmax_decode_kernel<<<cdiv(B * 17, 128), 128, 0, s>>>(keys, B * 17, W, amax->joints2d_px, amax->vis);
      rc = launch_status("argmax_decode_kernel");
    }
  }
  if (rc) return rc;
  rc = run_tc_conv(E, E->stem, xin, B, H, W, nullptr, 1, stem, s);
  if (rc) return rc;
  int ch = H / 2, cw = W / 2;
  rc = tap_copy_act(E, stem, (size_t)B * ch * cw, 64, &taps, s);
  if (rc) return rc;
  {
    const size_t total8 = (size_t)B * (ch / 2) * (cw / 2) * 64 / 8;
    if (E->split) maxpool3x3s2_split_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, s>>>(stem, ch, cw, 64, ch / 2, cw / 2, buf[0], total8);
    else maxpool3x3s2_f16_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, s>>>(stem, ch, cw, 64, ch / 2, cw / 2, buf[0], total8);
    rc = launch_status("maxpool3x3s2_f16_kernel");
    if (rc) return rc;
    ch /= 2; cw /= 2;
    rc = tap_copy_act(E, buf[0], (size_t)B * ch * cw, 64, &taps, s);
    if (rc) return rc;
  }
  __half* cur = buf[0];
  int free_idx[3] = {1, 2, 3};
  int C = 64;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < 2; ++b) {
      const TcConv& c1 = E->conv[l][b][0];
      const TcConv& c2 = E->conv[l][b][1];
      __half* t = buf[free_idx[0]];
      __half* y = buf[free_idx[1]];
      __half* d = buf[free_idx[2]];
      rc = run_tc_conv(E, c1, cur, B, ch, cw, nullptr, 1, t, s);
      if (rc) return rc;
      const int oh = ch / c1.stride, ow = cw / c1.stride;
      const __half* identity = cur;
      if (b == 0 && E->has_down[l]) {
        rc = run_tc_conv(E, E->down[l], cur, B, ch, cw, nullptr, 0, d, s);
        if (rc) return rc;
        identity = d;
      }
      rc = run_tc_conv(E, c2, t, B, oh, ow, identity, 1, y, s);
      if (rc) return rc;
      int cur_idx = 0;
      for (int i = 0; i < 4; ++i) if (buf[i] == cur) cur_idx = i;
      const int y_idx = free_idx[1];
      free_idx[1] = cur_idx;
      cur = buf[y_idx];
      ch = oh; cw = ow; C = c1.cout;
      rc = tap_copy_act(E, cur, (size_t)B * ch * cw, C, &taps, s);
      if (rc) return rc;
    }
  if (E->split) avgpool_split_kernel<<<B, 256, 0, s>>>(cur, ch * cw, C, feats, 1.f / in_scale);
  else avgpool_f16_kernel<<<B, 256, 0, s>>>(cur, ch * cw, C, feats);
  return launch_status("avgpool_f16_kernel");
}

}  // namespace hp3d
