// All-gather of this rank's slice by PEER STORES over NVLink (the reference has no multi-GPU path).
//
// Why not NCCL / copy engines (measured, profiles/README.md "multi-GPU"): ncclAllGather's CTAs (hundreds of threads, ~100
// registers each) cannot become resident next to the hot path's kernels, which fill every SM's registers / shared memory --
// the gather simply serialises with the compute (4 GPUs: 17.7 ms = 8.75 + 9.0, whether or not SMs are left free), and
// copy-engine pushes top out at ~430 GB/s per rank. This kernel is built to CO-RESIDE instead: tiny CTAs (128 threads,
// <= 32 registers, no shared memory) that fit into whatever the compute kernels leave on an SM, each thread streaming
// 16-byte pieces of the local slice into the same offset of every peer's buffer (symmetric memory: the peers' buffers are
// mapped into this process). Outbound bytes = (world - 1) x slice, inbound the same: NVLink-bound, but hidden behind the
// next chunk's / next step's kernels. A system-scope fence at the end orders the stores before the completion flag the host
// side exchanges afterwards (distributed.SymmPush.fence).
#include "common.cuh"
#include <algorithm>

using namespace hp3d;

namespace {
constexpr int MAX_PEERS = 15;
struct PeerDst { uint4* p[MAX_PEERS]; int n; };

template <int UNROLL>
__global__ void __launch_bounds__(128) peer_push_kernel(const uint4* __restrict__ src, const PeerDst d, size_t n16) {
  const size_t step = (size_t)gridDim.x * 128 * UNROLL;
  for (size_t base = (size_t)blockIdx.x * 128 * UNROLL + threadIdx.x; base < n16; base += step) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = base + (size_t)u * 128;
      if (i < n16) v[u] = __ldcs(src + i);                       // streaming: the slice is not re-read by this kernel
    }
    for (int p = 0; p < d.n; ++p) {
      uint4* dst = d.p[p];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = base + (size_t)u * 128;
        if (i < n16) __stcs(dst + i, v[u]);
      }
    }
  }
  __threadfence_system();
}
// The same through an NVSwitch MULTICAST address (NVLS): one multimem.st per 16 bytes lands in the buffer of every GPU of
// the group (this one included), so a rank sends its slice ONCE instead of (world - 1) times -- unicast peer stores and
// copy-engine pushes both measured ~400 GB/s per rank at 4 GPUs, NCCL's NVLS all-gather 730 GB/s.
template <int UNROLL>
__global__ void __launch_bounds__(128) peer_push_multicast_kernel(const uint4* __restrict__ src, uint4* mc_dst, size_t n16) {
  const size_t step = (size_t)gridDim.x * 128 * UNROLL;
  for (size_t base = (size_t)blockIdx.x * 128 * UNROLL + threadIdx.x; base < n16; base += step) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = base + (size_t)u * 128;
      if (i < n16) v[u] = __ldcs(src + i);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = base + (size_t)u * 128;
      if (i < n16)
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i), "f"(__uint_as_float(v[u].x)),
                     "f"(__uint_as_float(v[u].y)), "f"(__uint_as_float(v[u].z)), "f"(__uint_as_float(v[u].w))
                     : "memory");
    }
  }
  __threadfence_system();
}
}  // namespace

extern "C" int hp3d_peer_push_multicast(const void* src, void* multicast_dst, size_t bytes, int ctas, void* stream) {
  HP3D_ARG(src && multicast_dst, "null argument");
  HP3D_ARG(bytes % 16 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)multicast_dst & 15) == 0, "16-byte multiples required");
  if (bytes == 0) return 0;
  const size_t n16 = bytes / 16;
  if (ctas <= 0) ctas = 296;
  const int grid = (int)std::min<size_t>((size_t)ctas, (n16 + 128 * 4 - 1) / (128 * 4));
  peer_push_multicast_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)multicast_dst, n16);
  return launch_status("peer_push_multicast_kernel");
}

extern "C" int hp3d_peer_push(const void* src, void* const* peer_dsts, int n_peers, size_t bytes, int ctas, void* stream) {
  HP3D_ARG(src && peer_dsts && n_peers > 0 && n_peers <= MAX_PEERS, "need 1..15 peer destinations");
  HP3D_ARG(bytes % 16 == 0 && ((uintptr_t)src & 15) == 0, "src and size must be 16-byte multiples");
  PeerDst d;
  d.n = n_peers;
  for (int i = 0; i < n_peers; ++i) {
    HP3D_ARG(peer_dsts[i] && ((uintptr_t)peer_dsts[i] & 15) == 0, "peer destinations must be 16-byte aligned");
    d.p[i] = (uint4*)peer_dsts[i];
  }
  if (bytes == 0) return 0;
  const size_t n16 = bytes / 16;
  if (ctas <= 0) ctas = 296;
  const int grid = (int)std::min<size_t>((size_t)ctas, (n16 + 128 * 4 - 1) / (128 * 4));
  peer_push_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>((const uint4*)src, d, n16);
  return launch_status("peer_push_kernel");
}
