// Probe: does TMA multicast inside a thread-block cluster lift the per-SM L2 -> SM ingest limit that
// bounds the few-token projector GEMM ("Linear at M <= 256")?
//
// 112 CTAs replay the load pattern of Linear-2 at M = 256 (K = 3584: 56 k-blocks; per k-block a 16 KB token
// tile that every n-tile CTA of the same m-tile needs, and an 8 KB weight tile of its own) with no MMA
// behind it, in three modes:
//   unicast       every CTA fetches its token tile itself                      (24 KB requested per k-block)
//   multicast S   the S CTAs of a cluster (neighbouring n-tiles, same m-tile) fetch 1/S of the token tile
//                 each and multicast it to all S                              (8 + 16/S KB requested)
// Every CTA RECEIVES 24 KB per k-block in all modes.  If multicast is no faster, the limit is the SM's
// receive port and the GEMM cannot gain from it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_probe/mcast_probe tools/mcast_probe.cu
//   timeout 120 build_probe/mcast_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int kStages = 8;
constexpr int kXBytes = 128 * 64 * 2;     // token tile of one k-block
constexpr int kWBytes = 64 * 64 * 2;      // weight tile of one k-block (BN = 64)
constexpr int kStageBytes = kXBytes + kWBytes;
constexpr int kKBlocks = 56;
constexpr int kTilesN = 56;
constexpr int kCtas = 112;
constexpr uint32_t kSpinLimit = 1u << 26;   // a protocol bug traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (++spins > kSpinLimit) __trap();
  }
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
template <bool RELAXED>
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  if (RELAXED)
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
  else
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_mcast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// S = cluster size (1 = unicast, no cluster features used)
template <int S, bool MC, bool RELAXED>
__global__ void __launch_bounds__(64, 1)
probe_kernel(const uint8_t* __restrict__ x, const uint8_t* __restrict__ w, unsigned long long* __restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  const int tile = blockIdx.x;
  const int m_tile = tile / kTilesN, n_tile = tile % kTilesN;
  const uint32_t rank = S > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], S);      // every CTA of the cluster releases the stage
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (S > 1) cluster_sync(); else __syncthreads();

  const uint8_t* xs = x + size_t(m_tile) * kKBlocks * kXBytes;
  const uint8_t* ws = w + size_t(n_tile) * kKBlocks * kWBytes;
  if (threadIdx.x == 0) {               // producer
    for (int kb = 0; kb < kKBlocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* dst = smem + size_t(s) * kStageBytes;
      mbar_expect_tx(&full_bar[s], kStageBytes);
      if (S == 1 || !MC) {
        bulk_g2s(dst, xs + size_t(kb) * kXBytes, kXBytes, &full_bar[s]);
      } else {
        constexpr int kPart = kXBytes / S;
        bulk_g2s_mcast(dst + rank * kPart, xs + size_t(kb) * kXBytes + rank * kPart, kPart, &full_bar[s],
                       uint16_t((1u << S) - 1u));
      }
      bulk_g2s(dst + kXBytes, ws + size_t(kb) * kWBytes, kWBytes, &full_bar[s]);
    }
  } else if (threadIdx.x >= 32) {       // consumer warp: lane 0 touches one word per stage, lane r releases it in CTA r
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long acc = 0;
    for (int kb = 0; kb < kKBlocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&full_bar[s], ph);
      if (lane == 0) acc += *reinterpret_cast<const volatile uint32_t*>(smem + size_t(s) * kStageBytes + 4 * (kb & 63));
      __syncwarp();
      if (S == 1) {
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
      } else if (lane < uint32_t(S)) {
        mbar_arrive_remote<RELAXED>(&empty_bar[s], lane);
      }
    }
    if (acc == 0x1234567ull) sink[tile] = acc;
  }
  __syncwarp();
  if (S > 1) cluster_sync();            // nobody leaves while a peer may still signal its barriers
}

template <int S, bool MC, bool RELAXED = false> static float run(const uint8_t* x, const uint8_t* w, unsigned long long* sink, uint8_t* flush, size_t flush_bytes, bool cold) {
  const size_t smem = size_t(kStages) * kStageBytes;
  cudaFuncSetAttribute(probe_kernel<S, MC, RELAXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCtas);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f, sum = 0.f;
  const int reps = 20;
  for (int i = 0; i < reps + 3; ++i) {
    if (cold) cudaMemsetAsync(flush, i, flush_bytes);
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, probe_kernel<S, MC, RELAXED>, x, w, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) {
      printf("S=%d: %s\n", S, cudaGetErrorString(err));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 3) {
      best = ms < best ? ms : best;
      sum += ms;
    }
  }
  printf("  cluster %d  %-9s %-7s %-5s  mean %.2f us  best %.2f us\n", S, MC ? "multicast" : "unicast", RELAXED ? "relaxed" : "release", cold ? "cold" : "warm", 1e3f * sum / reps, 1e3f * best);
  return best;
}

int main() {
  uint8_t *x, *w, *flush;
  unsigned long long* sink;
  const size_t xb = size_t(2) * kKBlocks * kXBytes, wb = size_t(kTilesN) * kKBlocks * kWBytes, fb = size_t(512) << 20;
  cudaMalloc(&x, xb);
  cudaMalloc(&w, wb);
  cudaMalloc(&flush, fb);
  cudaMalloc(&sink, kCtas * 8);
  cudaMemset(x, 1, xb);
  cudaMemset(w, 1, wb);
  printf("load pattern of Linear-2 at M = 256: %d CTAs x %d k-blocks, %d KB received per CTA and k-block\n", kCtas,
         kKBlocks, kStageBytes / 1024);
  printf("bytes received per CTA: %.2f MB; weights %.1f MB, tokens %.1f MB\n", kKBlocks * kStageBytes / 1e6, wb / 1e6, xb / 1e6);
  for (int cold = 0; cold < 2; ++cold) {
    run<1, false>(x, w, sink, flush, fb, cold);
    run<2, false>(x, w, sink, flush, fb, cold);   // cluster-wide stage release, but every CTA loads for itself
    run<4, false>(x, w, sink, flush, fb, cold);
    run<2, true>(x, w, sink, flush, fb, cold);
    run<4, true>(x, w, sink, flush, fb, cold);
    run<2, false, true>(x, w, sink, flush, fb, cold);
    run<4, false, true>(x, w, sink, flush, fb, cold);
    run<2, true, true>(x, w, sink, flush, fb, cold);
    run<4, true, true>(x, w, sink, flush, fb, cold);
    run<8, true, true>(x, w, sink, flush, fb, cold);
  }
  return 0;
}
