// Kernel 4: one Linear (+ exact GELU) of the object projector,
//   y[m, n] = act(x[m, k] . w[n, k]^T + bias[n])
// Replaces nn.Linear / nn.GELU of feat_linear (reference ufvideo/model/layer.py:55-59,126).
//
// bf16 / fp16: tcgen05 tensor-core GEMM.  x and w are both K-major, so a 128 x BN output tile
// is D[tmem] = A[smem] . B[smem]^T with A = 128 rows of x, B = BN rows of w.  Warp roles:
//   warp 0   TMA producer: 128B-swizzled 2-D tensor-map loads of the A / B k-blocks into a
//            ring of shared-memory stages (mbarrier full / empty pairs)
//   warp 1   allocates TMEM (two accumulators); one elected lane issues tcgen05.mma (UMMA
//            128 x BN x 16, fp32 accumulate in TMEM) and commits stage release / accumulator-ready
//   warps 2-9  epilogue: tcgen05.ld the accumulator (thread = row), + bias, round to the model
//            dtype, GELU(erf), round, 16-byte stores; runs under the next tile's main loop
// The kernel is persistent (one CTA per SM walks the tile list).
// fp32: CUDA-core tiled GEMM (the 1e-5 fp32 tolerance rules out bf16/tf32 tensor-core inputs).
#include "common.cuh"

#include <cuda.h>
#include <cstdlib>

namespace ufv {

int make_tensor_map_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                       uint32_t box_rows, uint32_t box_cols, int swizzle128);

__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

// Exact-erf GELU for the tensor-core epilogue, whose result is rounded to bf16 / fp16 anyway:
// Phi(v) = 0.5 * erfc(-v / sqrt(2)) with erfc(z) = P(t) * exp(-z^2), t = 1 / (1 + p z) (Abramowitz &
// Stegun 7.1.26, |error| <= 1.5e-7 -- far below half a 16-bit ulp).  Evaluated on |v| and mirrored,
// so the negative tail has no 1 + erf cancellation.  14 FP32 instructions + 2 MUFU per element,
// against ~35 for erff(): with K = 1152 the epilogue would otherwise outlast the main loop.
__device__ __forceinline__ float gelu_erf_16bit(float v) {
  const float av = fabsf(v);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(av, 0.3275911f * 0.70710678118654752440f, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * (v * -0.72134752044448170368f)));   // exp(-v^2 / 2)
  float p = 0.5f * 1.061405429f;
  p = fmaf(p, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float q = p * t * e;                      // Phi(-|v|)
  return v * (v > 0.0f ? 1.0f - q : q);
}

// d/dv GELU(v) = Phi(v) + v * phi(v) for the backward epilogue (same erfc form; phi = exp(-v^2 / 2) / sqrt(2 pi)).
__device__ __forceinline__ float gelu_erf_grad_16bit(float v) {
  const float av = fabsf(v);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(av, 0.3275911f * 0.70710678118654752440f, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * (v * -0.72134752044448170368f)));   // exp(-v^2 / 2)
  float p = 0.5f * 1.061405429f;
  p = fmaf(p, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float q = p * t * e;                      // Phi(-|v|)
  const float cdf = v > 0.0f ? 1.0f - q : q;
  return fmaf(v * 0.39894228040143267794f, e, cdf);
}

// epilogue modes of the tensor-core kernel
constexpr int kEpiNone = 0;      // y = acc + bias
constexpr int kEpiGelu = 1;      // y = GELU(round(acc + bias)); aux (optional): the rounded pre-activation is stored there too
constexpr int kEpiGeluBwd = 2;   // y = (acc + bias) * GELU'(aux): aux = the forward's pre-activation (dgrad through GELU)

// ================================= tcgen05 path ===================================================
constexpr int kBM = 128;
constexpr int kBK = 64;             // 64 x 2 B = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;        // two per TMEM lane quarter, each takes every other 32-column chunk
constexpr int kGemmThreads = 32 * (2 + kEpiWarps);
constexpr int kTileBudget = 224 * 1024;   // shared memory for the stage ring (227 KB per CTA on sm_100)
constexpr int kPeerStageBytes = 20 * 1024;   // epilogue staging of the fused all-gather variant (static)

template <int BN, bool PEER = false> struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBudget = kTileBudget - (PEER ? kPeerStageBytes : 0);
  static constexpr int kStages = kBudget / kStageBytes < 8 ? kBudget / kStageBytes : 8;
  static constexpr int kSmem = kStages * kStageBytes + 1024;   // + alignment slack
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128
                                   : 2 * BN <= 256 ? 256 : 512;   // two accumulators
};

// K-major, 128-byte-swizzled shared-memory matrix descriptor (sm_100 format, version 1):
// rows are 128 B apart inside an 8-row swizzle atom, atoms are 1024 B apart (SBO).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);       // start address, 16-byte units
  d |= uint64_t(1) << 16;                            // leading byte offset (unused for SW128 K-major)
  d |= uint64_t(1024 >> 4) << 32;                    // stride byte offset
  d |= uint64_t(1) << 46;                            // descriptor version (sm_100)
  d |= uint64_t(2) << 61;                            // layout: SWIZZLE_128B
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc(int a_fmt, int bn) {
  return (1u << 4)                      // accumulator format: F32
         | (uint32_t(a_fmt) << 7)       // A format: 0 = F16, 1 = BF16
         | (uint32_t(a_fmt) << 10)      // B format
         | (uint32_t(bn >> 3) << 17)    // N
         | (uint32_t(kBM >> 4) << 24);  // M
}

template <typename T> struct Pack8;
template <> struct Pack8<__nv_bfloat16> {
  __device__ static uint32_t two(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ static float2 unpack(uint32_t r) {
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
};
template <> struct Pack8<__half> {
  __device__ static uint32_t two(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ static float2 unpack(uint32_t r) { return __half22float2(*reinterpret_cast<const __half2*>(&r)); }
};

// Tile order: m-tiles are taken in groups of kGroupM; inside a group the m-tile index runs fastest,
// then the n-tile.  CTAs running side by side therefore share weight tiles through L2, and the
// 2048 token rows of a group (<= 15 MB) stay L2-resident while every n-tile passes over them --
// without the grouping, M = 16384 re-reads the 117 MB activation matrix from HBM once per n-tile.
constexpr int kGroupM = 16;
__device__ __forceinline__ void tile_origin(int tile, int tiles_m, int tiles_n, int bn, int& m0, int& n0) {
  const int per_group = kGroupM * tiles_n;
  const int g = tile / per_group;
  const int r = tile - g * per_group;
  const int gm = min(kGroupM, tiles_m - g * kGroupM);   // m-tiles in this (possibly last, short) group
  const int nt = r / gm;
  m0 = (g * kGroupM + (r - nt * gm)) * kBM;
  n0 = nt * bn;
}

// Persistent: grid = min(tiles, SMs); CTA b takes tiles b, b + grid, ...  Three pipelines: the shared-memory stage
// ring (TMA -> MMA), the two TMEM accumulators (MMA -> epilogue: the epilogue of tile i overlaps
// the main loop of tile i + 1), and the tile loop itself.
// ---- stores of the fused all-gather variant ----------------------------------------------------------
__device__ __forceinline__ void st_multimem_v4(uint64_t addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}
__device__ __forceinline__ void st_multimem_u32(uint64_t addr, uint32_t v) {
  asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u32(uint64_t addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_multimem_release_u32(uint64_t addr, uint32_t v) {
  asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
// 16 bytes at byte offset `off` of this rank's rows in every destination copy (multimem / dst0 are the
// caller's register copies of peer.multimem / peer.dst[0]: the struct may live in global memory)
__device__ __forceinline__ void peer_store16(const ufv_peer_args& peer, bool multimem, uint64_t dst0, size_t off,
                                             uint4 v) {
  if (multimem) {
    st_multimem_v4(dst0 + off, v);
  } else {
    for (int d = 0; d < peer.n_dst; ++d) *reinterpret_cast<uint4*>(peer.dst[d] + off) = v;
  }
}

// Closing protocol of the fused all-gather.  Called by every thread of every CTA behind a __syncthreads()
// that follows the CTA's last remote store.  What sits between the last tile and the arrival flag is ONE
// NVLink round trip: thread 0 alone executes a system-scope fence -- it is cumulative over the stores of the
// CTA's other threads, which happen-before it through the barrier -- takes a ticket (local L2 atomic), and
// whoever takes the last ticket raises this rank's flag in every destination with a release store:
// flag >= flag_value  =>  every byte of the call has landed.  The tail (reserved rows + token counts, written
// by the merge kernel long before) is forwarded by CTA 0 ahead of its own fence, so nobody waits for it.
// (Round 1 fenced in all 320 threads, then let the last CTA forward the tail and fence again: three
// dependent round trips, 16 us at 2 GPUs -- profiles/r02c_timeline_n2.txt.)
__device__ __forceinline__ void peer_finish(const ufv_peer_args& peer, unsigned n_ctas) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && peer.tail_words > 0) {
    for (int wi = threadIdx.x; wi < peer.tail_words; wi += kGemmThreads) {
      const uint32_t v = __ldcg(reinterpret_cast<const uint32_t*>(peer.tail_src) + wi);
      if (peer.multimem) {
        st_multimem_u32(peer.tail_dst[0] + 4ull * wi, v);
      } else {
        for (int d = 0; d < peer.n_dst; ++d) *reinterpret_cast<uint32_t*>(peer.tail_dst[d] + 4ull * wi) = v;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  __threadfence_system();
  if (atomicAdd(peer.ticket, 1u) != n_ctas - 1) return;
  __threadfence();                     // acquire side of the ticket: the other CTAs' fenced stores precede the flag
  if (peer.multimem) {
    st_multimem_release_u32(peer.flag[0], uint32_t(peer.flag_value));
  } else {
    for (int d = 0; d < peer.n_dst; ++d) st_release_sys_u32(peer.flag[d], uint32_t(peer.flag_value));
  }
  *peer.ticket = 0u;                   // self-reset for the next call
}

template <typename T, int BN, int EPI, bool PEER>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const T* __restrict__ bias, T* __restrict__ y, int m, int n, int k, int tiles_m,
                 int n_tiles, const __grid_constant__ ufv_peer_args peer_param,
                 const ufv_dyn_args* __restrict__ dyn, const int32_t* __restrict__ row_map, T* __restrict__ aux) {
  constexpr bool GELU = EPI == kEpiGelu;
  using Cfg = GemmCfg<BN, PEER>;
  // fused all-gather: a warp's 32 x 32 sub-tile is turned around in shared memory so that its remote
  // stores are contiguous 64-byte row segments (4 lanes x 16 B) instead of 32 scattered 16-byte
  // pieces -- 16-byte NVLink writes ran at ~170 GB/s, and the packet rate, not the bandwidth, was the limit
  __shared__ uint4 s_stage[PEER ? kEpiWarps : 1][PEER ? 32 : 1][5];   // 80-byte rows: conflict-free both ways
  extern __shared__ uint8_t dyn_smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn_smem_raw) + 1023) &
                                              ~uintptr_t(1023));   // SW128 atoms need 1024-B alignment
  __shared__ __align__(8) uint64_t full_bar[Cfg::kStages];
  __shared__ __align__(8) uint64_t empty_bar[Cfg::kStages];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (k + kBK - 1) / kBK;
  const int tiles_n = n_tiles / tiles_m;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kEpiWarps);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem_base, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();                  // x (and, transitively, the parameters) come from earlier kernels
  pdl_launch_dependents();
  // graph replay: the output pointer / all-gather destinations of this call come through the device block
  if (dyn != nullptr) y = reinterpret_cast<T*>(dyn->tokens_out);
  const ufv_peer_args& peer = (PEER && dyn != nullptr) ? dyn->peer : peer_param;
  const bool peer_mm = PEER && peer.multimem != 0;
  const uint64_t peer_dst0 = PEER ? peer.dst[0] : 0;

  if (warp == 0) {
    // ------------------------------- TMA producer -----------------------------------------------
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int m0, n0;
        tile_origin(tile, tiles_m, tiles_n, BN, m0, n0);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* a_dst = tiles + size_t(s) * Cfg::kStageBytes;
          uint8_t* b_dst = a_dst + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
          tma_load_2d(a_dst, &tmap_x, kb * kBK, m0, &full_bar[s]);
          tma_load_2d(b_dst, &tmap_w, kb * kBK, n0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(Elem<T>::kDtype == UFV_BF16 ? 1 : 0, BN);
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
        mbar_wait(&acc_empty[acc], aph ^ 1u);   // the epilogue has drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * uint32_t(BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(tiles + size_t(s) * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
            const uint64_t da = smem_desc_sw128(a_addr + kk * kUmmaK * 2);
            const uint64_t db = smem_desc_sw128(b_addr + kk * kUmmaK * 2);
            umma_f16(d_tmem, da, db, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);   // stage reusable once these MMAs have read it
        }
        umma_commit(&acc_full[acc]);    // accumulator complete
      }
    }
  } else {
    // ------------------------------- epilogue warps -------------------------------------------------
    const int quarter = warp & 3;       // TMEM lanes a warp may read: 32 * (warp_id % 4) ..
    const int half = (warp - 2) >> 2;   // which of the two warps of this quarter
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      int m0, n0;
      tile_origin(tile, tiles_m, tiles_n, BN, m0, n0);
      const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
      mbar_wait(&acc_full[acc], aph);
      tc_fence_after_sync();
      const int row = m0 + quarter * 32 + lane;
      // scatter epilogue (the <region> splice): output row `row` lands in row row_map[row] of y, or nowhere (-1)
      const int dst_row = (!PEER && row_map != nullptr) ? (row < m ? __ldg(row_map + row) : -1) : row;
#pragma unroll 1
      for (int col0 = half * 32; col0 < BN; col0 += 64) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + acc * uint32_t(BN) + uint32_t(col0), v);
        tmem_ld_wait();
        const int gcol = n0 + col0;
        if ((PEER || (row < m && dst_row >= 0)) && gcol < n) {
          uint32_t packed[16];
          if (gcol + 32 <= n) {
            const uint4* bsrc = reinterpret_cast<const uint4*>(bias + gcol);   // warp-uniform: broadcast (bias may be null)
            T* aux_row = (EPI != kEpiNone && aux != nullptr && row < m) ? aux + size_t(row) * n + gcol : nullptr;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint4 bq = bias != nullptr ? __ldg(bsrc + q4) : make_uint4(0u, 0u, 0u, 0u);
              const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
              uint4 zq = make_uint4(0u, 0u, 0u, 0u);
              if (EPI == kEpiGeluBwd && aux_row != nullptr) zq = *reinterpret_cast<const uint4*>(aux_row + q4 * 8);
              const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
              uint32_t zout[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 bf = Pack8<T>::unpack(bw[j]);
                const int i = q4 * 8 + j * 2;
                float a = __uint_as_float(v[i]) + bf.x;
                float b = __uint_as_float(v[i + 1]) + bf.y;
                if (GELU) {   // the reference rounds to the model dtype before and after GELU
                  const float za = Elem<T>::to_f32(Elem<T>::from_f32(a)), zb = Elem<T>::to_f32(Elem<T>::from_f32(b));
                  zout[j] = Pack8<T>::two(za, zb);
                  a = gelu_erf_16bit(za);
                  b = gelu_erf_16bit(zb);
                }
                if (EPI == kEpiGeluBwd) {   // dgrad through GELU: the incoming gradient times GELU'(pre-activation)
                  const float2 z = Pack8<T>::unpack(zw[j]);
                  a *= gelu_erf_grad_16bit(z.x);
                  b *= gelu_erf_grad_16bit(z.y);
                }
                packed[i >> 1] = Pack8<T>::two(a, b);
              }
              if (GELU && aux_row != nullptr)     // training forward: keep the pre-activation for the backward pass
                *reinterpret_cast<uint4*>(aux_row + q4 * 8) = make_uint4(zout[0], zout[1], zout[2], zout[3]);
            }
            if (PEER) {                  // fused all-gather: the tile goes to every rank's gathered buffer
              uint4(*stage)[5] = s_stage[warp - 2];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                stage[lane][i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
              __syncwarp();
              const int row_base = m0 + quarter * 32;
#pragma unroll
              for (int i = 0; i < 4; ++i) {       // 8 rows per instruction, 4 lanes x 16 B per row
                const int r = i * 8 + (lane >> 2), c = lane & 3;
                if (row_base + r < m)
                  peer_store16(peer, peer_mm, peer_dst0, (size_t(row_base + r) * n + gcol) * sizeof(T) + 16 * c,
                               stage[r][c]);
              }
              __syncwarp();
            } else {
              T* dst = y + size_t(dst_row) * n + gcol;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<uint4*>(dst)[i] =
                    make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            }
          } else if (!PEER) {            // ragged right edge (n % 32 != 0); the gather variant requires n % 32 == 0
            T* dst = y + size_t(dst_row) * n + gcol;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (gcol + i < n) {
                float a = __uint_as_float(v[i]) + (bias != nullptr ? Elem<T>::to_f32(bias[gcol + i]) : 0.0f);
                if (GELU) {
                  const float za = Elem<T>::to_f32(Elem<T>::from_f32(a));
                  if (aux != nullptr) aux[size_t(row) * n + gcol + i] = Elem<T>::from_f32(za);
                  a = gelu_erf_16bit(za);
                }
                if (EPI == kEpiGeluBwd) a *= gelu_erf_grad_16bit(Elem<T>::to_f32(aux[size_t(row) * n + gcol + i]));
                dst[i] = Elem<T>::from_f32(a);
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);   // this warp's share of the accumulator is in registers / stored
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  if (PEER) peer_finish(peer, gridDim.x);
}

// ================================ split-K variant (few tokens) =====================================
// With M <= 256 tokens a Linear is a stream over its weights, and what bounds the persistent kernel
// above is not HBM but the 64 B/clk each SM can pull from L2: a full-K CTA of a 128 x BN tile ingests
// (128 + BN) x K x 2 bytes, of which the 128 token rows are the same bytes every other n-tile CTA pulls
// (M = 256, K = 3584, BN = 64: 1.38 MB per SM, 154 MB through the crossbar for 26 MB of weights).  Wide
// tiles cut the re-reads but leave too few CTAs -- so K is split over a thread-block CLUSTER of S CTAs:
// each computes the 128 x BN partial of its k-range in TMEM (BN = 256, S = 4: 0.67 MB per SM), then the
// cluster reduce-scatters the partials through L2: the tile is cut into 32 x 32 units (one tcgen05.ld of a
// warp), unit u belongs to CTA u mod S; every CTA stores the units it does not own into its fp32 slab
// (laid out so that a warp's store is one contiguous 512 B), one barrier.cluster (release / acquire,
// hardware co-scheduling: no flags, no spinning), then each owner adds the S partials of its units in the
// fixed order s = 0 .. S-1 -- its own straight from TMEM -- and runs the usual epilogue.  The result
// does not depend on timing, and not on M either (S is chosen from n and k alone inside this regime).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int BN> struct SplitCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBudget = kTileBudget - kPeerStageBytes;
  static constexpr int kStages = kBudget / kStageBytes < 8 ? kBudget / kStageBytes : 8;
  static constexpr int kSmem = kStages * kStageBytes + 1024;
  static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kChunks = BN / 32;              // 32-column chunks of the tile
  static constexpr int kUnitFloats = 32 * 32;          // one unit: 32 rows x 32 columns
  static constexpr int kTileFloats = kBM * BN;
};

template <typename T, int BN, bool GELU, bool PEER>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_splitk_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                     const T* __restrict__ bias, T* __restrict__ y, int m, int n, int k,
                     float* __restrict__ slab, const __grid_constant__ ufv_peer_args peer_param,
                     const ufv_dyn_args* __restrict__ dyn, const int32_t* __restrict__ row_map) {
  using Cfg = SplitCfg<BN>;
  __shared__ uint4 s_stage[PEER ? kEpiWarps : 1][PEER ? 32 : 1][5];
  extern __shared__ uint8_t dyn_smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[Cfg::kStages];
  __shared__ __align__(8) uint64_t empty_bar[Cfg::kStages];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_split = int(gridDim.x);                  // == cluster size: the cluster spans grid dimension x
  const int split = int(cluster_ctarank());
  const int n0 = int(blockIdx.y) * BN, m0 = int(blockIdx.z) * kBM;
  const int tile_id = int(blockIdx.z) * int(gridDim.y) + int(blockIdx.y);
  const int num_kb = (k + kBK - 1) / kBK;
  const int kb_begin = int((long long)split * num_kb / n_split);
  const int kb_end = int((long long)(split + 1) * num_kb / n_split);    // host guarantees n_split <= num_kb

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem_base, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();
  pdl_launch_dependents();
  if (dyn != nullptr) y = reinterpret_cast<T*>(dyn->tokens_out);
  const ufv_peer_args& peer = (PEER && dyn != nullptr) ? dyn->peer : peer_param;
  const bool peer_mm = PEER && peer.multimem != 0;
  const uint64_t peer_dst0 = PEER ? peer.dst[0] : 0;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = (it / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* a_dst = tiles + size_t(s) * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
        tma_load_2d(a_dst, &tmap_x, kb * kBK, m0, &full_bar[s]);
        tma_load_2d(a_dst + Cfg::kABytes, &tmap_w, kb * kBK, n0, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(Elem<T>::kDtype == UFV_BF16 ? 1 : 0, BN);
      for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = (it / Cfg::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(tiles + size_t(s) * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int kk = 0; kk < kBK / kUmmaK; ++kk)
          umma_f16(tmem_base, smem_desc_sw128(a_addr + kk * kUmmaK * 2), smem_desc_sw128(b_addr + kk * kUmmaK * 2),
                   idesc, (it | kk) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_full);
    }
  }

  // ---- reduce-scatter of the partial tiles through L2 ------------------------------------------------------
  const int quarter = warp & 3;             // TMEM lanes this warp may read (warps 2 .. 9 only)
  const int half = (warp - 2) >> 2;
  const uint32_t tmem_lane = tmem_base + (uint32_t(quarter * 32) << 16);
  float* my_slab = slab + (size_t(tile_id) * n_split + split) * Cfg::kTileFloats;
  if (warp >= 2) {
    mbar_wait(&acc_full, 0);
    tc_fence_after_sync();
    if (n_split > 1) {
#pragma unroll 1
      for (int c = half; c < Cfg::kChunks; c += 2) {
        if (n0 + c * 32 >= n || (quarter + c) % n_split == split) continue;      // outside the matrix / owned here
        uint32_t v[32];
        tmem_ld_32x32(tmem_lane + uint32_t(c * 32), v);
        tmem_ld_wait();
        uint4* dst = reinterpret_cast<uint4*>(my_slab + size_t(quarter * Cfg::kChunks + c) * Cfg::kUnitFloats) + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j * 32] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  if (n_split > 1) cluster_sync_all();      // every thread of every CTA: partial stores released, then acquired
  if (warp >= 2) {
    const int row = m0 + quarter * 32 + lane;
    const int dst_row = (!PEER && row_map != nullptr) ? (row < m ? __ldg(row_map + row) : -1) : row;
#pragma unroll 1
    for (int c = half; c < Cfg::kChunks; c += 2) {
      const int gcol = n0 + c * 32;
      if (gcol >= n || (n_split > 1 && (quarter + c) % n_split != split)) continue;
      float acc[32];
#pragma unroll 1
      for (int s = 0; s < n_split; ++s) {
        uint32_t v[32];
        if (s == split) {
          tmem_ld_32x32(tmem_lane + uint32_t(c * 32), v);
          tmem_ld_wait();
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(
                                 slab + (size_t(tile_id) * n_split + s) * Cfg::kTileFloats +
                                 size_t(quarter * Cfg::kChunks + c) * Cfg::kUnitFloats) + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 q4 = __ldcg(src + j * 32);
            v[4 * j] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = s == 0 ? __uint_as_float(v[i]) : __fadd_rn(acc[i], __uint_as_float(v[i]));
      }
      const bool live_row = PEER || (row < m && dst_row >= 0);   // nothing to store: the lane still stays in the loop
      uint32_t packed[16];
      const uint4* bsrc = reinterpret_cast<const uint4*>(bias + gcol);          // n % 32 == 0 on this path
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint4 bq = __ldg(bsrc + q4);
        const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 bf = Pack8<T>::unpack(bw[j]);
          const int i = q4 * 8 + j * 2;
          float a = acc[i] + bf.x;
          float b = acc[i + 1] + bf.y;
          if (GELU) {
            a = gelu_erf_16bit(Elem<T>::to_f32(Elem<T>::from_f32(a)));
            b = gelu_erf_16bit(Elem<T>::to_f32(Elem<T>::from_f32(b)));
          }
          packed[i >> 1] = Pack8<T>::two(a, b);
        }
      }
      if (PEER) {
        uint4(*stage)[5] = s_stage[warp - 2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          stage[lane][i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        __syncwarp();
        const int row_base = m0 + quarter * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2), cc = lane & 3;
          if (row_base + r < m)
            peer_store16(peer, peer_mm, peer_dst0, (size_t(row_base + r) * n + gcol) * sizeof(T) + 16 * cc, stage[r][cc]);
        }
        __syncwarp();
      } else if (live_row) {
        T* dst = y + size_t(dst_row) * n + gcol;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<uint4*>(dst)[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
      }
      __syncwarp();                            // the loop top holds warp-collective tcgen05.ld
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  if (PEER) peer_finish(peer, gridDim.x * gridDim.y * gridDim.z);
}

// an empty shard (no token rows) still has to forward its tail and raise its flag
__global__ void peer_tail_only_kernel(const __grid_constant__ ufv_peer_args peer) {
  pdl_wait();
  pdl_launch_dependents();
  for (int wi = threadIdx.x; wi < peer.tail_words; wi += blockDim.x) {
    const uint32_t v = __ldcg(reinterpret_cast<const uint32_t*>(peer.tail_src) + wi);
    if (peer.multimem) {
      st_multimem_u32(peer.tail_dst[0] + 4ull * wi, v);
    } else {
      for (int d = 0; d < peer.n_dst; ++d) *reinterpret_cast<uint32_t*>(peer.tail_dst[d] + 4ull * wi) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (peer.multimem) {
      st_multimem_release_u32(peer.flag[0], uint32_t(peer.flag_value));
    } else {
      for (int d = 0; d < peer.n_dst; ++d) st_release_sys_u32(peer.flag[d], uint32_t(peer.flag_value));
    }
  }
}

// ---- the same collection as a separate push, for callers that want it OFF the critical path --------------------
// With the stores fused into the last Linear, a rank's stream cannot move on before its 8 copies have crossed
// NVLink (a kernel is complete only when its stores are): at 8 GPUs every link has to absorb 8 x 1.8 MB per step,
// ~20 us that the next step's kernels -- which do not touch NVLink -- could be running under.  This kernel pushes
// rows the Linear wrote LOCALLY (this rank's slice of the symmetric buffer) to every destination and closes with
// the same protocol; launched on a side stream it overlaps the next step.  A few CTAs saturate the link.
__global__ void __launch_bounds__(kGemmThreads)
peer_push_kernel(const uint4* __restrict__ src, long long n_vec, const __grid_constant__ ufv_peer_args peer) {
  pdl_wait();
  pdl_launch_dependents();
  const bool mm = peer.multimem != 0;
  const uint64_t dst0 = peer.dst[0];
  const long long stride = (long long)gridDim.x * kGemmThreads;
  for (long long i = (long long)blockIdx.x * kGemmThreads + threadIdx.x; i < n_vec; i += stride)
    peer_store16(peer, mm, dst0, size_t(i) * 16, __ldcg(src + i));
  __syncthreads();
  peer_finish(peer, gridDim.x);
}

// ---- receiver side of the fused all-gather: spin until every rank's flag has reached `value` ---------
__global__ void wait_flags_kernel(const int32_t* __restrict__ flags, int n, int32_t value, long long timeout_ns,
                                  int32_t* __restrict__ timed_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = threadIdx.x;
  if (i >= n) return;
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
    if (v >= value) return;          // step counters only grow: a peer that is already further along also counts
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      if (timed_out != nullptr) {     // may point into pinned host memory: the host sees it without a sync
        *reinterpret_cast<volatile int32_t*>(timed_out) = 1;
        __threadfence_system();
      }
      return;
    }
  }
}

static int sm_count() {
  static const int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      v = 148;
    cudaGetLastError();
    return v;
  }();
  return n;
}

template <typename T, int BN>
static int launch_tc(const void* x, const void* w, const void* bias, void* y, int m, int n, int k,
                     int epi, const ufv_peer_args* peer, const ufv_dyn_args* dyn, const int32_t* row_map,
                     void* aux, cudaStream_t stream) {
  const int smem_bytes = peer != nullptr ? GemmCfg<BN, true>::kSmem : GemmCfg<BN, false>::kSmem;
  CUtensorMap tx, tw;
  int rc = make_tensor_map_2d(&tx, x, Elem<T>::kDtype, uint64_t(m), uint64_t(k), kBM, kBK, 1);
  if (rc != 0) return rc;
  rc = make_tensor_map_2d(&tw, w, Elem<T>::kDtype, uint64_t(n), uint64_t(k), BN, kBK, 1);
  if (rc != 0) return rc;
  const int tiles_m = (m + kBM - 1) / kBM;
  const int n_tiles = tiles_m * ((n + BN - 1) / BN);
  const dim3 grid(n_tiles < sm_count() ? n_tiles : sm_count());
  const int variant = peer != nullptr ? 3 : epi;
  auto kernel = variant == 3   ? linear_tc_kernel<T, BN, kEpiNone, true>
                : variant == 2 ? linear_tc_kernel<T, BN, kEpiGeluBwd, false>
                : variant == 1 ? linear_tc_kernel<T, BN, kEpiGelu, false>
                               : linear_tc_kernel<T, BN, kEpiNone, false>;
  static bool configured[4] = {false, false, false, false};   // idempotent attribute; a benign race sets it twice
  if (!configured[variant]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    configured[variant] = true;
  }
  static const ufv_peer_args no_peer = {};
  return check_launch("ufv_linear (tcgen05)",
                      launch_kernel(kernel, grid, dim3(kGemmThreads), smem_bytes, stream, tx, tw,
                                    static_cast<const T*>(bias), static_cast<T*>(y), m, n, k, tiles_m, n_tiles,
                                    peer != nullptr ? *peer : no_peer, dyn, row_map, static_cast<T*>(aux)));
}

// N-tile choice.  A launch costs (waves over the SMs) x (time of one tile).  Tile times per 64-deep
// k-block, in units fitted to B200 measurements (tools/gemm_sweep.py): narrow tiles are bound by the
// bytes a CTA pulls through its SM, (128 + BN) x 128 B; 128 x 128 and 128 x 256 tiles by the tensor
// pipe (a 256-wide tile takes 1.7x, not 2x, the time of a 128-wide one).  Small token counts thus
// take narrow tiles (more CTAs share the weight stream), large ones wide tiles.
// UFV_GEMM_BN overrides the choice (developer sweeps and tests).
static int choose_bn(int m, int n) {
  const char* env = getenv("UFV_GEMM_BN");
  const int forced = env ? atoi(env) : 0;
  if (forced == 32 || forced == 64 || forced == 128 || forced == 256) return forced;
  const int m_tiles = (m + kBM - 1) / kBM;
  static const struct { int bn; double tile_cost; } kChoices[] = {{32, 184.0}, {64, 216.0}, {128, 300.0}, {256, 512.0}};
  int best_bn = 32;
  double best_cost = -1.0;
  for (const auto& c : kChoices) {
    const long ctas = long(m_tiles) * ((n + c.bn - 1) / c.bn);
    const long waves = (ctas + sm_count() - 1) / sm_count();
    const double cost = double(waves) * c.tile_cost;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_bn = c.bn;
    }
  }
  return best_bn;
}

// ---- split-K configuration ------------------------------------------------------------------------------
// Chosen from (n, k) and the number of 128-token tiles only, so that inside one regime the summation
// order -- and with it every output bit -- is independent of how many tokens a call carries:
//   k >= 2048 (the 3584 -> 3584 Linear):   1-2 token tiles (M <= 256): BN = 256, S = 4  (<= 112 CTAs, 14 k-blocks each)
//                                          3-4 token tiles (M <= 512): BN = 256, S = 2
//   otherwise, or when the clusters would not fit one wave: full-K persistent kernel (S = 1).
// UFV_GEMM_SPLIT = 0 disables, = S forces S (with UFV_GEMM_SPLIT_BN = 64 | 128 | 256): developer sweeps.
struct SplitChoice { int s; int bn; };

// Measured on B200 (profiles/r02b_*): at M = 256 the split kernel's shorter main loop (0.67 instead of
// 1.38 MB per SM) is eaten by its extra phases (partial stores, cluster barrier, partial loads: ~5 us), 19.0
// against 17.6 us inside the step -- so it is opt-in (UFV_GEMM_SPLIT_DEFAULT=1) until that is fixed.
static bool split_by_default() {
  const char* e = getenv("UFV_GEMM_SPLIT_DEFAULT");
  return e != nullptr && atoi(e) != 0;
}

static SplitChoice choose_split(int m, int n, int k) {
  const char* env_s = getenv("UFV_GEMM_SPLIT");          // read per call: tests and sweeps toggle it
  const char* env_bn = getenv("UFV_GEMM_SPLIT_BN");
  const int tiles_m = (m + kBM - 1) / kBM;
  const int num_kb = (k + kBK - 1) / kBK;
  SplitChoice c{1, 256};
  if (n % 32 != 0) return c;
  if (env_s != nullptr) {
    c.s = atoi(env_s);
    if (env_bn != nullptr) c.bn = atoi(env_bn);
    if (c.bn != 64 && c.bn != 128 && c.bn != 256) c.bn = 256;
  } else if (split_by_default() && num_kb >= 32) {
    c.s = tiles_m <= 2 ? 4 : tiles_m <= 4 ? 2 : 1;
  }
  if (c.s < 2 || c.s > 8 || c.s > num_kb) return SplitChoice{1, c.bn};
  const long ctas = long(tiles_m) * ((n + c.bn - 1) / c.bn) * c.s;
  if (ctas > sm_count()) return SplitChoice{1, c.bn};     // the clusters must be co-resident in one wave
  return c;
}

static size_t split_ws_bytes(int m, int n, int k) {
  const SplitChoice c = choose_split(m, n, k);
  if (c.s <= 1) return 0;
  return size_t((m + kBM - 1) / kBM) * ((n + c.bn - 1) / c.bn) * c.s * kBM * c.bn * sizeof(float);
}

template <typename T, int BN>
static int launch_splitk(const void* x, const void* w, const void* bias, void* y, int m, int n, int k, int gelu,
                         int n_split, float* slab, const ufv_peer_args* peer, const ufv_dyn_args* dyn,
                         const int32_t* row_map, cudaStream_t stream) {
  CUtensorMap tx, tw;
  int rc = make_tensor_map_2d(&tx, x, Elem<T>::kDtype, uint64_t(m), uint64_t(k), kBM, kBK, 1);
  if (rc != 0) return rc;
  rc = make_tensor_map_2d(&tw, w, Elem<T>::kDtype, uint64_t(n), uint64_t(k), BN, kBK, 1);
  if (rc != 0) return rc;
  const dim3 grid(n_split, (n + BN - 1) / BN, (m + kBM - 1) / kBM);
  const int variant = peer != nullptr ? 2 : gelu ? 1 : 0;
  auto kernel = variant == 2   ? linear_splitk_kernel<T, BN, false, true>
                : variant == 1 ? linear_splitk_kernel<T, BN, true, false>
                               : linear_splitk_kernel<T, BN, false, false>;
  static bool configured[3] = {false, false, false};
  if (!configured[variant]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SplitCfg<BN>::kSmem);
    configured[variant] = true;
  }
  static const ufv_peer_args no_peer = {};
  return check_launch("ufv_linear (tcgen05, split-K cluster)",
                      launch_kernel_cluster(kernel, grid, dim3(kGemmThreads), SplitCfg<BN>::kSmem, stream,
                                            unsigned(n_split), tx, tw, static_cast<const T*>(bias),
                                            static_cast<T*>(y), m, n, k, slab, peer != nullptr ? *peer : no_peer, dyn,
                                            row_map));
}

template <typename T>
static int dispatch_tc(const void* x, const void* w, const void* bias, void* y, int m, int n, int k,
                       int gelu, const ufv_peer_args* peer, const ufv_dyn_args* dyn, void* ws, int64_t ws_bytes,
                       const int32_t* row_map, cudaStream_t stream, void* aux = nullptr) {
  const SplitChoice sc = choose_split(m, n, k);
  if (sc.s > 1 && ws != nullptr && size_t(ws_bytes) >= split_ws_bytes(m, n, k) && gelu <= 1 && aux == nullptr &&
      bias != nullptr) {
    float* slab = static_cast<float*>(ws);
    switch (sc.bn) {
      case 256: return launch_splitk<T, 256>(x, w, bias, y, m, n, k, gelu, sc.s, slab, peer, dyn, row_map, stream);
      case 128: return launch_splitk<T, 128>(x, w, bias, y, m, n, k, gelu, sc.s, slab, peer, dyn, row_map, stream);
      default: return launch_splitk<T, 64>(x, w, bias, y, m, n, k, gelu, sc.s, slab, peer, dyn, row_map, stream);
    }
  }
  switch (choose_bn(m, n)) {
    case 256: return launch_tc<T, 256>(x, w, bias, y, m, n, k, gelu, peer, dyn, row_map, aux, stream);
    case 128: return launch_tc<T, 128>(x, w, bias, y, m, n, k, gelu, peer, dyn, row_map, aux, stream);
    case 64: return launch_tc<T, 64>(x, w, bias, y, m, n, k, gelu, peer, dyn, row_map, aux, stream);
    default: return launch_tc<T, 32>(x, w, bias, y, m, n, k, gelu, peer, dyn, row_map, aux, stream);
  }
}

// ================================= fp32 CUDA-core path ============================================
constexpr int kSBM = 32, kSBN = 64, kSBK = 32, kSimtThreads = 256;

template <bool GELU>
__global__ void __launch_bounds__(kSimtThreads)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ bias, float* __restrict__ y, int m, int n, int k,
                  const int32_t* __restrict__ row_map) {
  __shared__ float sx[kSBK][kSBM + 1];
  __shared__ float sw[kSBK][kSBN + 1];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kSBM, n0 = blockIdx.x * kSBN;
  const int tr = tid / 16, tc = tid % 16;   // thread tile: rows tr*2.., cols tc*4..
  float acc[2][4] = {};
  pdl_wait();
  pdl_launch_dependents();
  for (int k0 = 0; k0 < k; k0 += kSBK) {
    for (int i = tid; i < kSBM * kSBK; i += kSimtThreads) {
      const int r = i / kSBK, kk = i % kSBK;
      sx[kk][r] = (m0 + r < m && k0 + kk < k) ? x[size_t(m0 + r) * k + k0 + kk] : 0.f;
    }
    for (int i = tid; i < kSBN * kSBK; i += kSimtThreads) {
      const int r = i / kSBK, kk = i % kSBK;
      sw[kk][r] = (n0 + r < n && k0 + kk < k) ? w[size_t(n0 + r) * k + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSBK; ++kk) {
      float a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = sx[kk][tr * 2 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sw[kk][tc * 4 + j];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = m0 + tr * 2 + i;
    if (row >= m) continue;
    const int dst_row = row_map != nullptr ? row_map[row] : row;      // scatter epilogue (-1: drop the row)
    if (dst_row < 0) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tc * 4 + j;
      if (col >= n) continue;
      float v = acc[i][j] + bias[col];
      if (GELU) v = gelu_erf(v);
      y[size_t(dst_row) * n + col] = v;
    }
  }
}

// Few tokens (m <= 16): the Linear is a stream over the weights.  One warp per two output columns reads
// their weight rows once with coalesced float4 loads and keeps all m x 2 dot products in registers; the
// m token rows (<= 229 KB) come through L1/L2.  224 CTAs of 8 warps for n = 3584.
constexpr int kSkinnyM = 16, kSkinnyCols = 2, kSkinnyWarps = 8;

template <bool GELU>
__global__ void __launch_bounds__(32 * kSkinnyWarps)
linear_f32_skinny_kernel(const float* __restrict__ x, const float* __restrict__ w,
                         const float* __restrict__ bias, float* __restrict__ y, int m, int n, int k,
                         const int32_t* __restrict__ row_map) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int col0 = (blockIdx.x * kSkinnyWarps + (threadIdx.x >> 5)) * kSkinnyCols;
  if (col0 >= n) return;
  float acc[kSkinnyM][kSkinnyCols];
#pragma unroll
  for (int r = 0; r < kSkinnyM; ++r)
#pragma unroll
    for (int c = 0; c < kSkinnyCols; ++c) acc[r][c] = 0.f;
  const bool has2 = col0 + 1 < n;
  const float* w0 = w + size_t(col0) * k;
  const float* w1 = w + size_t(has2 ? col0 + 1 : col0) * k;
  for (int k0 = lane * 4; k0 < k; k0 += 128) {       // k % 4 == 0 is checked by the caller
    const float4 a = *reinterpret_cast<const float4*>(w0 + k0);
    const float4 b = *reinterpret_cast<const float4*>(w1 + k0);
#pragma unroll
    for (int r = 0; r < kSkinnyM; ++r) {
      if (r < m) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + size_t(r) * k + k0));
        acc[r][0] = fmaf(v.x, a.x, fmaf(v.y, a.y, fmaf(v.z, a.z, fmaf(v.w, a.w, acc[r][0]))));
        acc[r][1] = fmaf(v.x, b.x, fmaf(v.y, b.y, fmaf(v.z, b.z, fmaf(v.w, b.w, acc[r][1]))));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kSkinnyM; ++r) {
    if (r < m) {
#pragma unroll
      for (int c = 0; c < kSkinnyCols; ++c) {
        float v = acc[r][c];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        const int dst_row = row_map != nullptr ? row_map[r] : r;
        if (lane == 0 && (c == 0 || has2) && dst_row >= 0) {
          v += bias[col0 + c];
          if (GELU) v = gelu_erf(v);
          y[size_t(dst_row) * n + col0 + c] = v;
        }
      }
    }
  }
}

}  // namespace ufv

extern "C" int64_t ufv_linear_ws_bytes(int m, int n, int k, int dtype) {
  if (m <= 0 || n <= 0 || k <= 0 || (dtype != UFV_BF16 && dtype != UFV_F16)) return 0;
  return int64_t(ufv::split_ws_bytes(m, n, k));
}

extern "C" int ufv_linear_scatter(const void* x, const void* w, const void* bias, void* y, int m, int n, int k,
                                  int dtype, int gelu, const int32_t* row_map, void* ws, int64_t ws_bytes,
                                  void* stream) {
  using namespace ufv;
  UFV_REQUIRE(ws == nullptr || aligned16(ws), UFV_E_ALIGN, "ufv_linear: ws must be 16-byte aligned");
  UFV_REQUIRE(m >= 0 && n >= 1 && k >= 1, UFV_E_SHAPE, "ufv_linear: m=%d n=%d k=%d", m, n, k);
  if (m == 0) return 0;
  UFV_REQUIRE(x && w && bias && y, UFV_E_NULL, "ufv_linear: null pointer");
  UFV_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y), UFV_E_ALIGN,
              "ufv_linear: x / w / y must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == UFV_F32 && m <= kSkinnyM && k % 4 == 0) {
    const int cols_per_cta = kSkinnyWarps * kSkinnyCols;
    auto kernel = gelu ? linear_f32_skinny_kernel<true> : linear_f32_skinny_kernel<false>;
    return check_launch("ufv_linear (fp32, skinny)",
                        launch_kernel(kernel, dim3((n + cols_per_cta - 1) / cols_per_cta), dim3(32 * kSkinnyWarps), 0,
                                      st, static_cast<const float*>(x), static_cast<const float*>(w),
                                      static_cast<const float*>(bias), static_cast<float*>(y), m, n, k, row_map));
  }
  if (dtype == UFV_F32) {
    const dim3 grid((n + kSBN - 1) / kSBN, (m + kSBM - 1) / kSBM);
    auto kernel = gelu ? linear_f32_kernel<true> : linear_f32_kernel<false>;
    return check_launch("ufv_linear (fp32)",
                        launch_kernel(kernel, grid, dim3(kSimtThreads), 0, st, static_cast<const float*>(x),
                                      static_cast<const float*>(w), static_cast<const float*>(bias),
                                      static_cast<float*>(y), m, n, k, row_map));
  }
  UFV_REQUIRE(dtype == UFV_BF16 || dtype == UFV_F16, UFV_E_DTYPE, "ufv_linear: unsupported dtype %d", dtype);
  UFV_REQUIRE(k % 8 == 0 && n % 8 == 0, UFV_E_SHAPE, "ufv_linear: k=%d and n=%d must be multiples of 8", k, n);
  if (dtype == UFV_BF16)
    return dispatch_tc<__nv_bfloat16>(x, w, bias, y, m, n, k, gelu, nullptr, nullptr, ws, ws_bytes, row_map, st);
  return dispatch_tc<__half>(x, w, bias, y, m, n, k, gelu, nullptr, nullptr, ws, ws_bytes, row_map, st);
}

extern "C" int ufv_linear(const void* x, const void* w, const void* bias, void* y, int m, int n, int k,
                          int dtype, int gelu, void* ws, int64_t ws_bytes, void* stream) {
  return ufv_linear_scatter(x, w, bias, y, m, n, k, dtype, gelu, nullptr, ws, ws_bytes, stream);
}

extern "C" int ufv_linear_ex(const void* x, const void* w, const void* bias, void* y, int m, int n, int k, int dtype,
                             int epilogue, void* aux, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(m >= 0 && n >= 1 && k >= 1, UFV_E_SHAPE, "ufv_linear_ex: m=%d n=%d k=%d", m, n, k);
  if (m == 0) return 0;
  UFV_REQUIRE(dtype == UFV_BF16 || dtype == UFV_F16, UFV_E_DTYPE, "ufv_linear_ex: dtype %d (bf16 / fp16 only)", dtype);
  UFV_REQUIRE(epilogue >= kEpiNone && epilogue <= kEpiGeluBwd, UFV_E_SHAPE, "ufv_linear_ex: epilogue %d", epilogue);
  UFV_REQUIRE(x && w && y && (epilogue != kEpiGeluBwd || aux), UFV_E_NULL, "ufv_linear_ex: null pointer");
  UFV_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(aux) && aligned16(bias), UFV_E_ALIGN,
              "ufv_linear_ex: buffers must be 16-byte aligned");
  UFV_REQUIRE(k % 8 == 0 && n % 8 == 0, UFV_E_SHAPE, "ufv_linear_ex: k=%d and n=%d must be multiples of 8", k, n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == UFV_BF16)
    return dispatch_tc<__nv_bfloat16>(x, w, bias, y, m, n, k, epilogue, nullptr, nullptr, nullptr, 0, nullptr, st, aux);
  return dispatch_tc<__half>(x, w, bias, y, m, n, k, epilogue, nullptr, nullptr, nullptr, 0, nullptr, st, aux);
}

// ---- helpers of the projector's backward pass ---------------------------------------------------------------
namespace ufv {
// out[c, r] = in[r, c] for 2-byte elements; out has `out_pitch` elements per row (>= rows; the columns
// [rows, out_pitch) are zero-filled so that the transposed matrix can serve as a K-major GEMM operand whose
// contraction length is padded to a multiple of 8).  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                          int rows, int cols, int out_pitch) {
  __shared__ uint16_t tile[32][34];
  pdl_wait();
  pdl_launch_dependents();
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[size_t(r) * cols + c] : uint16_t(0);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;                         // out row = c, out column = r
    if (c < cols && r < out_pitch) out[size_t(c) * out_pitch + r] = tile[tx][i];
  }
}

// out[c] = sum over rows of x[r, c], fp32 accumulation in ascending row order, rounded to T (bias gradients)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, T* __restrict__ out, int m, int n) {
  pdl_wait();
  pdl_launch_dependents();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n) return;
  float acc = 0.f;
  for (int r = 0; r < m; ++r) acc += Elem<T>::to_f32(x[size_t(r) * n + c]);
  out[c] = Elem<T>::from_f32(acc);
}
}  // namespace ufv

extern "C" int ufv_transpose16(const void* in, void* out, int rows, int cols, int out_pitch, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(rows >= 0 && cols >= 0 && out_pitch >= rows, UFV_E_SHAPE, "ufv_transpose16: rows=%d cols=%d out_pitch=%d",
              rows, cols, out_pitch);
  if (rows == 0 || cols == 0) return 0;
  UFV_REQUIRE(in && out, UFV_E_NULL, "ufv_transpose16: null pointer");
  const dim3 grid((cols + 31) / 32, (out_pitch + 31) / 32);
  return check_launch("ufv_transpose16",
                      launch_kernel(transpose16_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream),
                                    static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), rows, cols, out_pitch));
}

extern "C" int ufv_colsum(const void* x, void* out, int m, int n, int dtype, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(m >= 0 && n >= 1, UFV_E_SHAPE, "ufv_colsum: m=%d n=%d", m, n);
  UFV_REQUIRE(x && out, UFV_E_NULL, "ufv_colsum: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((n + 255) / 256);
  if (dtype == UFV_BF16)
    return check_launch("ufv_colsum", launch_kernel(colsum_kernel<__nv_bfloat16>, grid, dim3(256), 0, st,
                                                    static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), m, n));
  if (dtype == UFV_F16)
    return check_launch("ufv_colsum", launch_kernel(colsum_kernel<__half>, grid, dim3(256), 0, st,
                                                    static_cast<const __half*>(x), static_cast<__half*>(out), m, n));
  return fail(UFV_E_DTYPE, "ufv_colsum: dtype %d (bf16 / fp16 only)", dtype);
}

namespace ufv {
// last Linear of the chained path in graph-replay mode (tensor-core dtypes only): the output pointer and,
// with `peer`, the all-gather destinations are read from the device block `dyn` at run time
int last_linear_dyn(const void* x, const void* w, const void* bias, int m, int n, int k, int dtype,
                    const ufv_peer_args* peer, const ufv_dyn_args* dyn, void* ws, int64_t ws_bytes,
                    const int32_t* row_map, void* stream) {
  UFV_REQUIRE(dtype == UFV_BF16 || dtype == UFV_F16, UFV_E_DTYPE, "graph replay: bf16 / fp16 only (dtype %d)", dtype);
  UFV_REQUIRE(m >= 1 && k % 8 == 0 && n % 32 == 0, UFV_E_SHAPE, "graph replay: m=%d n=%d k=%d", m, n, k);
  UFV_REQUIRE(x && w && bias && dyn && aligned16(x) && aligned16(w), UFV_E_NULL, "graph replay: bad pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == UFV_BF16)
    return dispatch_tc<__nv_bfloat16>(x, w, bias, nullptr, m, n, k, 0, peer, dyn, ws, ws_bytes, row_map, st);
  return dispatch_tc<__half>(x, w, bias, nullptr, m, n, k, 0, peer, dyn, ws, ws_bytes, row_map, st);
}
}  // namespace ufv

extern "C" int ufv_linear_gather(const void* x, const void* w, const void* bias, int m, int n, int k,
                                 int dtype, const ufv_peer_args* peer, void* ws, int64_t ws_bytes, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(ws == nullptr || aligned16(ws), UFV_E_ALIGN, "ufv_linear_gather: ws must be 16-byte aligned");
  UFV_REQUIRE(m >= 0 && n >= 1 && k >= 1, UFV_E_SHAPE, "ufv_linear_gather: m=%d n=%d k=%d", m, n, k);
  UFV_REQUIRE(peer && (m == 0 || (x && w && bias)), UFV_E_NULL, "ufv_linear_gather: null pointer");
  UFV_REQUIRE(dtype == UFV_BF16 || dtype == UFV_F16, UFV_E_DTYPE,
              "ufv_linear_gather: dtype %d (bf16 / fp16 only)", dtype);
  UFV_REQUIRE(k % 8 == 0 && n % 32 == 0, UFV_E_SHAPE, "ufv_linear_gather: k=%d %% 8, n=%d %% 32 must be 0", k, n);
  UFV_REQUIRE(peer->n_dst >= 1 && peer->n_dst <= UFV_MAX_PEER_DST && (!peer->multimem || peer->n_dst == 1),
              UFV_E_SHAPE, "ufv_linear_gather: n_dst=%d multimem=%d", peer->n_dst, peer->multimem);
  UFV_REQUIRE(peer->ticket != nullptr && (peer->tail_words == 0 || peer->tail_src != nullptr), UFV_E_NULL,
              "ufv_linear_gather: ticket / tail_src is null");
  for (int d = 0; d < peer->n_dst; ++d)
    UFV_REQUIRE(peer->dst[d] != 0 && (peer->dst[d] & 15u) == 0 && peer->flag[d] != 0 &&
                    (peer->tail_words == 0 || peer->tail_dst[d] != 0),
                UFV_E_ALIGN, "ufv_linear_gather: destination %d is null or not 16-byte aligned", d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (m == 0)
    return check_launch("ufv_linear_gather (empty shard)",
                        launch_kernel(peer_tail_only_kernel, dim3(1), dim3(256), 0, st, *peer));
  UFV_REQUIRE(aligned16(x) && aligned16(w), UFV_E_ALIGN, "ufv_linear_gather: x / w must be 16-byte aligned");
  if (dtype == UFV_BF16)
    return dispatch_tc<__nv_bfloat16>(x, w, bias, nullptr, m, n, k, 0, peer, nullptr, ws, ws_bytes, nullptr, st);
  return dispatch_tc<__half>(x, w, bias, nullptr, m, n, k, 0, peer, nullptr, ws, ws_bytes, nullptr, st);
}

extern "C" int ufv_peer_push(const void* src, int64_t bytes, const ufv_peer_args* peer, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(peer != nullptr && bytes >= 0 && bytes % 16 == 0, UFV_E_SHAPE, "ufv_peer_push: bytes=%lld", (long long)bytes);
  UFV_REQUIRE(bytes == 0 || (src != nullptr && aligned16(src)), UFV_E_NULL, "ufv_peer_push: src is null or unaligned");
  UFV_REQUIRE(peer->n_dst >= 1 && peer->n_dst <= UFV_MAX_PEER_DST && (!peer->multimem || peer->n_dst == 1),
              UFV_E_SHAPE, "ufv_peer_push: n_dst=%d multimem=%d", peer->n_dst, peer->multimem);
  UFV_REQUIRE(peer->ticket != nullptr && (peer->tail_words == 0 || peer->tail_src != nullptr), UFV_E_NULL,
              "ufv_peer_push: ticket / tail_src is null");
  const long long n_vec = bytes / 16;
  long long ctas = (n_vec + kGemmThreads * 8 - 1) / (kGemmThreads * 8);      // >= 8 vectors per thread
  ctas = ctas < 1 ? 1 : ctas > 16 ? 16 : ctas;
  return check_launch("ufv_peer_push",
                      launch_kernel(peer_push_kernel, dim3(unsigned(ctas)), dim3(kGemmThreads), 0,
                                    static_cast<cudaStream_t>(stream), static_cast<const uint4*>(src), n_vec, *peer));
}

extern "C" int ufv_wait_flags(const int32_t* flags, int n, int32_t value, int timeout_ms, int32_t* timed_out,
                              void* stream) {
  using namespace ufv;
  UFV_REQUIRE(n >= 0 && n <= 1024, UFV_E_SHAPE, "ufv_wait_flags: n=%d", n);
  if (n == 0) return 0;
  UFV_REQUIRE(flags != nullptr, UFV_E_NULL, "ufv_wait_flags: flags is null");
  const long long ns = (timeout_ms > 0 ? (long long)timeout_ms : 2000LL) * 1000000LL;
  return check_launch("ufv_wait_flags",
                      launch_kernel(wait_flags_kernel, dim3(1), dim3((n + 31) / 32 * 32), 0,
                                    static_cast<cudaStream_t>(stream), flags, n, value, ns, timed_out));
}
