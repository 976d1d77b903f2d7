// Kernel 1: mask resize + binarise -> per object-frame patch bitmask, count and index list,
// plus the per-group "union plan" the pool kernel streams from.
//
// Replaces F.interpolate(mask, (27, 27), 'bilinear', align_corners=False) followed by
// (mask > 0) and mask.sum (reference ufvideo/model/layer.py:139,143,145).  With non-negative
// masks the interpolated value is a sum of non-negative weight * value products, so
// "interp > 0" is the OR of the (at most four) taps whose weight is non-zero: integer-exact,
// and only 4 * n_out^2 mask elements are read instead of H * W.
//
// One CTA per object-frame.  (Round 1 also built a per-frame "union plan" here, by whichever CTA of a frame
// finished last; the pool kernel now derives what it needs from the bitmasks itself, per 32-patch window.)
#include "common.cuh"

#include <cmath>

namespace ufv {

// ---- host: tap table ------------------------------------------------------------------------------
// fp32 arithmetic exactly as ATen evaluates it (area_pixel_compute_source_index, then
// guard_index_and_lambda): scale = in / out; src = max(fma(scale, i + 0.5, -0.5), 0) -- ATen's
// `scale * (dst + 0.5) - 0.5` is compiled to ONE fused multiply-add on both of its back ends (gcc
// -ffp-contract for the CPU kernels, nvcc's default contraction for CUDA), i.e. a single rounding of the
// exact product minus 0.5.  Rounding the product first moves src across an integer for n_in = 3, 5, 9 and
// 2049 (out = 27) and changes one tap each.  i0 = min(floor(src), in - 1); lambda1 = clamp(src - i0, 0, 1);
// lambda0 = 1 - lambda1; i1 = i0 + (i0 < in - 1).  A tap is kept iff its lambda is > 0.
static void axis_taps(int n_in, int n_out, int32_t* t0, int32_t* t1) {
  const float scale = static_cast<float>(n_in) / static_cast<float>(n_out);
  for (int i = 0; i < n_out; ++i) {
    // the product scale * (i + 0.5) has at most 24 + 6 significant bits and the subtraction of 0.5 stays
    // inside 53 bits: the double expression is exact, and its single rounding to fp32 IS the fma result
    // (no dependence on the host compiler's contraction flags or on fmaf being hardware-backed)
    const double exact = static_cast<double>(scale) * (static_cast<double>(i) + 0.5) - 0.5;
    volatile float src = static_cast<float>(exact);
    if (src < 0.0f) src = 0.0f;
    int i0 = static_cast<int>(floorf(src));
    if (i0 > n_in - 1) i0 = n_in - 1;
    volatile float lam1 = src - static_cast<float>(i0);
    if (lam1 < 0.0f) lam1 = 0.0f;
    if (lam1 > 1.0f) lam1 = 1.0f;
    volatile float lam0 = 1.0f - lam1;
    const int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    t0[i] = lam0 > 0.0f ? i0 : -1;
    t1[i] = lam1 > 0.0f ? i1 : -1;
  }
}

static void shift_into_image(int32_t* t, int n, int offset, int extent) {
  for (int i = 0; i < n; ++i) {
    if (t[i] < 0) continue;
    const int s = t[i] - offset;
    t[i] = (s >= 0 && s < extent) ? s : -1;   // taps that land in the zero padding read zero
  }
}

// ---- device ---------------------------------------------------------------------------------------
// COCO run-length mask: cum[i] = end (exclusive) of run i in column-major pixel order; runs alternate
// off / on starting with off.  Pixel q is on iff the first run with cum[i] > q has an odd index.
__device__ __forceinline__ bool rle_positive(const int32_t* __restrict__ cum, int n_runs, int q) {
  int lo = 0, hi = n_runs;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cum + mid) <= q) lo = mid + 1; else hi = mid;
  }
  return lo < n_runs && (lo & 1);
}

// one tap of an object-frame's mask: dense planes are indexed row-major, run-length masks column-major
__device__ __forceinline__ bool tap_positive(const ufv_mask_desc& d, int r, int c);

__device__ __forceinline__ bool mask_positive(uint64_t base, int64_t off, int dtype) {
  switch (dtype) {
    case UFV_U8:
      return reinterpret_cast<const uint8_t*>(base)[off] != 0;
    case UFV_F32:
      return reinterpret_cast<const float*>(base)[off] > 0.0f;
    case UFV_BF16:
      return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off]) > 0.0f;
    default:
      return __half2float(reinterpret_cast<const __half*>(base)[off]) > 0.0f;
  }
}

__device__ __forceinline__ bool tap_positive(const ufv_mask_desc& d, int r, int c) {
  if (d.dtype == UFV_RLE) return rle_positive(reinterpret_cast<const int32_t*>(d.addr), d.pitch, c * d.aux + r);
  return mask_positive(d.addr, int64_t(r) * d.pitch + c, d.dtype);
}

// Tap mode for one dense element type: all 4 * ITERS tap loads of a thread are issued back to back (invalid
// taps read element 0 and are masked afterwards), so the thread waits for ONE memory round trip instead of
// one per tap -- with the dtype switch inside the tap the compiler serialised them (r01d: 65 % of the
// kernel's stall samples sat on the compares behind the individual loads).
template <typename E> __device__ __forceinline__ bool elem_positive(E v);
template <> __device__ __forceinline__ bool elem_positive<uint8_t>(uint8_t v) { return v != 0; }
template <> __device__ __forceinline__ bool elem_positive<float>(float v) { return v > 0.0f; }
template <> __device__ __forceinline__ bool elem_positive<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v) > 0.0f;
}
template <> __device__ __forceinline__ bool elem_positive<__half>(__half v) { return __half2float(v) > 0.0f; }

template <typename E> __device__ __forceinline__ uint32_t elem_bits(E v);
template <> __device__ __forceinline__ uint32_t elem_bits<uint8_t>(uint8_t v) { return v; }
template <> __device__ __forceinline__ uint32_t elem_bits<float>(float v) { return __float_as_uint(v); }
template <> __device__ __forceinline__ uint32_t elem_bits<__nv_bfloat16>(__nv_bfloat16 v) {
  return uint32_t(__bfloat16_as_ushort(v)) << 16;      // bf16 -> the fp32 with the same value
}
template <> __device__ __forceinline__ uint32_t elem_bits<__half>(__half v) { return __float_as_uint(__half2float(v)); }
template <typename E> __device__ __forceinline__ bool bits_positive(uint32_t b) { return __uint_as_float(b) > 0.0f; }
template <> __device__ __forceinline__ bool bits_positive<uint8_t>(uint32_t b) { return b != 0u; }

template <typename E, int ITERS, int THREADS>
__device__ __forceinline__ void gather_taps(const ufv_mask_desc& d, const int32_t* h0, const int32_t* h1,
                                            const int32_t* w0, const int32_t* w1, int n_out, int n_patch, int tid,
                                            bool (&on)[ITERS]) {
  static_assert(ITERS <= 3, "the register fence below lists 12 values");
  const E* __restrict__ base = reinterpret_cast<const E*>(d.addr);
  uint32_t v[3][4] = {};
  uint32_t ok = 0;                                        // bit 4 * it + t
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int p = it * THREADS + tid;
    const bool live = p < n_patch;
    const int i = live ? p / n_out : 0, jx = live ? p - i * n_out : 0;
    const int r[2] = {h0[i], h1[i]}, c[2] = {w0[jx], w1[jx]};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int rr = r[t >> 1], cc = c[t & 1];
      const bool valid = live && rr >= 0 && cc >= 0;
      ok |= uint32_t(valid) << (4 * it + t);
      v[it][t] = elem_bits<E>(base[valid ? int64_t(rr) * d.pitch + cc : int64_t(0)]);
    }
  }
  // register fence: every load above is issued before any compare below (ptxas otherwise interleaves them
  // three at a time and the thread pays four memory round trips instead of one)
  asm volatile("" : "+r"(v[0][0]), "+r"(v[0][1]), "+r"(v[0][2]), "+r"(v[0][3]), "+r"(v[1][0]), "+r"(v[1][1]),
                    "+r"(v[1][2]), "+r"(v[1][3]), "+r"(v[2][0]), "+r"(v[2][1]), "+r"(v[2][2]), "+r"(v[2][3]));
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    bool any = false;
#pragma unroll
    for (int t = 0; t < 4; ++t) any |= ((ok >> (4 * it + t)) & 1u) && bits_positive<E>(v[it][t]);
    on[it] = any;
  }
}

// "element > 0" flags of one 16-byte chunk, bit e = element e of the chunk
__device__ __forceinline__ uint32_t chunk_flags(uint4 v, int dtype) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t f = 0;
  switch (dtype) {
    case UFV_F32:
#pragma unroll
      for (int e = 0; e < 4; ++e) f |= uint32_t(__uint_as_float(w[e]) > 0.0f) << e;
      break;
    case UFV_U8:
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t t = __vcmpne4(w[e], 0u);   // 0xff per non-zero byte
        f |= ((t & 1u) | ((t >> 7) & 2u) | ((t >> 14) & 4u) | ((t >> 21) & 8u)) << (4 * e);
      }
      break;
    case UFV_BF16:
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        f |= uint32_t(__uint_as_float(w[e] << 16) > 0.0f) << (2 * e);
        f |= uint32_t(__uint_as_float(w[e] & 0xffff0000u) > 0.0f) << (2 * e + 1);
      }
      break;
    default:
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
        f |= uint32_t(h.x > 0.0f) << (2 * e);
        f |= uint32_t(h.y > 0.0f) << (2 * e + 1);
      }
      break;
  }
  return f;
}

// One 16-byte chunk of a source row in row mode.  Chunks are aligned to 16 bytes around the tap span
// [lo, hi) of the row, so the first and the last chunk of a row may reach past the span: those are
// assembled from element loads of the in-span elements only (zeros elsewhere), and the kernel never
// touches a byte outside the columns its taps name -- caller-owned pinned host memory included.
__device__ __forceinline__ uint4 load_chunk_clamped(uint64_t chunk_addr, uint64_t lo, uint64_t hi, int es) {
  if (chunk_addr >= lo && chunk_addr + 16 <= hi) return *reinterpret_cast<const uint4*>(chunk_addr);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  for (int b = 0; b < 16; b += es) {
    const uint64_t a = chunk_addr + b;
    if (a < lo || a + es > hi) continue;
    const uint32_t e = es == 4   ? *reinterpret_cast<const uint32_t*>(a)
                       : es == 2 ? uint32_t(*reinterpret_cast<const uint16_t*>(a))
                                 : uint32_t(*reinterpret_cast<const uint8_t*>(a));
    w[b >> 2] |= e << (8 * (b & 3));
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// exclusive prefix of per-word popcounts, computed by warp 0; s_prefix[UFV_BITS_WORDS] = total
__device__ __forceinline__ void word_prefix(const uint32_t* s_words, int32_t* s_prefix, int tid) {
  if (tid < 32) {
    const int mine = tid < UFV_BITS_WORDS ? __popc(s_words[tid]) : 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, d);
      if (tid >= d) incl += up;
    }
    if (tid < UFV_BITS_WORDS) s_prefix[tid] = incl - mine;
    if (tid == UFV_BITS_WORDS - 1) s_prefix[UFV_BITS_WORDS] = incl;
  }
}

constexpr int kRowChunks = 256;                  // 16-byte chunks of one staged row span (4 KB)
constexpr int kRowUnroll = 4;                    // chunk loads in flight per thread

// Two instantiations: ROWS = true can read in either mode and carries the 27 KB flag table (4 CTAs
// of 384 threads per SM); ROWS = false is the lean tap-mode kernel for batches whose masks all live
// in HBM (256 threads, ~1 KB of shared memory, 8 CTAs per SM -- the kernel is a latency chain per
// CTA, so resident CTAs are what hides it).
template <bool ROWS> struct PatchCfg {
  static constexpr int kThreads = ROWS ? 384 : 256;
  static constexpr int kMinCtas = ROWS ? 4 : 5;   // 5 x 256 threads: 51 registers, room for 12 tap loads in flight
  static constexpr int kFlagRows = ROWS ? 2 * UFV_MAX_PATCH_SIDE : 1;
  static constexpr int kFlagCols = ROWS ? kRowChunks : 1;
};

// Two ways to read a mask, chosen per object-frame (desc.flags bit 0 asks for row mode):
//   row mode  (column span of the taps <= ~4 KB per row): the CTA pulls the 2 * n_out source rows with
//             coalesced 16-byte loads, reduces every chunk to "element > 0" flags in shared memory and
//             picks the tap columns out of the flags.  Few, wide requests: this is what keeps PCIe
//             efficient when the mask lives in pinned HOST memory and is read in place.
//   tap mode  (default; also wide masks): every thread gathers the four taps of its patches directly --
//             the lowest latency for masks in HBM (15.7 us vs 20.3 us in row mode at 512 masks of 384 x 384).
template <bool ROWS>
__global__ void __launch_bounds__(PatchCfg<ROWS>::kThreads, PatchCfg<ROWS>::kMinCtas)
mask_to_patches_kernel(const ufv_mask_desc* __restrict__ desc, const int32_t* __restrict__ taps, int n_out,
                       uint32_t* __restrict__ bits_out, int32_t* __restrict__ cnt_out,
                       uint16_t* __restrict__ idx_out, int idx_pitch,
                       const uint32_t* __restrict__ dyn_src, uint32_t* __restrict__ dyn_dev) {
  __shared__ int32_t s_taps[4 * UFV_MAX_PATCH_SIDE];
  __shared__ uint32_t s_words[UFV_BITS_WORDS];
  __shared__ int32_t s_prefix[UFV_BITS_WORDS + 1];
  constexpr int kPatchThreads = PatchCfg<ROWS>::kThreads;
  __shared__ uint16_t s_flags[PatchCfg<ROWS>::kFlagRows][PatchCfg<ROWS>::kFlagCols];
  __shared__ int s_span[2];

  const int j = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  pdl_wait();                  // masks / descriptors may come from the previous kernel in the stream
  pdl_launch_dependents();     // the pool kernel may get resident and set up while this one runs
  // graph replay: CTA 0 forwards the caller's per-call block (pinned host memory) to device memory for
  // the later kernels; the PCIe read is requested here and only consumed by the store at the very end
  uint32_t dyn_word = 0;
  const bool dyn_copy = dyn_src != nullptr && j == 0 && tid < int(sizeof(ufv_dyn_args) / 4);
  if (dyn_copy) dyn_word = *reinterpret_cast<const volatile uint32_t*>(dyn_src + tid);
  const ufv_mask_desc d = desc[j];
  if (tid < 4 * n_out) s_taps[tid] = taps[d.tap_off + tid];
  __syncthreads();

  const int n_patch = n_out * n_out;
  const int32_t* h0 = s_taps;
  const int32_t* h1 = s_taps + n_out;
  const int32_t* w0 = s_taps + 2 * n_out;
  const int32_t* w1 = s_taps + 3 * n_out;
  const int es_shift = d.dtype == UFV_F32 ? 2 : d.dtype == UFV_U8 ? 0 : 1;
  const int es = 1 << es_shift;
  if (warp == 0) {   // column span [cmin, cmax] of the valid taps
    int lo = 0x7fffffff, hi = -1;
    if (lane < n_out) {
      const int a = w0[lane], b = w1[lane];
      if (a >= 0) { lo = min(lo, a); hi = max(hi, a); }
      if (b >= 0) { lo = min(lo, b); hi = max(hi, b); }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (lane == 0) { s_span[0] = lo; s_span[1] = hi; }
  }
  __syncthreads();
  const int cmin = s_span[0], cmax = s_span[1];
  const bool row_mode = ROWS && (d.flags & 1) != 0 && d.dtype != UFV_RLE && cmax >= 0 && ((cmax - cmin + 1) * es + 30) >> 4 <= kRowChunks;

  constexpr int kIters = (UFV_BITS_WORDS * 32 + kPatchThreads - 1) / kPatchThreads;
  bool on[kIters];
  if (ROWS && row_mode) {
    const int span_bytes = (cmax - cmin + 1) * es;
    const int nch = (span_bytes + 30) >> 4;        // chunks per source row, whatever its misalignment
    const int total = 2 * n_out * nch;             // s_taps[0 .. 2 * n_out) = h0 then h1: one slot per source row
    for (int f0 = tid; f0 < total; f0 += kRowUnroll * kPatchThreads) {
      uint4 v[kRowUnroll];
      int slot[kRowUnroll], chunk[kRowUnroll];
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) {       // all loads of the round in flight before any flag
        const int f = f0 + u * kPatchThreads;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        slot[u] = -1;
        if (f < total) {
          slot[u] = f / nch;
          chunk[u] = f - slot[u] * nch;
          const int r = s_taps[slot[u]];
          const uint64_t first = d.addr + uint64_t(int64_t(max(r, 0)) * d.pitch + cmin) * es;
          const int mis = int(first & 15u);
          if (r >= 0 && chunk[u] < ((mis + span_bytes + 15) >> 4))
            v[u] = load_chunk_clamped(first - mis + 16ull * chunk[u], first, first + span_bytes, es);
        }
      }
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u)
        if (slot[u] >= 0) s_flags[slot[u]][chunk[u]] = uint16_t(chunk_flags(v[u], d.dtype));
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int p = it * kPatchThreads + tid;
      bool hit = false;
      if (p < n_patch) {
        const int i = p / n_out, jx = p - i * n_out;
        const int c[2] = {w0[jx], w1[jx]};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int r = s_taps[s * n_out + i];
          if (r < 0) continue;
          const int mis = int((d.addr + uint64_t(int64_t(r) * d.pitch + cmin) * es) & 15u);
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (c[u] >= 0) {
              const int off = mis + (c[u] - cmin) * es;
              hit |= (s_flags[s * n_out + i][off >> 4] >> ((off & 15) >> es_shift)) & 1u;
            }
        }
      }
      on[it] = hit;
    }
  } else if (d.dtype == UFV_F32) {
    gather_taps<float, kIters, kPatchThreads>(d, h0, h1, w0, w1, n_out, n_patch, tid, on);
  } else if (d.dtype == UFV_U8) {
    gather_taps<uint8_t, kIters, kPatchThreads>(d, h0, h1, w0, w1, n_out, n_patch, tid, on);
  } else if (d.dtype == UFV_BF16) {
    gather_taps<__nv_bfloat16, kIters, kPatchThreads>(d, h0, h1, w0, w1, n_out, n_patch, tid, on);
  } else if (d.dtype == UFV_F16) {
    gather_taps<__half, kIters, kPatchThreads>(d, h0, h1, w0, w1, n_out, n_patch, tid, on);
  } else {                                  // run-length masks: one binary search per tap
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int p = it * kPatchThreads + tid;
      on[it] = false;
      if (p < n_patch) {
        const int i = p / n_out, jx = p - i * n_out;
        const int ra = h0[i], rb = h1[i], ca = w0[jx], cb = w1[jx];
        const bool t00 = (ra >= 0 && ca >= 0) && tap_positive(d, ra, ca);
        const bool t01 = (ra >= 0 && cb >= 0) && tap_positive(d, ra, cb);
        const bool t10 = (rb >= 0 && ca >= 0) && tap_positive(d, rb, ca);
        const bool t11 = (rb >= 0 && cb >= 0) && tap_positive(d, rb, cb);
        on[it] = t00 | t01 | t10 | t11;
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const uint32_t word = __ballot_sync(0xffffffffu, on[it]);
    const int wi = (it * kPatchThreads + tid) >> 5;
    if (lane == 0 && wi < UFV_BITS_WORDS) s_words[wi] = word;
  }
  __syncthreads();
  word_prefix(s_words, s_prefix, tid);
  __syncthreads();
  if (tid < UFV_BITS_WORDS) bits_out[size_t(j) * UFV_BITS_WORDS + tid] = s_words[tid];
  if (tid == 0) cnt_out[j] = s_prefix[UFV_BITS_WORDS];
  if (idx_out != nullptr) {
    for (int p = tid; p < n_patch; p += kPatchThreads) {
      const uint32_t word = s_words[p >> 5];
      const uint32_t bit = 1u << (p & 31);
      if (word & bit)
        idx_out[size_t(j) * idx_pitch + s_prefix[p >> 5] + __popc(word & (bit - 1u))] =
            static_cast<uint16_t>(p);
    }
  }
  if (dyn_copy) dyn_dev[tid] = dyn_word;
}

}  // namespace ufv

extern "C" int ufv_tap_table(int h, int w, int n_out, int pad_square, int32_t* taps_host) {
  UFV_REQUIRE(taps_host != nullptr, UFV_E_NULL, "ufv_tap_table: taps_host is null");
  UFV_REQUIRE(h >= 1 && w >= 1 && n_out >= 1 && n_out <= UFV_MAX_PATCH_SIDE, UFV_E_SHAPE,
              "ufv_tap_table: h=%d w=%d n_out=%d out of range", h, w, n_out);
  int eff_h = h, eff_w = w, top = 0, left = 0;
  if (pad_square) {   // layer.py:77-86: centred zero padding of the short side
    const int side = h > w ? h : w;
    top = (side - h) / 2;
    left = (side - w) / 2;
    eff_h = eff_w = side;
  }
  ufv::axis_taps(eff_h, n_out, taps_host, taps_host + n_out);
  ufv::axis_taps(eff_w, n_out, taps_host + 2 * n_out, taps_host + 3 * n_out);
  if (pad_square) {
    ufv::shift_into_image(taps_host, 2 * n_out, top, h);
    ufv::shift_into_image(taps_host + 2 * n_out, 2 * n_out, left, w);
  }
  return 0;
}

namespace ufv {
int launch_mask_to_patches(const ufv_mask_desc* desc, const int32_t* taps, int n_masks, int n_out, int any_row_mode,
                           uint32_t* bits_out, int32_t* cnt_out, uint16_t* idx_out, int idx_pitch,
                           const ufv_dyn_args* dyn_src, ufv_dyn_args* dyn_dev, void* stream) {
  UFV_REQUIRE(n_masks >= 0 && n_out >= 1 && n_out <= UFV_MAX_PATCH_SIDE, UFV_E_SHAPE,
              "ufv_mask_to_patches: n_masks=%d n_out=%d out of range", n_masks, n_out);
  if (n_masks == 0) return 0;
  UFV_REQUIRE(desc && taps && bits_out && cnt_out, UFV_E_NULL, "ufv_mask_to_patches: null pointer");
  UFV_REQUIRE(idx_out == nullptr || idx_pitch >= n_out * n_out, UFV_E_SHAPE,
              "ufv_mask_to_patches: idx_pitch %d < %d", idx_pitch, n_out * n_out);
  auto kernel = any_row_mode ? ufv::mask_to_patches_kernel<true> : ufv::mask_to_patches_kernel<false>;
  const int threads = any_row_mode ? ufv::PatchCfg<true>::kThreads : ufv::PatchCfg<false>::kThreads;
  return ufv::check_launch(
      "ufv_mask_to_patches",
      ufv::launch_kernel(kernel, dim3(n_masks), dim3(threads), 0, static_cast<cudaStream_t>(stream), desc,
                         taps, n_out, bits_out, cnt_out, idx_out, idx_pitch,
                         reinterpret_cast<const uint32_t*>(dyn_src), reinterpret_cast<uint32_t*>(dyn_dev)));
}
}  // namespace ufv

extern "C" int ufv_mask_to_patches(const ufv_mask_desc* desc, const int32_t* taps, int n_masks, int n_out,
                                   int any_row_mode, uint32_t* bits_out, int32_t* cnt_out, uint16_t* idx_out,
                                   int idx_pitch, void* stream) {
  return ufv::launch_mask_to_patches(desc, taps, n_masks, n_out, any_row_mode, bits_out, cnt_out, idx_out, idx_pitch,
                                     nullptr, nullptr, stream);
}
