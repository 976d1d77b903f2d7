// Kernel 3: temporal token merge (TTM).
//
// Replaces token_merge (reference ufvideo/model/layer.py:6-33), its per-object dispatch
// (layer.py:110-119) and the downcast to the model dtype (layer.py:123).  On a GPU the reference
// pays one device->host sync per frame per object (the python `if sim[0,i] < kth`) plus O(T)
// tiny launches.  Here two launches cover every object of the batch:
//   3a  ttm_sims_kernel   one warp per (object, adjacent pair):
//         m_t = max(sqrt(sum x_t^2), 1e-12)                              (F.normalize, :13-14)
//         s_i = sum (x_i / m_i) * (x_{i+1} / m_{i+1})                    (:15)
//   3b  ttm_merge_kernel  one CTA per (object, output slot): r-th largest s by rank counting
//         (torch.topk, :17-18), strict-below cuts (:24), the slot's run bounds by popcount
//         prefix, the run mean in ascending token order (:26,31), downcast, count.
// Work is spread over (objects x pairs) and (objects x K) CTAs, so a single 512-frame object
// fills the GPU instead of one SM (r01's one-CTA-per-object kernel: 185 us at T = 512).
// Sums follow the canonical order documented in oracle/restatement.py (one accumulator per lane
// over elements 128k + 4*lane + j, xor-butterfly across lanes, no FMA contraction; run sums in
// ascending token order), so sims, cuts and merged tokens are bit-identical to the oracle.
//
// Roofline: L2 (pooled rows were just written by kernel 2); reads ~3 T C 4 bytes per object.
#include "common.cuh"

#include <cstdlib>

// Developer build (-DUFV_TTM_TRACE): CTA 0 of the fused kernel stamps %globaltimer at its phase boundaries
// into a device array that tools/ttm_trace.py reads back.  Not compiled into the shipped library.
#ifdef UFV_TTM_TRACE
__device__ unsigned long long g_ttm_trace[16];
#define UFV_TRACE(i)                                                                  \
  do {                                                                                \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                        \
      unsigned long long t_;                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
      g_ttm_trace[i] = t_;                                                            \
    }                                                                                 \
  } while (0)
extern "C" int ufv_debug_ttm_trace(unsigned long long* out_host) {
  return int(cudaMemcpyFromSymbol(out_host, g_ttm_trace, sizeof(g_ttm_trace)));
}
#else
#define UFV_TRACE(i) do { } while (0)
#endif

namespace ufv {

constexpr int kSimWarps = 4;                 // adjacent pairs per CTA of the similarity kernel
constexpr int kMergeThreads = 288;           // 1152 channels / 4 per thread
constexpr int kTtmBatch = 5;                 // float4 loads in flight per lane per row

__device__ __forceinline__ float butterfly_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// topk order: NaN ranks above every number
__device__ __forceinline__ bool ranks_above(float a, float b) {
  return a > b || (a != a && b == b);
}

// ---- exact quotients with the reciprocal hoisted ------------------------------------------------------------
// F.normalize divides every element by the row norm (true division).  The IEEE quotient RN(a / m) costs ~10
// instructions apiece, and the similarity phase needs 2 x 1152 of them per pair.  With r = RN(1 / m) computed once
// per row,  q = RN(a r);  e = fma(-q, m, a) (the exact remainder);  q' = fma(e, r, q)  IS RN(a / m) as long as
// nothing underflows: checked bit for bit on 8e10 pairs with 2^-60 <= |a| < 2^41 and 2^-40 <= m < 2^47, random and
// all-ones / all-zeros / +-1-ulp significands (tools/div_probe.cu).  Whether a row qualifies is found while its norm
// is computed (one integer max per element); rows holding a zero, a denormal, a huge value, an inf or a NaN take the
// IEEE division.  A row's norm is kept with the flag in its sign: +norm = every element in range.
constexpr uint32_t kFastLo = 67u << 23;                          // 2^-60
constexpr uint32_t kFastSpan = (168u << 23) - 1u - kFastLo;      // up to the largest value below 2^41
__device__ __forceinline__ uint32_t range_excess(float x) { return (__float_as_uint(x) & 0x7fffffffu) - kFastLo; }
__device__ __forceinline__ float quot_hoisted(float a, float m, float r) {
  const float q = __fmul_rn(a, r);
  return __fmaf_rn(__fmaf_rn(-q, m, a), r, q);
}

// canonical 32-lane strided sum of squares of one row (oracle/restatement.py::_rowsum); returns
// max(||x||, 1e-12) (F.normalize), negated when some element is outside the hoisted-reciprocal range
__device__ __forceinline__ float row_norm(const float* __restrict__ row, int c, int lane) {
  float acc = 0.f;
  uint32_t worst = 0u;
  for (int e0 = lane * 4; e0 < c; e0 += 128 * kTtmBatch) {
    float4 v[kTtmBatch];
#pragma unroll
    for (int u = 0; u < kTtmBatch; ++u)
      if (e0 + u * 128 < c) v[u] = *reinterpret_cast<const float4*>(row + e0 + u * 128);
#pragma unroll
    for (int u = 0; u < kTtmBatch; ++u) {
      if (e0 + u * 128 < c) {
        acc = __fadd_rn(acc, __fmul_rn(v[u].x, v[u].x));
        acc = __fadd_rn(acc, __fmul_rn(v[u].y, v[u].y));
        acc = __fadd_rn(acc, __fmul_rn(v[u].z, v[u].z));
        acc = __fadd_rn(acc, __fmul_rn(v[u].w, v[u].w));
        worst = max(max(worst, range_excess(v[u].x)), max(range_excess(v[u].y), max(range_excess(v[u].z), range_excess(v[u].w))));
      }
    }
  }
  const float norm = fmaxf(__fsqrt_rn(butterfly_sum(acc)), 1e-12f);
  return __all_sync(0xffffffffu, worst <= kFastSpan) ? norm : -norm;
}

// <x_a / |x_a|, x_b / |x_b|> of two rows in the canonical order; sa / sb = their signed norms (row_norm)
__device__ __forceinline__ float pair_dot(const float* __restrict__ ra, const float* __restrict__ rb, float sa,
                                          float sb, int c, int lane) {
  const float ma = fabsf(sa), mb = fabsf(sb);
  float acc = 0.f;
  if (sa > 0.f && sb > 0.f) {                                    // warp-uniform
    const float ia = __frcp_rn(ma), ib = __frcp_rn(mb);
    for (int e0 = lane * 4; e0 < c; e0 += 128 * kTtmBatch) {
      float4 a[kTtmBatch], b[kTtmBatch];
#pragma unroll
      for (int u = 0; u < kTtmBatch; ++u) {
        if (e0 + u * 128 < c) {
          a[u] = *reinterpret_cast<const float4*>(ra + e0 + u * 128);
          b[u] = *reinterpret_cast<const float4*>(rb + e0 + u * 128);
        }
      }
#pragma unroll
      for (int u = 0; u < kTtmBatch; ++u) {
        if (e0 + u * 128 < c) {
          acc = __fadd_rn(acc, __fmul_rn(quot_hoisted(a[u].x, ma, ia), quot_hoisted(b[u].x, mb, ib)));
          acc = __fadd_rn(acc, __fmul_rn(quot_hoisted(a[u].y, ma, ia), quot_hoisted(b[u].y, mb, ib)));
          acc = __fadd_rn(acc, __fmul_rn(quot_hoisted(a[u].z, ma, ia), quot_hoisted(b[u].z, mb, ib)));
          acc = __fadd_rn(acc, __fmul_rn(quot_hoisted(a[u].w, ma, ia), quot_hoisted(b[u].w, mb, ib)));
        }
      }
    }
  } else {
    for (int e0 = lane * 4; e0 < c; e0 += 128 * kTtmBatch) {
      float4 a[kTtmBatch], b[kTtmBatch];
#pragma unroll
      for (int u = 0; u < kTtmBatch; ++u) {
        if (e0 + u * 128 < c) {
          a[u] = *reinterpret_cast<const float4*>(ra + e0 + u * 128);
          b[u] = *reinterpret_cast<const float4*>(rb + e0 + u * 128);
        }
      }
#pragma unroll
      for (int u = 0; u < kTtmBatch; ++u) {
        if (e0 + u * 128 < c) {
          acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(a[u].x, ma), __fdiv_rn(b[u].x, mb)));
          acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(a[u].y, ma), __fdiv_rn(b[u].y, mb)));
          acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(a[u].z, ma), __fdiv_rn(b[u].z, mb)));
          acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(a[u].w, ma), __fdiv_rn(b[u].w, mb)));
        }
      }
    }
  }
  return butterfly_sum(acc);
}

// ---- kernel 3a: adjacent cosine similarities, one warp per (object, pair) --------------------------
// Every pair of every object runs in parallel (a 512-frame object is 511 warps, not one CTA); each
// warp derives both norms itself, so there is no norm pass and no synchronisation.
__global__ void __launch_bounds__(32 * kSimWarps)
ttm_sims_kernel(const float* __restrict__ pooled, int c, const int32_t* __restrict__ obj_start,
                const int32_t* __restrict__ obj_len, int k_keep, float* __restrict__ sims, int sims_pitch) {
  pdl_wait();                  // pooled rows come from kernel 2
  pdl_launch_dependents();
  const int o = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * kSimWarps + (threadIdx.x >> 5);
  const int t_len = obj_len[o];
  if (t_len <= k_keep || i >= t_len - 1) return;     // layer.py:115: nothing to merge / no such pair
  const float* ra = pooled + (size_t(obj_start[o]) + i) * c;
  const float* rb = ra + c;
  const float acc = pair_dot(ra, rb, row_norm(ra, c, lane), row_norm(rb, c, lane), c, lane);   // L1 hits the second time
  if (lane == 0) sims[size_t(o) * sims_pitch + i] = acc;
}

// Token count of object o: to the device array and, when the caller asked for it, straight into
// pinned host memory as ONE self-describing word, (epoch << 16) | count.  The host polls until every
// word of the call carries the call's epoch: no fence, ticket, copy, event or stream synchronisation
// sits between the merge decisions and the reference's list[int].
__device__ __forceinline__ void publish_count(int32_t* counts_out, int32_t* counts_host, int32_t epoch,
                                              int o, int count) {
  counts_out[o] = count;
  if (counts_host != nullptr)
    *reinterpret_cast<volatile int32_t*>(counts_host + o) = (epoch << 16) | (count & 0xffff);
}

template <typename T>
__device__ __forceinline__ void store_token(T* tokens_out, float* tokens_f32_out, size_t dst, float4 v) {
  if (tokens_f32_out != nullptr) *reinterpret_cast<float4*>(tokens_f32_out + dst) = v;
  tokens_out[dst + 0] = Elem<T>::from_f32(v.x);
  tokens_out[dst + 1] = Elem<T>::from_f32(v.y);
  tokens_out[dst + 2] = Elem<T>::from_f32(v.z);
  tokens_out[dst + 3] = Elem<T>::from_f32(v.w);
}
template <>
__device__ __forceinline__ void store_token<float>(float* tokens_out, float* tokens_f32_out, size_t dst, float4 v) {
  if (tokens_f32_out != nullptr) *reinterpret_cast<float4*>(tokens_f32_out + dst) = v;
  *reinterpret_cast<float4*>(tokens_out + dst) = v;
}
template <>
__device__ __forceinline__ void store_token<__nv_bfloat16>(__nv_bfloat16* tokens_out, float* tokens_f32_out,
                                                           size_t dst, float4 v) {
  if (tokens_f32_out != nullptr) *reinterpret_cast<float4*>(tokens_f32_out + dst) = v;
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(tokens_out + dst) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
template <>
__device__ __forceinline__ void store_token<__half>(__half* tokens_out, float* tokens_f32_out, size_t dst,
                                                    float4 v) {
  if (tokens_f32_out != nullptr) *reinterpret_cast<float4*>(tokens_f32_out + dst) = v;
  const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(tokens_out + dst) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

// ---- kernel 3b: threshold, cuts and ONE merged token per CTA (object, slot) -------------------------
//   threshold   kth = r-th largest s (duplicates counted), r = T - K   (torch.topk, :17-18)
//   cuts        cut after token i  <=>  s_i < kth (strict)             (:24)
//   merge       mean of the slot's run, ascending token order          (:26,31)
// Every slot CTA of an object repeats the (cheap) selection; slot 0 also publishes the count.
template <typename T>
__global__ void __launch_bounds__(kMergeThreads)
ttm_merge_kernel(const float* __restrict__ pooled, int c, const int32_t* __restrict__ obj_start,
                 const int32_t* __restrict__ obj_len, const int32_t* __restrict__ slot_off, int k_keep,
                 int max_len, const float* __restrict__ sims, int sims_pitch, T* __restrict__ tokens_out,
                 float* __restrict__ tokens_f32_out, int32_t* __restrict__ counts_out,
                 uint32_t* __restrict__ cuts_out, int cut_pitch_words, int32_t* __restrict__ counts_host,
                 int32_t epoch, const ufv_dyn_args* __restrict__ dyn) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  const int len_words = (max_len + 31) / 32;
  float* s_sim = reinterpret_cast<float*>(dyn_smem);                 // [max_len]
  uint32_t* s_cutw = reinterpret_cast<uint32_t*>(s_sim + max_len);   // [len_words]
  int32_t* s_wpre = reinterpret_cast<int32_t*>(s_cutw + len_words);  // [len_words + 1]
  __shared__ float s_kth;
  __shared__ int s_first, s_last;

  const int o = blockIdx.y, g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();                  // similarities come from kernel 3a, pooled rows from kernel 2
  pdl_launch_dependents();
  if (dyn != nullptr) {        // graph replay: per-call values come through the device block
    epoch = dyn->epoch;
    if (dyn->counts_out != 0) counts_out = reinterpret_cast<int32_t*>(dyn->counts_out);
  }
  const int t_len = obj_len[o];
  const int n_slots = min(t_len, k_keep);
  if (g >= n_slots) {          // this object reserved fewer than K slots (T_o < K)
    if (g == 0 && tid == 0) publish_count(counts_out, counts_host, epoch, o, 0);
    return;
  }
  const float* x = pooled + size_t(obj_start[o]) * c;
  const int c4 = c >> 2;
  const size_t out_row = size_t(slot_off[o] + g) * c;

  if (g == 0 && cuts_out != nullptr)
    for (int w = tid; w < cut_pitch_words; w += kMergeThreads) cuts_out[size_t(o) * cut_pitch_words + w] = 0u;

  if (t_len <= k_keep) {       // layer.py:115: nothing to merge, token g passes through
    for (int q = tid; q < c4; q += kMergeThreads)
      store_token<T>(tokens_out, tokens_f32_out, out_row + q * 4,
                     *reinterpret_cast<const float4*>(x + size_t(g) * c + q * 4));
    if (g == 0 && tid == 0) publish_count(counts_out, counts_host, epoch, o, t_len);
    return;
  }

  // ---- r-th largest similarity by rank counting ------------------------------------------------------
  const int n_sim = t_len - 1;
  for (int i = tid; i < n_sim; i += kMergeThreads) s_sim[i] = sims[size_t(o) * sims_pitch + i];
  __syncthreads();
  const int r = t_len - k_keep;
  for (int i = tid; i < n_sim; i += kMergeThreads) {
    const float si = s_sim[i];
    int above = 0, not_below = 0;
    for (int j = 0; j < n_sim; ++j) {
      const float sj = s_sim[j];          // broadcast read
      above += ranks_above(sj, si);
      not_below += !ranks_above(si, sj);
    }
    if (above < r && r <= not_below) s_kth = si;   // every writer holds an equal value
  }
  __syncthreads();

  // ---- cuts -> bounds of run g ---------------------------------------------------------------------------
  const float kth = s_kth;
  for (int base = 0; base < len_words * 32; base += 32 * (kMergeThreads / 32)) {
    const int i = base + tid;
    const bool cut = i < n_sim && s_sim[i] < kth;
    const uint32_t word = __ballot_sync(0xffffffffu, cut);
    if (lane == 0 && (i >> 5) < len_words) s_cutw[i >> 5] = word;
  }
  __syncthreads();
  if (warp == 0) {   // exclusive prefix over the cut words
    int carry = 0;
    for (int w0 = 0; w0 < len_words; w0 += 32) {
      const int w = w0 + lane;
      const int mine = w < len_words ? __popc(s_cutw[w]) : 0;
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      if (w < len_words) s_wpre[w] = carry + incl - mine;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_wpre[len_words] = carry;
  }
  if (tid == 0) {
    s_first = 0;
    s_last = t_len - 1;        // the last run always ends at the last token (:29-31)
  }
  __syncthreads();
  const int n_cut = s_wpre[len_words];
  const int count = n_cut + 1;            // <= k_keep because at least r sims are >= kth
  for (int i = tid; i < n_sim; i += kMergeThreads) {   // cut number g - 1 opens run g, cut number g closes it
    const uint32_t word = s_cutw[i >> 5];
    const uint32_t bit = 1u << (i & 31);
    if (word & bit) {
      const int ord = s_wpre[i >> 5] + __popc(word & (bit - 1u));
      if (ord == g - 1) s_first = i + 1;
      if (ord == g) s_last = i;
    }
  }
  if (g == 0) {
    if (tid == 0) publish_count(counts_out, counts_host, epoch, o, count);
    if (cuts_out != nullptr)
      for (int w = tid; w < min(len_words, cut_pitch_words); w += kMergeThreads)
        cuts_out[size_t(o) * cut_pitch_words + w] = s_cutw[w];
  }
  __syncthreads();

  // ---- mean of run g, ascending token order; slots past the count are zero-filled -----------------------
  const int first = s_first, last = s_last;
  for (int q = tid; q < c4; q += kMergeThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < count) {
      const float* src = x + size_t(first) * c + q * 4;
#pragma unroll 8
      for (int t = first; t <= last; ++t, src += c) {
        const float4 v = *reinterpret_cast<const float4*>(src);
        acc.x = __fadd_rn(acc.x, v.x);
        acc.y = __fadd_rn(acc.y, v.y);
        acc.z = __fadd_rn(acc.z, v.z);
        acc.w = __fadd_rn(acc.w, v.w);
      }
      const int len = last - first + 1;
      const float n = float(len);
      if ((len & (len - 1)) == 0) {         // 1, 2, 4, ...: scaling by 2^-k rounds exactly like the division
        const float inv = __frcp_rn(n);
        acc.x = __fmul_rn(acc.x, inv);
        acc.y = __fmul_rn(acc.y, inv);
        acc.z = __fmul_rn(acc.z, inv);
        acc.w = __fmul_rn(acc.w, inv);
      } else {
        acc.x = __fdiv_rn(acc.x, n);
        acc.y = __fdiv_rn(acc.y, n);
        acc.z = __fdiv_rn(acc.z, n);
        acc.w = __fdiv_rn(acc.w, n);
      }
    }
    store_token<T>(tokens_out, tokens_f32_out, out_row + q * 4, acc);
  }
}

// ---- fused variant for short objects (max_len <= kFusedMaxLen): one CTA per object does 3a and 3b ------
// With T = 16 the whole merge is a few microseconds of latency; one launch and one CTA per object
// beat two launches (measured at 32 objects x 16 frames: 11.7 us against 8.2 + 15.7 us).
constexpr int kTtmThreads = 512;
constexpr int kTtmWarps = kTtmThreads / 32;
constexpr int kFusedMaxLen = 64;
constexpr size_t kTtmMaxSmem = 220 * 1024;   // staged rows + small arrays (227 KB per CTA on sm_100, minus static)

// The object's T pooled rows (T * C * 4 bytes, contiguous) are first pulled into shared memory with the TMA
// engine -- one bulk copy per row, all in flight at once, ONE L2 round trip -- and every later phase
// (norms, adjacent dots, run means: three dependent passes over the rows) reads them from there.  Reading
// them from L2 in every phase made this kernel a chain of ~600-cycle round trips on 32 of 148 SMs: 17 us
// inside the pipeline for a few microseconds of work (profiles/r02a_timeline_n1.txt).  `stage_rows` = 0
// keeps the global reads (objects too long for shared memory).  The arithmetic and its order are the same
// either way.
template <typename T>
__global__ void __launch_bounds__(kTtmThreads)
ttm_fused_kernel(const float* __restrict__ pooled, int c, const int32_t* __restrict__ obj_start,
           const int32_t* __restrict__ obj_len, const int32_t* __restrict__ slot_off, int k_keep,
           int max_len, T* __restrict__ tokens_out, float* __restrict__ tokens_f32_out,
           int32_t* __restrict__ counts_out, uint32_t* __restrict__ cuts_out, int cut_pitch_words,
           float* __restrict__ sims_out, int sims_pitch, int32_t* __restrict__ counts_host, int32_t epoch,
           const ufv_dyn_args* __restrict__ dyn, int stage_rows) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  const int len_words = (max_len + 31) / 32;
  float* s_norm = reinterpret_cast<float*>(dyn_smem);          // [max_len]
  float* s_sim = s_norm + max_len;                             // [max_len]
  uint32_t* s_cutw = reinterpret_cast<uint32_t*>(s_sim + max_len);   // [len_words]
  int32_t* s_wpre = reinterpret_cast<int32_t*>(s_cutw + len_words);  // [len_words + 1]
  int32_t* s_gend = s_wpre + len_words + 1;                    // [k_keep + 1]
  // staged rows start at the next 128-byte boundary after the small arrays (same formula on the host)
  float* s_rows = reinterpret_cast<float*>(dyn_smem + ((size_t(max_len) * 8 + size_t(len_words) * 8 + 4 +
                                                        size_t(k_keep + 1) * 4 + 127) & ~size_t(127)));
  __shared__ float s_kth;
  __shared__ __align__(8) uint64_t s_rows_bar;

  const int o = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the object's extent is plan data (uploaded before kernel 1 ran): fetched while the pool kernel drains
  UFV_TRACE(0);
  const int t_len = obj_len[o];
  const int slot = slot_off[o];
  const int start = obj_start[o];
  if (stage_rows && tid == 0) {
    mbar_init(&s_rows_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  UFV_TRACE(1);
  pdl_wait();                  // pooled rows come from kernel 2
  UFV_TRACE(2);
  pdl_launch_dependents();
  if (dyn != nullptr) {        // graph replay: per-call values come through the device block
    epoch = dyn->epoch;
    if (dyn->counts_out != 0) counts_out = reinterpret_cast<int32_t*>(dyn->counts_out);
  }
  const float* x = pooled + size_t(start) * c;
  const int c4 = c >> 2;
  const int n_slots = min(t_len, k_keep);
  const bool staged = stage_rows && t_len > k_keep;
  if (staged && warp == 0) {
    if (lane == 0) mbar_arrive_expect_tx(&s_rows_bar, uint32_t(t_len) * uint32_t(c) * 4u);
    __syncwarp();
    for (int t = lane; t < t_len; t += 32) bulk_g2s(s_rows + size_t(t) * c, x + size_t(t) * c, uint32_t(c) * 4u, &s_rows_bar);
  }

  if (cuts_out != nullptr)
    for (int w = tid; w < cut_pitch_words; w += kTtmThreads) cuts_out[size_t(o) * cut_pitch_words + w] = 0u;

  if (t_len <= k_keep) {   // layer.py:115: nothing to merge, tokens pass through
    for (int u = tid; u < t_len * c4; u += kTtmThreads) {
      const int t = u / c4, q = u - t * c4;
      const float4 v = *reinterpret_cast<const float4*>(x + size_t(t) * c + q * 4);
      store_token<T>(tokens_out, tokens_f32_out, size_t(slot + t) * c + q * 4, v);
    }
    if (tid == 0) publish_count(counts_out, counts_host, epoch, o, t_len);
    return;
  }
  if (staged) {                 // from here on the rows are read from shared memory
    mbar_wait(&s_rows_bar, 0);
    x = s_rows;
  }
  UFV_TRACE(3);

  // ---- 1. norms: one warp per token (kTtmBatch float4 loads in flight per lane) -----------------
  for (int t = warp; t < t_len; t += kTtmWarps) {
    const float sn = row_norm(x + size_t(t) * c, c, lane);       // signed: see row_norm
    if (lane == 0) s_norm[t] = sn;
  }
  __syncthreads();

  UFV_TRACE(4);
  // ---- 2. adjacent cosine similarities: one warp per pair ----------------------------------------
  const int n_sim = t_len - 1;
  for (int i = warp; i < n_sim; i += kTtmWarps) {
    const float* ra = x + size_t(i) * c;
    const float acc = pair_dot(ra, ra + c, s_norm[i], s_norm[i + 1], c, lane);
    if (lane == 0) {
      s_sim[i] = acc;
      if (sims_out != nullptr) sims_out[size_t(o) * sims_pitch + i] = acc;
    }
  }
  __syncthreads();

  UFV_TRACE(5);
  const int r = t_len - k_keep;
  if (n_sim <= 32) {
    // ---- 3 + 4 for short objects: one warp, one similarity per lane, no block-wide barriers in between -----
    if (warp == 0) {
      const float si = lane < n_sim ? s_sim[lane] : 0.f;
      int above = 0, not_below = 0;
      for (int j = 0; j < n_sim; ++j) {
        const float sj = __shfl_sync(0xffffffffu, si, j);
        above += ranks_above(sj, si);
        not_below += !ranks_above(si, sj);
      }
      const uint32_t holders = __ballot_sync(0xffffffffu, lane < n_sim && above < r && r <= not_below);
      const float kth = __shfl_sync(0xffffffffu, si, __ffs(holders) - 1);   // every holder has an equal value
      const uint32_t word = __ballot_sync(0xffffffffu, lane < n_sim && si < kth);
      const int n_cut = __popc(word);
      for (int w = lane; w < len_words; w += 32) {
        s_cutw[w] = w == 0 ? word : 0u;
        s_wpre[w] = w == 0 ? 0 : n_cut;
      }
      if ((word >> lane) & 1u) s_gend[__popc(word & ((1u << lane) - 1u))] = lane;
      if (lane == 0) {
        s_wpre[len_words] = n_cut;
        s_gend[n_cut] = t_len - 1;          // the last run always ends at the last token (:29-31)
        publish_count(counts_out, counts_host, epoch, o, n_cut + 1);
      }
      if (cuts_out != nullptr)
        for (int w = lane; w < min(len_words, cut_pitch_words); w += 32)
          cuts_out[size_t(o) * cut_pitch_words + w] = w == 0 ? word : 0u;
    }
    __syncthreads();
  } else {
    // ---- 3. r-th largest by rank counting -------------------------------------------------------------
    for (int i = tid; i < n_sim; i += kTtmThreads) {
      const float si = s_sim[i];
      int above = 0, not_below = 0;
      for (int j = 0; j < n_sim; ++j) {
        const float sj = s_sim[j];          // broadcast read
        above += ranks_above(sj, si);
        not_below += !ranks_above(si, sj);
      }
      if (above < r && r <= not_below) s_kth = si;   // every writer holds an equal value
    }
    __syncthreads();

    // ---- 4. cuts -> run ends ------------------------------------------------------------------------------
    const float kth = s_kth;
    for (int base = 0; base < len_words * 32; base += kTtmThreads) {
      const int i = base + tid;
      const bool cut = i < n_sim && s_sim[i] < kth;
      const uint32_t word = __ballot_sync(0xffffffffu, cut);
      if (lane == 0 && (i >> 5) < len_words) s_cutw[i >> 5] = word;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive prefix over the cut words
      int carry = 0;
      for (int w0 = 0; w0 < len_words; w0 += 32) {
        const int w = w0 + lane;
        const int mine = w < len_words ? __popc(s_cutw[w]) : 0;
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += up;
        }
        if (w < len_words) s_wpre[w] = carry + incl - mine;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) s_wpre[len_words] = carry;
    }
    __syncthreads();
    const int n_cut = s_wpre[len_words];
    for (int i = tid; i < n_sim; i += kTtmThreads) {
      const uint32_t word = s_cutw[i >> 5];
      const uint32_t bit = 1u << (i & 31);
      if (word & bit) s_gend[s_wpre[i >> 5] + __popc(word & (bit - 1u))] = i;
    }
    if (tid == 0) {
      s_gend[n_cut] = t_len - 1;            // the last run always ends at the last token (:29-31)
      publish_count(counts_out, counts_host, epoch, o, n_cut + 1);
    }
    if (cuts_out != nullptr)
      for (int w = tid; w < min(len_words, cut_pitch_words); w += kTtmThreads)
        cuts_out[size_t(o) * cut_pitch_words + w] = s_cutw[w];
    __syncthreads();
  }
  const int count = s_wpre[len_words] + 1;  // <= k_keep because at least r sims are >= kth
  UFV_TRACE(6);
  // ---- 5. run means, ascending token order; unused slots are zero-filled -----------------------------------
  // one warp per (output token, 128 channels): the run bounds are warp-uniform (broadcast reads, no divergence in
  // the row loop), and no per-thread integer division
  const int chunks = (c4 + 31) >> 5;
  for (int item = warp; item < n_slots * chunks; item += kTtmWarps) {
    const int g = item / chunks, q = (item - g * chunks) * 32 + lane;
    if (q >= c4) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < count) {
      const int first = g == 0 ? 0 : s_gend[g - 1] + 1;
      const int last = s_gend[g];
      const float* src = x + size_t(first) * c + q * 4;
#pragma unroll 4
      for (int t = first; t <= last; ++t, src += c) {
        const float4 v = *reinterpret_cast<const float4*>(src);
        acc.x = __fadd_rn(acc.x, v.x);
        acc.y = __fadd_rn(acc.y, v.y);
        acc.z = __fadd_rn(acc.z, v.z);
        acc.w = __fadd_rn(acc.w, v.w);
      }
      const int len = last - first + 1;
      const float n = float(len);
      if ((len & (len - 1)) == 0) {         // 1, 2, 4, ...: scaling by 2^-k rounds exactly like the division
        const float inv = __frcp_rn(n);
        acc.x = __fmul_rn(acc.x, inv);
        acc.y = __fmul_rn(acc.y, inv);
        acc.z = __fmul_rn(acc.z, inv);
        acc.w = __fmul_rn(acc.w, inv);
      } else {
        acc.x = __fdiv_rn(acc.x, n);
        acc.y = __fdiv_rn(acc.y, n);
        acc.z = __fdiv_rn(acc.z, n);
        acc.w = __fdiv_rn(acc.w, n);
      }
    }
    store_token<T>(tokens_out, tokens_f32_out, size_t(slot + g) * c + q * 4, acc);
  }
  UFV_TRACE(7);
}


template <typename T>
static int launch_ttm(const float* pooled, int c, const int32_t* obj_start, const int32_t* obj_len,
                      const int32_t* slot_off, int n_obj, int max_len, int k_keep, void* tokens_out,
                      float* tokens_f32_out, int32_t* counts_out, uint32_t* cuts_out,
                      int cut_pitch_words, float* sims, int sims_pitch, int32_t* counts_host, int32_t epoch,
                      const ufv_dyn_args* dyn, cudaStream_t stream) {
  const int len_words = (max_len + 31) / 32;
  if (max_len <= kFusedMaxLen) {
    const size_t small = (size_t(max_len) * 8 + size_t(len_words) * 8 + 4 + size_t(k_keep + 1) * 4 + 127) & ~size_t(127);
    const size_t rows = size_t(max_len) * c * sizeof(float);
    static const bool no_stage = getenv("UFV_TTM_NO_STAGE") != nullptr;      // developer A/B knob
    const int stage_rows = !no_stage && c % 4 == 0 && small + rows <= kTtmMaxSmem ? 1 : 0;
    const size_t smem = small + (stage_rows ? rows : 0);
    auto kernel = ttm_fused_kernel<T>;
    static size_t configured = 48 * 1024;     // the attribute only ever grows; a benign race sets it twice
    if (smem > configured) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTtmMaxSmem));
      configured = kTtmMaxSmem;
    }
    return check_launch("ufv_ttm (fused)",
                        launch_kernel(kernel, dim3(n_obj), dim3(kTtmThreads), smem, stream, pooled, c, obj_start,
                                      obj_len, slot_off, k_keep, max_len, static_cast<T*>(tokens_out),
                                      tokens_f32_out, counts_out, cuts_out, cut_pitch_words, sims, sims_pitch,
                                      counts_host, epoch, dyn, stage_rows));
  }
  if (max_len > k_keep) {
    const dim3 grid((max_len - 1 + kSimWarps - 1) / kSimWarps, n_obj);
    const int rc = check_launch("ufv_ttm (similarities)",
                                launch_kernel(ttm_sims_kernel, grid, dim3(32 * kSimWarps), 0, stream, pooled, c,
                                              obj_start, obj_len, k_keep, sims, sims_pitch));
    if (rc != 0) return rc;
  }
  const size_t smem = size_t(max_len) * 4 + size_t(len_words) * 4 + size_t(len_words + 1) * 4;
  auto kernel = ttm_merge_kernel<T>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  const dim3 grid(max_len < k_keep ? max_len : k_keep, n_obj);
  return check_launch("ufv_ttm (merge)",
                      launch_kernel(kernel, grid, dim3(kMergeThreads), smem, stream, pooled, c, obj_start,
                                    obj_len, slot_off, k_keep, max_len, static_cast<const float*>(sims),
                                    sims_pitch, static_cast<T*>(tokens_out), tokens_f32_out, counts_out,
                                    cuts_out, cut_pitch_words, counts_host, epoch, dyn));
}

}  // namespace ufv

namespace ufv {
int ttm_dispatch(const float* pooled, int c, const int32_t* obj_start, const int32_t* obj_len,
                 const int32_t* slot_off, int n_obj, int max_len, int k_keep, void* tokens_out,
                 int out_dtype, float* tokens_f32_out, int32_t* counts_out, uint32_t* cuts_out,
                 int cut_pitch_words, float* sims_out, int sims_pitch, int32_t* counts_host,
                 int32_t epoch, const ufv_dyn_args* dyn, void* stream) {
  UFV_REQUIRE(n_obj >= 0, UFV_E_SHAPE, "ufv_ttm: n_obj=%d", n_obj);
  if (n_obj == 0) return 0;
  UFV_REQUIRE(pooled && obj_start && obj_len && slot_off && tokens_out && counts_out, UFV_E_NULL,
              "ufv_ttm: null pointer");
  UFV_REQUIRE(c >= 4 && c % 4 == 0, UFV_E_SHAPE, "ufv_ttm: c=%d must be a multiple of 4", c);
  UFV_REQUIRE(k_keep >= 1 && k_keep <= 4096, UFV_E_SHAPE, "ufv_ttm: k_keep=%d out of range", k_keep);
  UFV_REQUIRE(max_len >= 1 && max_len <= 16384, UFV_E_SHAPE, "ufv_ttm: max_len=%d out of range", max_len);
  UFV_REQUIRE(aligned16(pooled) && aligned16(tokens_out) && aligned16(tokens_f32_out), UFV_E_ALIGN,
              "ufv_ttm: buffers must be 16-byte aligned");
  UFV_REQUIRE(cuts_out == nullptr || cut_pitch_words >= (max_len + 31) / 32, UFV_E_SHAPE,
              "ufv_ttm: cut_pitch_words too small");
  UFV_REQUIRE(max_len <= k_keep || sims_out != nullptr, UFV_E_NULL,
              "ufv_ttm: sims_out (fp32 [n_obj * sims_pitch] scratch) is required when an object has more "
              "than k_keep frames");
  UFV_REQUIRE(sims_out == nullptr || sims_pitch >= max_len - 1, UFV_E_SHAPE, "ufv_ttm: sims_pitch too small");
  UFV_REQUIRE(counts_host == nullptr || dyn != nullptr || (epoch > 0 && epoch < 32768), UFV_E_SHAPE,
              "ufv_ttm: epoch %d must be in [1, 32767] when counts_host is given", epoch);
  UFV_REQUIRE(counts_host == nullptr || k_keep < 65536, UFV_E_SHAPE, "ufv_ttm: k_keep too large for the host word");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (out_dtype) {
    case UFV_F32:
      return launch_ttm<float>(pooled, c, obj_start, obj_len, slot_off, n_obj, max_len, k_keep, tokens_out,
                               tokens_f32_out, counts_out, cuts_out, cut_pitch_words, sims_out, sims_pitch, counts_host, epoch, dyn, st);
    case UFV_BF16:
      return launch_ttm<__nv_bfloat16>(pooled, c, obj_start, obj_len, slot_off, n_obj, max_len, k_keep,
                                       tokens_out, tokens_f32_out, counts_out, cuts_out, cut_pitch_words,
                                       sims_out, sims_pitch, counts_host, epoch, dyn, st);
    case UFV_F16:
      return launch_ttm<__half>(pooled, c, obj_start, obj_len, slot_off, n_obj, max_len, k_keep, tokens_out,
                                tokens_f32_out, counts_out, cuts_out, cut_pitch_words, sims_out, sims_pitch, counts_host, epoch, dyn, st);
    default:
      return fail(UFV_E_DTYPE, "ufv_ttm: unsupported output dtype %d", out_dtype);
  }
}
}  // namespace ufv

extern "C" int ufv_ttm(const float* pooled, int c, const int32_t* obj_start, const int32_t* obj_len,
                       const int32_t* slot_off, int n_obj, int max_len, int k_keep, void* tokens_out,
                       int out_dtype, float* tokens_f32_out, int32_t* counts_out, uint32_t* cuts_out,
                       int cut_pitch_words, float* sims_out, int sims_pitch, int32_t* counts_host,
                       int32_t epoch, void* stream) {
  return ufv::ttm_dispatch(pooled, c, obj_start, obj_len, slot_off, n_obj, max_len, k_keep, tokens_out, out_dtype,
                           tokens_f32_out, counts_out, cuts_out, cut_pitch_words, sims_out, sims_pitch, counts_host,
                           epoch, nullptr, stream);
}
