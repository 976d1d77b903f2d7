// Kernel 3: fused temporal token merge (TTM).
//
// Replaces token_merge (reference ufvideo/model/layer.py:6-33), its per-object dispatch
// (layer.py:110-119) and the downcast to the model dtype (layer.py:123).  On a GPU the reference
// pays one device->host sync per frame per object (the python `if sim[0,i] < kth`) plus O(T)
// tiny launches; here one CTA per object does everything on chip:
//   1. norms       m_t = max(sqrt(sum x_t^2), 1e-12)                      (F.normalize, :13-14)
//   2. sims        s_i = sum (x_i / m_i) * (x_{i+1} / m_{i+1})            (:15)
//   3. threshold   kth = r-th largest s (duplicates counted), r = T - K   (torch.topk, :17-18)
//   4. cuts        cut after token i  <=>  s_i < kth (strict)             (:24)
//   5. merge       mean of every maximal run, in temporal order           (:26,31)
// Sums follow the canonical order documented in oracle/restatement.py (one accumulator per lane
// over elements 128k + 4*lane + j, xor-butterfly across lanes, no FMA contraction; run sums in
// ascending token order), so sims, cuts and merged tokens are bit-identical to the oracle.
//
// Roofline: HBM/L2 (reads T*C*4 bytes per object twice from L2, writes <= K*C tokens).
#include "common.cuh"

namespace ufv {

constexpr int kTtmThreads = 512;
constexpr int kTtmWarps = kTtmThreads / 32;
constexpr int kTtmBatch = 5;   // float4 loads in flight per lane while a row is reduced

__device__ __forceinline__ float butterfly_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// topk order: NaN ranks above every number
__device__ __forceinline__ bool ranks_above(float a, float b) {
  return a > b || (a != a && b == b);
}

// Token count of object o: to the device array and, when the caller asked for it, straight into
// pinned host memory.  The CTA that publishes last stamps `epoch` behind the counts, which is what
// the host polls: the reference's list[int] is ready as soon as the merge decisions are, with no
// copy, event or stream synchronisation in between.
__device__ __forceinline__ void publish_count(int32_t* counts_out, int32_t* counts_host,
                                              uint32_t* done_ticket, int32_t epoch, int o, int count) {
  counts_out[o] = count;
  if (counts_host == nullptr) return;
  counts_host[o] = count;
  __threadfence_system();
  if (atomicAdd(done_ticket, 1u) == gridDim.x - 1) {
    *done_ticket = 0u;                    // self-reset for the next call
    __threadfence_system();
    *reinterpret_cast<volatile int32_t*>(counts_host + gridDim.x) = epoch;
  }
}

template <typename T>
__global__ void __launch_bounds__(kTtmThreads)
ttm_kernel(const float* __restrict__ pooled, int c, const int32_t* __restrict__ obj_start,
           const int32_t* __restrict__ obj_len, const int32_t* __restrict__ slot_off, int k_keep,
           int max_len, T* __restrict__ tokens_out, float* __restrict__ tokens_f32_out,
           int32_t* __restrict__ counts_out, uint32_t* __restrict__ cuts_out, int cut_pitch_words,
           float* __restrict__ sims_out, int sims_pitch, int32_t* __restrict__ counts_host,
           uint32_t* __restrict__ done_ticket, int32_t epoch) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const int len_words = (max_len + 31) / 32;
  float* s_norm = reinterpret_cast<float*>(dyn_smem);          // [max_len]
  float* s_sim = s_norm + max_len;                             // [max_len]
  uint32_t* s_cutw = reinterpret_cast<uint32_t*>(s_sim + max_len);   // [len_words]
  int32_t* s_wpre = reinterpret_cast<int32_t*>(s_cutw + len_words);  // [len_words + 1]
  int32_t* s_gend = s_wpre + len_words + 1;                    // [k_keep + 1]
  __shared__ float s_kth;

  const int o = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();                  // pooled rows come from kernel 2
  pdl_launch_dependents();
  const int t_len = obj_len[o];
  const int slot = slot_off[o];
  const float* x = pooled + size_t(obj_start[o]) * c;
  const int c4 = c >> 2;
  const int n_slots = min(t_len, k_keep);

  if (cuts_out != nullptr)
    for (int w = tid; w < cut_pitch_words; w += kTtmThreads) cuts_out[size_t(o) * cut_pitch_words + w] = 0u;

  if (t_len <= k_keep) {   // layer.py:115: nothing to merge, tokens pass through
    for (int u = tid; u < t_len * c4; u += kTtmThreads) {
      const int t = u / c4, q = u - t * c4;
      const float4 v = *reinterpret_cast<const float4*>(x + size_t(t) * c + q * 4);
      const size_t dst = size_t(slot + t) * c + q * 4;
      if (tokens_f32_out != nullptr) *reinterpret_cast<float4*>(tokens_f32_out + dst) = v;
      tokens_out[dst + 0] = Elem<T>::from_f32(v.x);
      tokens_out[dst + 1] = Elem<T>::from_f32(v.y);
      tokens_out[dst + 2] = Elem<T>::from_f32(v.z);
      tokens_out[dst + 3] = Elem<T>::from_f32(v.w);
    }
    if (tid == 0) publish_count(counts_out, counts_host, done_ticket, epoch, o, t_len);
    return;
  }

  // ---- 1. norms: one warp per token (kTtmBatch float4 loads in flight per lane) -----------------
  for (int t = warp; t < t_len; t += kTtmWarps) {
    const float* row = x + size_t(t) * c;
    float acc = 0.f;
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
        }
      }
    }
    acc = butterfly_sum(acc);
    if (lane == 0) s_norm[t] = fmaxf(__fsqrt_rn(acc), 1e-12f);
  }
  __syncthreads();

  // ---- 2. adjacent cosine similarities: one warp per pair ----------------------------------------
  const int n_sim = t_len - 1;
  for (int i = warp; i < n_sim; i += kTtmWarps) {
    const float* ra = x + size_t(i) * c;
    const float* rb = ra + c;
    const float ma = s_norm[i], mb = s_norm[i + 1];
    float acc = 0.f;
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
    acc = butterfly_sum(acc);
    if (lane == 0) {
      s_sim[i] = acc;
      if (sims_out != nullptr) sims_out[size_t(o) * sims_pitch + i] = acc;
    }
  }
  __syncthreads();

  // ---- 3. r-th largest by rank counting -------------------------------------------------------------
  const int r = t_len - k_keep;
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
  const int count = n_cut + 1;            // <= k_keep because at least r sims are >= kth
  for (int i = tid; i < n_sim; i += kTtmThreads) {
    const uint32_t word = s_cutw[i >> 5];
    const uint32_t bit = 1u << (i & 31);
    if (word & bit) s_gend[s_wpre[i >> 5] + __popc(word & (bit - 1u))] = i;
  }
  if (tid == 0) {
    s_gend[n_cut] = t_len - 1;            // the last run always ends at the last token (:29-31)
    publish_count(counts_out, counts_host, done_ticket, epoch, o, count);
  }
  if (cuts_out != nullptr)
    for (int w = tid; w < min(len_words, cut_pitch_words); w += kTtmThreads)
      cuts_out[size_t(o) * cut_pitch_words + w] = s_cutw[w];
  __syncthreads();

  // ---- 5. run means, ascending token order; unused slots are zero-filled -----------------------------------
  for (int u = tid; u < n_slots * c4; u += kTtmThreads) {
    const int g = u / c4, q = u - g * c4;
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
      const float n = float(last - first + 1);
      acc.x = __fdiv_rn(acc.x, n);
      acc.y = __fdiv_rn(acc.y, n);
      acc.z = __fdiv_rn(acc.z, n);
      acc.w = __fdiv_rn(acc.w, n);
    }
    const size_t dst = size_t(slot + g) * c + q * 4;
    if (tokens_f32_out != nullptr) *reinterpret_cast<float4*>(tokens_f32_out + dst) = acc;
    tokens_out[dst + 0] = Elem<T>::from_f32(acc.x);
    tokens_out[dst + 1] = Elem<T>::from_f32(acc.y);
    tokens_out[dst + 2] = Elem<T>::from_f32(acc.z);
    tokens_out[dst + 3] = Elem<T>::from_f32(acc.w);
  }
}

template <typename T>
static int launch_ttm(const float* pooled, int c, const int32_t* obj_start, const int32_t* obj_len,
                      const int32_t* slot_off, int n_obj, int max_len, int k_keep, void* tokens_out,
                      float* tokens_f32_out, int32_t* counts_out, uint32_t* cuts_out,
                      int cut_pitch_words, float* sims_out, int sims_pitch, int32_t* counts_host,
                      uint32_t* done_ticket, int32_t epoch, cudaStream_t stream) {
  const int len_words = (max_len + 31) / 32;
  const size_t smem = size_t(max_len) * 8 + size_t(len_words) * 4 + size_t(len_words + 1) * 4 +
                      size_t(k_keep + 1) * 4;
  auto kernel = ttm_kernel<T>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  return check_launch("ufv_ttm",
                      launch_kernel(kernel, dim3(n_obj), dim3(kTtmThreads), smem, stream, pooled, c, obj_start,
                                    obj_len, slot_off, k_keep, max_len, static_cast<T*>(tokens_out),
                                    tokens_f32_out, counts_out, cuts_out, cut_pitch_words, sims_out,
                                    sims_pitch, counts_host, done_ticket, epoch));
}

}  // namespace ufv

extern "C" int ufv_ttm(const float* pooled, int c, const int32_t* obj_start, const int32_t* obj_len,
                       const int32_t* slot_off, int n_obj, int max_len, int k_keep, void* tokens_out,
                       int out_dtype, float* tokens_f32_out, int32_t* counts_out, uint32_t* cuts_out,
                       int cut_pitch_words, float* sims_out, int sims_pitch, int32_t* counts_host,
                       uint32_t* done_ticket, int32_t epoch, void* stream) {
  using namespace ufv;
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
  UFV_REQUIRE(sims_out == nullptr || sims_pitch >= max_len - 1, UFV_E_SHAPE, "ufv_ttm: sims_pitch too small");
  UFV_REQUIRE(counts_host == nullptr || done_ticket != nullptr, UFV_E_NULL,
              "ufv_ttm: counts_host needs done_ticket");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (out_dtype) {
    case UFV_F32:
      return launch_ttm<float>(pooled, c, obj_start, obj_len, slot_off, n_obj, max_len, k_keep, tokens_out,
                               tokens_f32_out, counts_out, cuts_out, cut_pitch_words, sims_out, sims_pitch, counts_host, done_ticket, epoch, st);
    case UFV_BF16:
      return launch_ttm<__nv_bfloat16>(pooled, c, obj_start, obj_len, slot_off, n_obj, max_len, k_keep,
                                       tokens_out, tokens_f32_out, counts_out, cuts_out, cut_pitch_words,
                                       sims_out, sims_pitch, counts_host, done_ticket, epoch, st);
    case UFV_F16:
      return launch_ttm<__half>(pooled, c, obj_start, obj_len, slot_off, n_obj, max_len, k_keep, tokens_out,
                                tokens_f32_out, counts_out, cuts_out, cut_pitch_words, sims_out, sims_pitch, counts_host, done_ticket, epoch, st);
    default:
      return fail(UFV_E_DTYPE, "ufv_ttm: unsupported output dtype %d", out_dtype);
  }
}
