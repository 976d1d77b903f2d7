// Error plumbing, the chained whole-path entry point and the row-gather used for compaction.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace ufv {

int launch_mask_to_patches(const ufv_mask_desc* desc, const int32_t* taps, int n_masks, int n_out, int any_row_mode,
                           uint32_t* bits_out, int32_t* cnt_out, uint16_t* idx_out, int idx_pitch,
                           const ufv_dyn_args* dyn_src, ufv_dyn_args* dyn_dev, void* stream);
int ttm_dispatch(const float* pooled, int c, const int32_t* obj_start, const int32_t* obj_len,
                 const int32_t* slot_off, int n_obj, int max_len, int k_keep, void* tokens_out,
                 int out_dtype, float* tokens_f32_out, int32_t* counts_out, uint32_t* cuts_out,
                 int cut_pitch_words, float* sims_out, int sims_pitch, int32_t* counts_host,
                 int32_t epoch, const ufv_dyn_args* dyn, void* stream);
int last_linear_dyn(const void* x, const void* w, const void* bias, int m, int n, int k, int dtype,
                    const ufv_peer_args* peer, const ufv_dyn_args* dyn, void* ws, int64_t ws_bytes,
                    const int32_t* row_map, void* stream);

static_assert(sizeof(ufv_dyn_args) == 256, "ufv_dyn_args must stay 256 bytes (kernel 1 copies 64 words)");

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static const bool on = getenv("UFV_NO_PDL") == nullptr;
  return on;
}

int check_launch(const char* what, cudaError_t launch_rc) {
  const cudaError_t e = launch_rc != cudaSuccess ? launch_rc : cudaGetLastError();
  if (launch_rc != cudaSuccess) cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return static_cast<int>(e);
}

__global__ void gather_rows_kernel(const uint4* __restrict__ in, const int32_t* __restrict__ row_map,
                                   uint4* __restrict__ out, int vec_per_row) {
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x;
  const uint4* src = in + size_t(row_map[r]) * vec_per_row;
  uint4* dst = out + size_t(r) * vec_per_row;
  for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) dst[i] = src[i];
}

// Object o owns rows [slot_off[o], slot_off[o] + counts[o]) of `in`; they move to rows
// [sum of counts before o, ...) of `out`.  One CTA per object; the offset is a block-wide sum.
__global__ void compact_rows_kernel(const uint4* __restrict__ in, const int32_t* __restrict__ slot_off,
                                    const int32_t* __restrict__ counts, uint4* __restrict__ out,
                                    int vec_per_row) {
  __shared__ int s_part[32];
  pdl_wait();
  pdl_launch_dependents();
  const int o = blockIdx.x;
  int before = 0;
  for (int i = threadIdx.x; i < o; i += blockDim.x) before += counts[i];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) before += __shfl_xor_sync(0xffffffffu, before, off);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = before;
  __syncthreads();
  before = 0;
  for (int w = 0; w < int(blockDim.x >> 5); ++w) before += s_part[w];
  const int n = counts[o] * vec_per_row;
  const uint4* src = in + size_t(slot_off[o]) * vec_per_row;
  uint4* dst = out + size_t(before) * vec_per_row;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// ---- <region> splice (the consumer of the path, videorefer_arch.py:300-311) ----------------------------
// The flattened token sequence has one placeholder row per object (region_pos ascending); the output is
// that sequence with placeholder o replaced by object o's counts[o] token rows.  One CTA per work item:
// items [0, n_text) are text rows, items [n_text, n_text + m_pad) are padded token rows.  Where a row
// lands depends on the counts of the objects before it, which live on the device -- so the caller never
// needs the host copy of region_token_nums to build inputs_embeds.
__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, int v) {   // first index with a[i] >= v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void splice_rows_kernel(const uint4* __restrict__ text, int n_text, const int32_t* __restrict__ region_pos,
                                   const uint4* __restrict__ tokens, const int32_t* __restrict__ slot_off,
                                   const int32_t* __restrict__ counts, int n_obj, int m_pad, uint4* __restrict__ out,
                                   int32_t* __restrict__ out_len, int32_t* __restrict__ row_src, int vec_per_row) {
  __shared__ int s_part[32];
  __shared__ int s_dst;
  pdl_wait();
  pdl_launch_dependents();
  const int item = blockIdx.x;
  const uint4* src;
  int n_before;          // objects whose placeholder precedes this row
  int extra = 0;         // offset inside the object's own tokens
  int tag;
  bool live = true;
  if (item < n_text) {
    n_before = lower_bound_i32(region_pos, n_obj, item);
    live = !(n_before < n_obj && region_pos[n_before] == item);   // placeholders themselves disappear
    src = text + size_t(item) * vec_per_row;
    tag = item;
  } else {
    const int r = item - n_text;
    int o = lower_bound_i32(slot_off, n_obj, r + 1) - 1;           // last object with slot_off[o] <= r
    live = o >= 0 && r - slot_off[o] < counts[o];
    n_before = max(o, 0);
    extra = o >= 0 ? r - slot_off[o] : 0;
    src = tokens + size_t(r) * vec_per_row;
    tag = -(r + 1);
  }
  // rows before this one: its own text position (or its placeholder's) + sum over earlier objects of (count - 1)
  int shift = 0;
  for (int i = threadIdx.x; i < n_before; i += blockDim.x) shift += counts[i] - 1;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) shift += __shfl_xor_sync(0xffffffffu, shift, off);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = shift;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) total += s_part[w];
    const int base = item < n_text ? item : region_pos[n_before];
    s_dst = base + total + extra;
    if (item == 0 && out_len != nullptr) {     // total length: every placeholder turns into counts[o] rows
      int len = n_text;
      for (int i = 0; i < n_obj; ++i) len += counts[i] - 1;
      *out_len = len;
    }
  }
  __syncthreads();
  if (!live) return;
  const int dst = s_dst;
  uint4* d = out + size_t(dst) * vec_per_row;
  for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) d[i] = src[i];
  if (threadIdx.x == 0 && row_src != nullptr) row_src[dst] = tag;
}

// ---- static <region> splice with batch padding and labels (videorefer_arch.py:291-368) ---------------------
// The batch's final [B, L_max] layout is known on the host whenever no object ties below its reserved token
// count (the normal case): src_map[i] says what output row i is -- a text row (>= 0: its index), padding
// (-1: zero embedding, IGNORE label, attention 0) or a region-token row (-2: the projector's scatter epilogue
// writes the embedding; label IGNORE, attention 1).  One CTA per output row.
__global__ void splice_static_kernel(const uint4* __restrict__ text, const int64_t* __restrict__ labels_in,
                                     const int32_t* __restrict__ src_map, uint4* __restrict__ out,
                                     int64_t* __restrict__ labels_out, uint8_t* __restrict__ attn_out,
                                     int vec_per_row, int64_t ignore_index) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = blockIdx.x;
  const int s = src_map[i];
  if (threadIdx.x == 0) {
    if (labels_out != nullptr) labels_out[i] = (s >= 0 && labels_in != nullptr) ? labels_in[s] : ignore_index;
    if (attn_out != nullptr) attn_out[i] = s == -1 ? 0 : 1;
  }
  if (s == -2) return;
  uint4* d = out + size_t(i) * vec_per_row;
  if (s >= 0) {
    const uint4* src = text + size_t(s) * vec_per_row;
    for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) d[v] = src[v];
  } else {
    for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) d[v] = make_uint4(0u, 0u, 0u, 0u);
  }
}

}  // namespace ufv

extern "C" int ufv_splice_static(const void* text, const int64_t* labels_in, const int32_t* src_map, void* out,
                                 int64_t* labels_out, uint8_t* attn_out, int n_out_rows, int row_bytes,
                                 int64_t ignore_index, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(n_out_rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0, UFV_E_SHAPE,
              "ufv_splice_static: n_out_rows=%d row_bytes=%d", n_out_rows, row_bytes);
  if (n_out_rows == 0) return 0;
  UFV_REQUIRE(src_map && out, UFV_E_NULL, "ufv_splice_static: null pointer");
  UFV_REQUIRE(aligned16(text) && aligned16(out), UFV_E_ALIGN, "ufv_splice_static: unaligned buffer");
  return check_launch("ufv_splice_static",
                      launch_kernel(splice_static_kernel, dim3(n_out_rows), dim3(128), 0,
                                    static_cast<cudaStream_t>(stream), static_cast<const uint4*>(text), labels_in,
                                    src_map, static_cast<uint4*>(out), labels_out, attn_out, row_bytes / 16,
                                    ignore_index));
}

extern "C" int ufv_abi_version(void) { return UFV_ABI_VERSION; }

extern "C" const char* ufv_last_error(void) { return ufv::g_error; }

extern "C" int ufv_struct_size(const char* name) {
  if (name == nullptr) return -1;
  if (strcmp(name, "ufv_mask_desc") == 0) return int(sizeof(ufv_mask_desc));
  if (strcmp(name, "ufv_peer_args") == 0) return int(sizeof(ufv_peer_args));
  if (strcmp(name, "ufv_dyn_args") == 0) return int(sizeof(ufv_dyn_args));
  if (strcmp(name, "ufv_encode_args") == 0) return int(sizeof(ufv_encode_args));
  return -1;
}

extern "C" int ufv_device_address(const void* host_ptr, uint64_t* dev_addr_host) {
  using namespace ufv;
  UFV_REQUIRE(host_ptr && dev_addr_host, UFV_E_NULL, "ufv_device_address: null pointer");
  void* dev = nullptr;
  const cudaError_t e = cudaHostGetDevicePointer(&dev, const_cast<void*>(host_ptr), 0);
  if (e != cudaSuccess) {
    cudaGetLastError();   // clear the sticky-less error state
    return fail(int(e), "ufv_device_address: %s (not mapped pinned memory?)", cudaGetErrorString(e));
  }
  *dev_addr_host = reinterpret_cast<uint64_t>(dev);
  return 0;
}

extern "C" int ufv_gather_rows(const void* in, const int32_t* row_map, void* out, int n_out_rows,
                               int row_bytes, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(n_out_rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0, UFV_E_SHAPE,
              "ufv_gather_rows: n_out_rows=%d row_bytes=%d", n_out_rows, row_bytes);
  if (n_out_rows == 0) return 0;
  UFV_REQUIRE(in && row_map && out, UFV_E_NULL, "ufv_gather_rows: null pointer");
  UFV_REQUIRE(aligned16(in) && aligned16(out), UFV_E_ALIGN, "ufv_gather_rows: unaligned buffer");
  return check_launch("ufv_gather_rows",
                      launch_kernel(gather_rows_kernel, dim3(n_out_rows), dim3(128), 0,
                                    static_cast<cudaStream_t>(stream), static_cast<const uint4*>(in), row_map,
                                    static_cast<uint4*>(out), row_bytes / 16));
}

extern "C" int ufv_compact_rows(const void* in, const int32_t* slot_off, const int32_t* counts, int n_obj,
                                void* out, int row_bytes, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(n_obj >= 0 && row_bytes > 0 && row_bytes % 16 == 0, UFV_E_SHAPE,
              "ufv_compact_rows: n_obj=%d row_bytes=%d", n_obj, row_bytes);
  if (n_obj == 0) return 0;
  UFV_REQUIRE(in && slot_off && counts && out, UFV_E_NULL, "ufv_compact_rows: null pointer");
  UFV_REQUIRE(aligned16(in) && aligned16(out), UFV_E_ALIGN, "ufv_compact_rows: unaligned buffer");
  return check_launch("ufv_compact_rows",
                      launch_kernel(compact_rows_kernel, dim3(n_obj), dim3(256), 0,
                                    static_cast<cudaStream_t>(stream), static_cast<const uint4*>(in), slot_off,
                                    counts, static_cast<uint4*>(out), row_bytes / 16));
}

extern "C" int ufv_splice_rows(const void* text, int n_text, const int32_t* region_pos, const void* tokens,
                               const int32_t* slot_off, const int32_t* counts, int n_obj, int m_pad, void* out,
                               int32_t* out_len, int32_t* row_src, int row_bytes, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(n_text >= 0 && n_obj >= 0 && m_pad >= 0 && row_bytes > 0 && row_bytes % 16 == 0, UFV_E_SHAPE,
              "ufv_splice_rows: n_text=%d n_obj=%d m_pad=%d row_bytes=%d", n_text, n_obj, m_pad, row_bytes);
  UFV_REQUIRE(n_obj <= n_text, UFV_E_SHAPE, "ufv_splice_rows: more placeholders (%d) than rows (%d)", n_obj, n_text);
  if (n_text + m_pad == 0) return 0;
  UFV_REQUIRE(out && (n_text == 0 || text) && (n_obj == 0 || (region_pos && slot_off && counts)) &&
                  (m_pad == 0 || tokens), UFV_E_NULL, "ufv_splice_rows: null pointer");
  UFV_REQUIRE(aligned16(text) && aligned16(tokens) && aligned16(out), UFV_E_ALIGN, "ufv_splice_rows: unaligned buffer");
  return check_launch("ufv_splice_rows",
                      launch_kernel(splice_rows_kernel, dim3(n_text + m_pad), dim3(128), 0,
                                    static_cast<cudaStream_t>(stream), static_cast<const uint4*>(text), n_text,
                                    region_pos, static_cast<const uint4*>(tokens), slot_off, counts, n_obj, m_pad,
                                    static_cast<uint4*>(out), out_len, row_src, row_bytes / 16));
}

extern "C" int ufv_encode(const ufv_encode_args* a, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(a != nullptr, UFV_E_NULL, "ufv_encode: args is null");
  const int side = a->n_patch_side;
  const bool dyn_mode = a->dyn_src != nullptr && a->dyn_dev != nullptr;
  const ufv_dyn_args* dyn = dyn_mode ? a->dyn_dev : nullptr;
  if (dyn_mode)
    UFV_REQUIRE(a->n_masks > 0 && a->n_obj > 0 && a->m_pad > 0, UFV_E_SHAPE,
                "ufv_encode: graph-replay mode needs a non-empty batch");
  int rc = launch_mask_to_patches(a->mask_desc, a->taps, a->n_masks, side, a->any_row_mode, a->bits, a->cnt, a->idx,
                                  a->idx_pitch, dyn_mode ? a->dyn_src : nullptr, dyn_mode ? a->dyn_dev : nullptr,
                                  stream);
  if (rc != 0) return rc;
  rc = ufv_mask_pool(a->feats, a->feat_dtype, a->n_rows, side * side, a->c, a->bits, a->cnt, a->grp_row, a->grp_off,
                     a->grp_member, a->n_groups, a->max_group, a->pooled, stream);
  if (rc != 0) return rc;
  rc = ttm_dispatch(a->pooled, a->c, a->obj_start, a->obj_len, a->slot_off, a->n_obj, a->max_len, a->k_keep,
                    a->merged, a->feat_dtype, nullptr, a->counts, nullptr, 0, a->sims, a->sims_pitch,
                    a->counts_host, a->epoch, dyn, stream);
  if (rc != 0) return rc;
  if (a->m_pad == 0)
    return a->peer != nullptr
               ? ufv_linear_gather(nullptr, nullptr, nullptr, 0, a->hid, a->hid, a->feat_dtype, a->peer, nullptr, 0, stream)
               : 0;
  rc = ufv_linear(a->merged, a->w1, a->b1, a->hidden, a->m_pad, a->hid, a->c, a->feat_dtype, 1, a->gemm_ws,
                  a->gemm_ws_bytes, stream);
  if (rc != 0) return rc;
  if (dyn_mode)             // output pointer / all-gather destinations are read from the device block
    return last_linear_dyn(a->hidden, a->w2, a->b2, a->m_pad, a->hid, a->hid, a->feat_dtype, a->peer, dyn,
                           a->gemm_ws, a->gemm_ws_bytes, a->tokens_row_map, stream);
  if (a->peer != nullptr)   // last Linear fused with the all-gather: tiles go straight to every rank
    return ufv_linear_gather(a->hidden, a->w2, a->b2, a->m_pad, a->hid, a->hid, a->feat_dtype, a->peer, a->gemm_ws,
                             a->gemm_ws_bytes, stream);
  return ufv_linear_scatter(a->hidden, a->w2, a->b2, a->tokens_out, a->m_pad, a->hid, a->hid, a->feat_dtype, 0,
                            a->tokens_row_map, a->gemm_ws, a->gemm_ws_bytes, stream);
}

// ---- graph replay ---------------------------------------------------------------------------------------
extern "C" int ufv_encode_graph_create(const ufv_encode_args* a, void** graph_out) {
  using namespace ufv;
  UFV_REQUIRE(a != nullptr && graph_out != nullptr, UFV_E_NULL, "ufv_encode_graph_create: null pointer");
  UFV_REQUIRE(a->dyn_src != nullptr && a->dyn_dev != nullptr, UFV_E_NULL,
              "ufv_encode_graph_create: args->dyn_src / dyn_dev must be set");
  *graph_out = nullptr;
  cudaStream_t cap = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
  if (e != cudaSuccess) return fail(int(e), "ufv_encode_graph_create: stream: %s", cudaGetErrorString(e));
  e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    cudaStreamDestroy(cap);
    return fail(int(e), "ufv_encode_graph_create: begin capture: %s", cudaGetErrorString(e));
  }
  const int rc = ufv_encode(a, cap);
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(cap, &graph);
  cudaStreamDestroy(cap);
  if (rc != 0) {
    if (graph != nullptr) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess || graph == nullptr) {
    cudaGetLastError();
    return fail(int(e), "ufv_encode_graph_create: end capture: %s", cudaGetErrorString(e));
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(int(e), "ufv_encode_graph_create: instantiate: %s", cudaGetErrorString(e));
  }
  *graph_out = exec;
  return 0;
}

extern "C" int ufv_encode_graph_launch(void* graph, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(graph != nullptr, UFV_E_NULL, "ufv_encode_graph_launch: graph is null");
  const cudaError_t e = cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(int(e), "ufv_encode_graph_launch: %s", cudaGetErrorString(e));
  }
  return 0;
}

extern "C" int ufv_encode_graph_destroy(void* graph) {
  if (graph != nullptr) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph));
  return 0;
}
