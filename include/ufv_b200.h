/*
 * ufv_b200.h -- C ABI of the B200-native object encoder (libufv_b200.so).
 *
 * Drop-in boundary.  The reference (Heven-Pan/UFVideo) has no FFI for this path: its object
 * encoder is the Python module `MaskExtractor` in ufvideo/model/layer.py, built at
 * ufvideo/model/videorefer_arch.py:39,92 and called at videorefer_arch.py:236.  Each entry
 * point below replaces the chain of ATen library calls named beside it; the Python host side
 * (ufvideo_b200/layer.py) keeps the reference module's constructor, forward signature,
 * parameter names and return types and binds these symbols through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 *     `_host`; `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = ok, < 0 = bad argument (UFV_E_*), > 0 = cudaError_t of a failed launch.
 *     ufv_last_error() returns a thread-local message for the last non-zero return.
 *   - no global mutable state, no hidden allocation: the caller owns every buffer, calls are
 *     stream-ordered and re-entrant across streams.
 *   - kernels are launched with programmatic stream serialization and open with
 *     griddepcontrol.wait, so back-to-back calls overlap their launch latency and prologues but
 *     never their data dependencies.
 *   - there is no CPU path: without a CUDA device every compute call returns an error.
 */
#ifndef UFV_B200_H
#define UFV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UFV_ABI_VERSION 9

/* element types */
enum { UFV_F32 = 0, UFV_BF16 = 1, UFV_F16 = 2, UFV_U8 = 3, UFV_RLE = 4 /* masks only */ };

/* argument errors */
enum {
  UFV_E_NULL = -1,      /* required pointer is null */
  UFV_E_SHAPE = -2,     /* size out of the supported range */
  UFV_E_DTYPE = -3,     /* unsupported element type */
  UFV_E_ALIGN = -4,     /* pointer / row pitch not 16-byte aligned */
  UFV_E_DRIVER = -5     /* CUDA driver entry point unavailable (tensor-map encode) */
};

#define UFV_BITS_WORDS 24        /* uint32 words per patch bitmask row (729 bits -> 23, padded) */
#define UFV_MAX_PATCH_SIDE 27    /* kernels are sized for up to 27 x 27 patches              */
#define UFV_MAX_GROUP 64         /* object-frames pooled together from one staged frame tile */

/* One object-frame's mask plane (32 bytes).  dtype UFV_RLE: a COCO run-length mask -- `addr` points to the
 * int32 cumulative run ends, `pitch` = number of runs, `aux` = image height; pixel p is on iff the first
 * run whose end exceeds p has an odd index (runs alternate, starting with zeros).  Dense dtypes:  `addr` may point into device memory or into pinned,
 * device-mapped HOST memory (ufv_device_address): the kernel then reads the mask in place over
 * PCIe, touching only the rows its taps need. */
typedef struct ufv_mask_desc {
  uint64_t addr;      /* device address of element (0,0) of the mask plane                     */
  int32_t pitch;      /* row pitch in elements                                                  */
  int32_t dtype;      /* UFV_U8 (also bool) / UFV_F32 / UFV_BF16 / UFV_F16                      */
  int32_t tap_off;    /* offset of the plane's tap table inside `taps`, in int32 units          */
  int32_t group;      /* pool group this object-frame belongs to (informational; kernels do not read it) */
  int32_t flags;      /* bit 0: read in row mode when the tap span fits (see ufv_mask_to_patches) */
  int32_t aux;        /* UFV_RLE: image height h (pixels are numbered column-major, p = x * h + y)  */
} ufv_mask_desc;

int ufv_abi_version(void);
const char* ufv_last_error(void);
/* sizeof() of an ABI struct as this library was compiled, by name ("ufv_mask_desc", "ufv_peer_args",
 * "ufv_dyn_args", "ufv_encode_args"); -1 for an unknown name.  Lets a binding verify its struct mirrors. */
int ufv_struct_size(const char* name);

/* Device-visible address of a pinned (page-locked, mapped) host buffer, so that kernels can read
 * host-resident masks in place.  Returns 0 and sets *dev_addr, or the cudaError_t when `host_ptr`
 * is not mapped pinned memory. */
int ufv_device_address(const void* host_ptr, uint64_t* dev_addr_host);

/* ---------------------------------------------------------------------------------------------
 * Host helper: tap table of the bilinear resize of an h x w mask to n_out x n_out.
 * Replaces the index/weight computation inside F.interpolate(mode='bilinear',
 * align_corners=False) at layer.py:139, evaluated in fp32 exactly as ATen does, and folds the
 * 'pad' aspect mode (layer.py:77-86) into the indices.
 * taps_host[4 * n_out] = h0[n_out], h1[n_out], w0[n_out], w1[n_out]; an entry is the source
 * row / column of that tap, or -1 when the tap contributes nothing (zero weight, or it falls
 * into the zero padding).  Pure integer output; runs on the host.
 * -------------------------------------------------------------------------------------------*/
int ufv_tap_table(int h, int w, int n_out, int pad_square, int32_t* taps_host);

/* ---------------------------------------------------------------------------------------------
 * Kernel 1: mask resize + binarise -> patch bitmask, count and index list per object-frame.
 * Replaces F.interpolate + (mask > 0) + mask.sum at layer.py:139,143,145.  Bit-exact.
 * Per object-frame the kernel reads in tap mode (every tap gathered individually: lowest latency
 * for masks in HBM) or, when desc.flags bit 0 is set and the tap columns of a row span <= 4 KB,
 * in row mode (the source rows of every output row are read with coalesced 16-byte loads: few,
 * wide requests, the efficient pattern over PCIe for pinned host masks).  Both give the same bits.
 * Row mode reads 16-byte chunks aligned around the tap span of each source row; the first and last
 * chunk of a row are clamped to the span, so no byte outside the tapped columns is ever touched.
 *   desc[n_masks]        one ufv_mask_desc per object-frame
 *   any_row_mode         non-zero iff some descriptor sets flags bit 0 (selects the kernel variant
 *                        that carries the row-mode flag table; 0 = lean tap-mode kernel)
 *   taps                 concatenated tap tables (ufv_tap_table) the descriptors point into
 *   bits_out[n_masks*UFV_BITS_WORDS]  bit p%32 of word p/32 = patch p (row-major h,w) is on
 *   cnt_out[n_masks]     number of on patches
 *   idx_out              optional (may be null): [n_masks * idx_pitch] uint16, ascending patch
 *                        indices of the on patches, first cnt entries valid
 * -------------------------------------------------------------------------------------------*/
int ufv_mask_to_patches(const ufv_mask_desc* desc, const int32_t* taps, int n_masks, int n_out,
                        int any_row_mode, uint32_t* bits_out, int32_t* cnt_out, uint16_t* idx_out, int idx_pitch,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 2: segmented mask pool.  Replaces the gather feats[ann_index], the layout permute,
 * the fp32 upcast (layer.py:98-104) and the masked mean (layer.py:145-147).
 *   feats [n_rows, n_patch, c] of feat_dtype (UFV_F32 / UFV_BF16 / UFV_F16), contiguous
 *   bits / cnt: patch bitmasks and on-counts of all object-frames, from ufv_mask_to_patches
 *   groups: group g pools object-frames grp_member[grp_off[g] .. grp_off[g+1]) (at most
 *           UFV_MAX_GROUP of them), all of which read feature row grp_row[g] -- the row is streamed ONCE
 *           for the whole group, in windows of 32 consecutive patches: the kernel ORs the members'
 *           bitmasks, skips windows nobody needs, fetches mostly-needed windows as one 2-D tile and
 *           sparse ones row by row.  Up to 8 members: every staged row is added into the members whose
 *           bit is set (warp-uniform predicates).  9 .. 64 members: each consumer warp owns a few members
 *           and walks only the rows its members pool (16-bit features, up to 32 members: 256-channel
 *           slices).  Several groups may name the same feature row: with many object-frames on a frame
 *           the kernel is bound by instructions, not bytes, and equal sub-groups of <= 16 run faster
 *           than one group of 64 (what ufvideo_b200/packer.py builds; every split gives the same bits).
 *           max_group = the largest group size in this call (selects the kernel variant)
 *   pooled_out fp32 [n_masks, c]:  sum over on patches in ascending patch order, divided by
 *           (float(cnt) + 1e-8f); an all-off mask gives an exact zero row.
 * -------------------------------------------------------------------------------------------*/
int ufv_mask_pool(const void* feats, int feat_dtype, int64_t n_rows, int n_patch, int c,
                  const uint32_t* bits, const int32_t* cnt, const int32_t* grp_row, const int32_t* grp_off,
                  const int32_t* grp_member, int n_groups, int max_group, float* pooled_out, void* stream);

/* Adjoint of ufv_mask_pool w.r.t. the features (training; the reference gets it from autograd through
 * layer.py:98-104,145-147).  w fp32 [n_masks, c] = d_pooled / (cnt + 1e-8); the object-frames that pool
 * from feature row r are row_member[row_off[r] .. row_off[r+1]) (at most max_members per row);
 * d_feats [n_rows, n_patch, c] of feat_dtype is written completely (rows nobody pools from: zeros):
 *   d_feats[r, p, :] = sum over those object-frames j with patch p on (bits) of w[j, :]. */
int ufv_mask_pool_backward(const float* w, const uint32_t* bits, const int32_t* row_off,
                           const int32_t* row_member, int64_t n_rows, int max_members, int n_patch, int c,
                           void* d_feats, int feat_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 3: fused temporal token merge.  Replaces token_merge (layer.py:6-33), its dispatch
 * (layer.py:110-119) and the downcast (layer.py:123).
 *   object o owns pooled rows [obj_start[o], obj_start[o] + obj_len[o]) and writes its tokens
 *   to rows slot_off[o] .. of tokens_out, which has min(obj_len[o], k_keep) rows reserved for
 *   it; rows past counts_out[o] are zero-filled.
 *   tokens_out [m_pad, c] of out_dtype; tokens_f32_out (optional) the same rows before the
 *   downcast; cuts_out (optional) [n_obj * cut_pitch_words] uint32, bit i set = run boundary
 *   after token i; sims_out [n_obj * sims_pitch] fp32 adjacent cosine similarities: scratch the two
 *   launches communicate through (sims_pitch >= max_len - 1), required when max_len > k_keep.
 *   Early read-back of the counts (optional, pass counts_host = null to skip): counts_host is the
 *   device-visible address (ufv_device_address) of a pinned int32[n_obj]; object o's count is also
 *   stored there as ONE word, (epoch << 16) | count, epoch in [1, 32767].  The host polls until all
 *   n_obj words carry the call's epoch -- no fence, copy, event or stream synchronisation needed.
 *   Objects of up to 64 frames are merged by one fused launch (a CTA per object), longer ones by
 *   a similarity launch (a warp per adjacent pair) plus a merge launch (a CTA per output token).
 * -------------------------------------------------------------------------------------------*/
int ufv_ttm(const float* pooled, int c, const int32_t* obj_start, const int32_t* obj_len,
            const int32_t* slot_off, int n_obj, int max_len, int k_keep, void* tokens_out,
            int out_dtype, float* tokens_f32_out, int32_t* counts_out, uint32_t* cuts_out,
            int cut_pitch_words, float* sims_out, int sims_pitch, int32_t* counts_host,
            int32_t epoch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 4: one Linear (+ optional exact-erf GELU) of the object projector,
 *   y[m, n] = act(x[m, k] . w[n, k]^T + bias[n]),  act = GELU applied to the value rounded to
 *   `dtype` (the reference rounds the hidden activation between modules, layer.py:55-59).
 * UFV_BF16 / UFV_F16: tcgen05 tensor-core GEMM with TMEM accumulators (fp32 accumulate);
 * UFV_F32: fp32 CUDA-core GEMM.  x, w, y row-major, 16-byte aligned, k % 8 == 0, n % 8 == 0.
 * Few tokens (m <= 512) and a deep contraction (k >= 2048): K is split over a thread-block cluster
 * whose CTAs exchange fp32 partial tiles through `ws` (device scratch of at least
 * ufv_linear_ws_bytes(m, n, k, dtype) bytes, 16-byte aligned; contents are don't-care on entry and exit).
 * With ws == null (or too small) the full-K kernel runs instead: same result up to fp32 summation order.
 * -------------------------------------------------------------------------------------------*/
int64_t ufv_linear_ws_bytes(int m, int n, int k, int dtype);
int ufv_linear(const void* x, const void* w, const void* bias, void* y, int m, int n, int k,
               int dtype, int gelu, void* ws, int64_t ws_bytes, void* stream);
/* Same, with a scatter epilogue: output row r is stored as row row_map[r] of y (row pitch n elements), or
 * dropped when row_map[r] < 0.  row_map int32 [m] on the device; null = identity.  This is how the last Linear
 * of the projector writes the object tokens straight into the caller's inputs_embeds at their <region>
 * positions (videorefer_arch.py:300-311), with no tokens tensor and no copy in between. */
int ufv_linear_scatter(const void* x, const void* w, const void* bias, void* y, int m, int n, int k,
                       int dtype, int gelu, const int32_t* row_map, void* ws, int64_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward pass of the projector (the region encoder is trainable in the reference,
 * videorefer_arch.py:94-96, train.py:887-890; there it is torch autograd over cuBLAS).  Every product runs
 * on the same tcgen05 kernel as the forward pass:
 *   ufv_linear_ex   y = epilogue(x . w^T + bias), bf16 / fp16, bias may be null.  epilogue:
 *                   0  none
 *                   1  GELU of the rounded sum; aux (optional, [m, n]) also receives the rounded
 *                      pre-activation -- what the training forward keeps for the backward pass
 *                   2  multiply by GELU'(aux[m, n]): the dgrad through Linear-2 and the GELU in one pass,
 *                      dZ1 = (dY . W2) * GELU'(Z1), with w = W2^T
 *   ufv_transpose16 out[c, r] = in[r, c] for 2-byte elements, output rows padded with zeros to out_pitch
 *                   (>= rows, a multiple of 8): turns W, dY, H, X into the K-major operands the kernel takes:
 *                   dgrad  dX = dZ . W      -> x = dZ,       w = W^T
 *                   wgrad  dW = dZ^T . X    -> x = dZ^T,     w = X^T   (contraction over the tokens)
 *   ufv_colsum      out[c] = sum_r x[r, c] (fp32 accumulation): the bias gradients
 * -------------------------------------------------------------------------------------------*/
int ufv_linear_ex(const void* x, const void* w, const void* bias, void* y, int m, int n, int k, int dtype,
                  int epilogue, void* aux, void* stream);
int ufv_transpose16(const void* in, void* out, int rows, int cols, int out_pitch, void* stream);
int ufv_colsum(const void* x, void* out, int m, int n, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 4 fused with the result-collection all-gather (multi-GPU, clips sharded over ranks).
 * The reference collects per-rank results through files (eval/inference_PixRQA.py:214); here the
 * epilogue of the last Linear stores every output tile straight into the gathered buffer of
 * EVERY rank over NVLink -- one multimem.st per 16 bytes through the NVSwitch multicast address,
 * or one st per peer -- so the transfer rides under the GEMM instead of following it.
 *   dst[i]       address of element (0, 0) of THIS rank's token rows inside destination copy i
 *                (the multicast alias of the symmetric gathered buffer when multimem != 0, else
 *                one entry per rank, this rank included)
 *   tail_src     int32 [tail_words] in local memory (reserved rows per object + token counts,
 *                written earlier in the stream), copied to tail_dst[i] by the last CTA
 *   flag[i]      address of this rank's int32 arrival flag in destination i; written with
 *                release.sys semantics = flag_value after every store of the call is fenced
 *   ticket       one zeroed uint32 in local device memory (zero again on completion)
 * Receivers wait with ufv_wait_flags.
 * -------------------------------------------------------------------------------------------*/
#define UFV_MAX_PEER_DST 8
typedef struct ufv_peer_args {
  uint64_t dst[UFV_MAX_PEER_DST];
  uint64_t tail_dst[UFV_MAX_PEER_DST];
  uint64_t flag[UFV_MAX_PEER_DST];
  const int32_t* tail_src;
  uint32_t* ticket;
  int32_t tail_words;
  int32_t n_dst;
  int32_t multimem;
  int32_t flag_value;
} ufv_peer_args;

/* y[m, n] = x . w^T + bias (bf16 / fp16, no activation), written to every peer->dst instead of a
 * local y. */
int ufv_linear_gather(const void* x, const void* w, const void* bias, int m, int n, int k, int dtype,
                      const ufv_peer_args* peer_host, void* ws, int64_t ws_bytes, void* stream);

/* The same collection as a separate step: push `bytes` (multiple of 16) starting at `src` -- rows the last Linear
 * wrote locally, normally this rank's own slice of the symmetric gathered buffer -- to every peer->dst, forward the
 * tail and raise the flags exactly as ufv_linear_gather does.  Launched on a side stream it takes the NVLink
 * transfer (world x payload per link and step) off the compute stream's critical path. */
int ufv_peer_push(const void* src, int64_t bytes, const ufv_peer_args* peer_host, void* stream);

/* Block the stream until flags[0 .. n) (int32, local memory) have all reached `value` (>=, acquire.sys
 * loads; step counters only grow).  Gives up after ~timeout_ms (0 = 2000) and stores 1 to *timed_out
 * (optional int32 in device memory or device-mapped pinned host memory, where the host can poll it
 * without synchronising). */
int ufv_wait_flags(const int32_t* flags, int n, int32_t value, int timeout_ms, int32_t* timed_out,
                   void* stream);

/* Per-call values of a replayed launch sequence (see ufv_encode_graph_create): 256 bytes. */
typedef struct ufv_dyn_args {
  uint64_t tokens_out;     /* output of the last Linear for this call                              */
  uint64_t counts_out;     /* int32 [n_obj] device array for the counts, 0 = args->counts           */
  int32_t epoch;           /* tag of this call's counts words (ufv_ttm)                             */
  int32_t reserved;
  ufv_peer_args peer;      /* fused all-gather destinations of this call (when args->peer != null)   */
  uint64_t pad;
} ufv_dyn_args;

/* ---------------------------------------------------------------------------------------------
 * The whole path in one call (kernels 1-4 chained on `stream`).  Replaces
 * MaskExtractor.forward (layer.py:63-128) minus the host read-back of counts.
 * -------------------------------------------------------------------------------------------*/
typedef struct ufv_encode_args {
  /* features */
  const void* feats; int32_t feat_dtype; int32_t n_patch_side; int64_t n_rows; int32_t c; int32_t hid;
  /* masks -> patches */
  const ufv_mask_desc* mask_desc; const int32_t* taps; int32_t n_masks; int32_t idx_pitch;
  int32_t any_row_mode; int32_t reserved1;
  uint32_t* bits; int32_t* cnt; uint16_t* idx;            /* idx optional */
  /* pool */
  const int32_t* grp_row; const int32_t* grp_off; const int32_t* grp_member; int32_t n_groups;
  int32_t max_group;
  float* pooled;
  /* merge */
  const int32_t* obj_start; const int32_t* obj_len; const int32_t* slot_off;
  int32_t n_obj; int32_t max_len; int32_t k_keep; int32_t m_pad;
  void* merged; int32_t* counts;
  float* sims; int32_t sims_pitch; int32_t reserved0;      /* fp32 [n_obj * sims_pitch] scratch of the merge */
  /* optional early read-back of the token counts (see ufv_ttm): the merge kernel stores them into
   * the pinned buffer behind counts_host tagged with `epoch`, so the caller can build the
   * reference's list[int] while the projector is still running */
  int32_t* counts_host; int32_t epoch; int32_t reserved;
  /* projector: feat_linear.0 / feat_linear.2 (layer.py:55-59) */
  const void* w1; const void* b1; const void* w2; const void* b2;
  void* hidden; void* tokens_out;
  /* scratch of the split-K Linears (ufv_linear): max over the two layers of ufv_linear_ws_bytes; may be null */
  void* gemm_ws; int64_t gemm_ws_bytes;
  /* optional scatter epilogue of the last Linear (ufv_linear_scatter): int32 [m_pad] on the device, row r of
   * the padded tokens is stored as row tokens_row_map[r] of tokens_out (< 0: dropped).  Ignored with `peer`. */
  const int32_t* tokens_row_map;
  /* optional: fuse the result-collection all-gather into the last Linear (ufv_linear_gather);
   * tokens_out is then unused */
  const ufv_peer_args* peer;
  /* optional: per-call values through a device block instead of kernel parameters (graph replay) */
  const ufv_dyn_args* dyn_src;   /* device-visible address of the caller's pinned block */
  ufv_dyn_args* dyn_dev;         /* 256 bytes of device scratch */
} ufv_encode_args;

int ufv_encode(const ufv_encode_args* args_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Replaying the chained path as a CUDA graph.  A repeated call differs from the previous one only
 * in a few values (the output tensor, the tag of the counts words, the all-gather slot); those are
 * kept OUT of the kernel parameters: the caller writes them into a pinned ufv_dyn_args block,
 * kernel 1 copies the block to device memory (dyn_dev) and the later kernels read it there.  With
 * args->dyn_src / dyn_dev set, the whole launch sequence is therefore constant and can be captured
 * once (ufv_encode_graph_create) and relaunched with one driver call (ufv_encode_graph_launch).
 * The block must stay unchanged until kernel 1 of that launch has run (callers that wait for the
 * counts before their next call satisfy this).
 * -------------------------------------------------------------------------------------------*/
int ufv_encode_graph_create(const ufv_encode_args* args_host, void** graph_out_host);
int ufv_encode_graph_launch(void* graph, void* stream);
int ufv_encode_graph_destroy(void* graph);

/* Compaction after merge ties: object o's first counts[o] rows (starting at row slot_off[o] of
 * `in`) are packed back to back into `out` in object order; `row_bytes` per row (multiple of 16).
 * out must hold sum(counts) rows.  Replaces the variable-length torch.cat at layer.py:121. */
int ufv_compact_rows(const void* in, const int32_t* slot_off, const int32_t* counts, int n_obj,
                     void* out, int row_bytes, void* stream);

/* <region> splice, the consumer of the path (videorefer_arch.py:300-311), on the device.
 *   text [n_text, row]    embeddings of the flattened token sequence, one placeholder row per object
 *   region_pos[n_obj]     row of object o's placeholder in `text`, strictly ascending
 *   tokens [m_pad, row]   padded object tokens (object o at slot_off[o]), counts[n_obj] on the device
 *   out                   the sequence with placeholder o replaced by object o's counts[o] rows; must
 *                         hold n_text - n_obj + m_pad rows; *out_len (optional) = rows actually written
 *   row_src (optional)    int32 per output row: text row index, or -(token row + 1)
 * Driven by the device-side counts, so inputs_embeds can be built without the host copy of
 * region_token_nums. */
int ufv_splice_rows(const void* text, int n_text, const int32_t* region_pos, const void* tokens,
                    const int32_t* slot_off, const int32_t* counts, int n_obj, int m_pad, void* out,
                    int32_t* out_len, int32_t* row_src, int row_bytes, void* stream);

/* <region> splice with batch padding and labels for a layout known on the host (the reference's
 * prepare_inputs_labels_for_multimodal, videorefer_arch.py:291-368, for the batch [B, L_max] flattened to
 * n_out_rows = B * L_max rows).  src_map[i] >= 0: output row i is text row src_map[i] (embedding copied, label
 * labels_in[src_map[i]], attention 1); -1: padding (zero embedding, label ignore_index, attention 0); -2: a
 * region-token row -- its embedding is written by ufv_linear_scatter / ufv_encode (tokens_row_map), here it
 * only gets label ignore_index and attention 1.  labels_in / labels_out / attn_out are optional (null). */
int ufv_splice_static(const void* text, const int64_t* labels_in, const int32_t* src_map, void* out,
                      int64_t* labels_out, uint8_t* attn_out, int n_out_rows, int row_bytes,
                      int64_t ignore_index, void* stream);

/* Gather rows: out[i, :] = in[row_map[i], :], `row_bytes` per row (multiple of 16). */
int ufv_gather_rows(const void* in, const int32_t* row_map, void* out, int n_out_rows,
                    int row_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UFV_B200_H */
