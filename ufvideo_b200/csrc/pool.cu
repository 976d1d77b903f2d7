// Kernel 2: segmented mask pool.
//
// Replaces, for every object-frame, the reference's gather feats[ann_index], the NCHW permute,
// the fp32 upcast (ufvideo/model/layer.py:98-104) and the masked mean
// (x * mask / denorm).sum(-1).sum(-1) (layer.py:145-147), which materialise three dense fp32
// [q, C, 27, 27] temporaries.  Here each needed feature row is read from HBM once per frame
// and shared by every object on that frame.
//
// Work item = (group, 128-channel slice).  A group is one feature row (frame) plus up to 64
// object-frames pooled from it.  The frame's 729 patch rows are walked in WINDOWS of 32 consecutive
// patches.  Every CTA first ORs its members' patch bitmasks (kernel 1's output, 96 bytes each) into the
// frame's union; then the producer warp moves, per non-empty window, the slice of the needed rows into a
// ring of shared-memory stages with the TMA engine: ONE 2-D tiled tensor-map load of all 32 rows when most
// of the window is needed (a single request; the few unneeded rows ride along), otherwise one 1-D bulk
// copy per needed row (measured: 256-byte row copies stream at ~3.4 TB/s, 32-row tiles at ~5.4 TB/s).
// Windows no member needs are skipped entirely.  The row is fetched ONCE however many objects pool from it.
// Two consumer designs:
//   mask_pool_kernel         (<= 8 members)  two consumer warps, 64 channels each, 2 per lane.  Every staged row
//            is added, in ascending patch order, into the fp32 accumulators of the members whose bit is set
//            (warp-uniform predicates, packed f32x2 adds).  Best when most members pool most rows.  (A single
//            consumer warp with 4 channels per lane needs fewer instructions but measured slower: 43.6 vs
//            41.3 us on c2, 133 vs 89 us with 8 objects per frame -- two warps hide each other's latencies.)
//   mask_pool_sparse_kernel  (9 .. 64 members)  eight consumer warps, each owning a few members and all 128
//            channels (4 per lane): a warp walks the set bits of its member's window word and adds only
//            those rows.  With 16 - 64 objects on a frame a row is pooled by a small fraction of them, and
//            the predicated design spends its issue slots on adds that are switched off (measured on c4,
//            16 blobs per frame: 306 us predicated in one pass, 263 us re-streaming the frame per 8 members).
// Accumulation order per (object, channel) is the plain ascending-patch sequence, independent
// of any blocking, which is what oracle/restatement.py::mask_pool restates bit-for-bit.
//
// Roofline: HBM.  Algorithmic bytes per group = n_union_patches * C * sizeof(feat).
// kTileMin: a full window whose union has at least this many rows is fetched as one tile.
#include "common.cuh"

#include <cuda.h>
#include <cstdlib>

namespace ufv {

constexpr int kPoolCh = 128;          // channels per CTA slice
// 24.6 KB of ring per CTA (8 CTAs per SM) is the sweet spot; 32 rows x 3 stages measured 0-8 % faster than
// 16 x 6 at the same footprint (tools/pool_variant_sweep.sh), more or fewer bytes per CTA are both slower.
#ifndef UFV_POOL_ROWS
#define UFV_POOL_ROWS 32
#endif
#ifndef UFV_POOL_STAGES
#define UFV_POOL_STAGES 3
#endif
constexpr int kPoolRows = UFV_POOL_ROWS;      // patch rows per stage (multiple of 16, <= 32)
static_assert(kPoolRows % 16 == 0 && kPoolRows <= 32, "stage rows: one producer lane per row, 16-byte mask copies");

constexpr int kPoolStages = UFV_POOL_STAGES;
constexpr int kPoolConsumers = 2;             // dense kernel: consumer warps, 64 channels each
constexpr int kPoolThreads = 32 * (kPoolConsumers + 1);
// Sparse kernel: NW consumer warps (128 channels each), member j belongs to warp j % NW, slot j / NW.  What the
// kernel needs is warps in flight: 8 warps with 2 / 4 / 8 members each at seven CTAs per SM and 32 registers
// (measured on c4: 238 us, against 256 us at five CTAs with 40 registers and a deeper ring; on c5-wide, 64 members:
// 82 us with 8 x 8 at four to five CTAs and ~50 registers, 123 us at seven CTAs with the accumulators in local
// memory, 103 us with 16 warps x 4 members at three CTAs and an 8-stage ring, 111 us with 8 x 8 at three CTAs).
template <int MPW, int NW> struct SparseCfg {
  static constexpr int kThreads = 32 * (NW + 1);
  static constexpr int kMinCtas = NW > 8 ? 3 : MPW <= 4 ? 7 : 4;   // 8 members per warp: 32 accumulator registers alone
  static constexpr int kStages = NW <= 8 ? UFV_POOL_STAGES : 8;
  // 256-channel slices (16-bit features, up to 32 members).  Measured on c4 (16 members) and on 2 x 64 frames x 32
  // members, stages x CTAs per SM: 2 members per warp  3x4 210 us, 2x5 213, 2x6 201 (32 registers, 16 B of spill),
  // 4x3 236;  4 members per warp  3x4 92.6 us, 4x3 86.4 (64 registers), 2x5 102, 2x6 232 (spills in the loop).
  static constexpr int kWideStages = MPW <= 2 ? 2 : 4;
  static constexpr int kWideCtas = MPW <= 2 ? 6 : 3;
};
static_assert(kPoolRows == 32, "a window is one 32-bit word of the patch bitmasks");

// One ring stage holds a CHUNK of the frame: either one full window that is mostly needed, fetched as a 2-D tile
// (row r of the window sits in slot r), or a run of consecutive windows whose needed rows -- at most 32 in
// total -- are copied one by one and packed into slots 0, 1, 2, ... in ascending patch order (a sparse frame
// of 80 needed rows is 3 stages, not 23).  Producer and consumers derive the same chunks from the union words.
struct PoolChunk {
  int w0, w1;      // windows [w0, w1)
  int rows;        // needed rows in the chunk
  bool tile;
};

__device__ __forceinline__ bool window_is_tile(int win, uint32_t u, int n_patch, int use_tmap, int tile_min) {
  return use_tmap && 32 * win + 32 <= n_patch && __popc(u) >= tile_min;
}

// The chunk table of a frame, built once per CTA by one thread (at most 23 windows): entry = w0 | w1 << 8 |
// rows << 16 | tile << 24.  Returns the number of chunks.
__device__ __forceinline__ int build_chunk_table(const uint32_t* s_union, int n_win, int n_patch, int use_tmap,
                                                 int tile_min, uint32_t* s_chunk) {
  int n = 0, win = 0;
  while (true) {
    while (win < n_win && s_union[win] == 0u) ++win;
    if (win >= n_win) break;
    const uint32_t u = s_union[win];
    int rows = __popc(u);
    const bool tile = window_is_tile(win, u, n_patch, use_tmap, tile_min);
    int e = win + 1;
    if (!tile) {
      while (e < n_win) {
        const uint32_t u2 = s_union[e];
        if (u2 != 0u) {
          if (window_is_tile(e, u2, n_patch, use_tmap, tile_min) || rows + __popc(u2) > kPoolRows) break;
          rows += __popc(u2);
        }
        ++e;
      }
    }
    s_chunk[n++] = uint32_t(win) | (uint32_t(e) << 8) | (uint32_t(rows) << 16) | (tile ? 1u << 24 : 0u);
    win = e;
  }
  return n;
}

__device__ __forceinline__ PoolChunk unpack_chunk(uint32_t v) {
  PoolChunk ch;
  ch.w0 = int(v & 0xffu);
  ch.w1 = int((v >> 8) & 0xffu);
  ch.rows = int((v >> 16) & 0xffu);
  ch.tile = (v >> 24) != 0u;
  return ch;
}

// Frames whose needed rows are spread thinly over many windows take the packed path (chunk table, several
// windows per stage); everything else walks the windows directly, one stage per non-empty window -- the two
// paths are separate loops so that the common, dense one carries none of the packing machinery.
// Called by ONE thread after the union words are complete; returns the number of chunks, or 0 for the direct path.
__device__ __forceinline__ int plan_frame(const uint32_t* s_union, int n_win, int n_patch, int use_tmap, int tile_min,
                                          uint32_t* s_chunk) {
  int windows = 0, rows = 0;
  for (int w = 0; w < n_win; ++w) {
    const uint32_t u = s_union[w];
    windows += u != 0u;
    rows += __popc(u);
  }
  if (windows <= 2 * ((rows + kPoolRows - 1) / kPoolRows)) return 0;
  return build_chunk_table(s_union, n_win, n_patch, use_tmap, tile_min, s_chunk);
}

// Direct path, per non-empty window: the producer warp's part.  `u` = union word of the window (bit r = patch
// 32 * win + r is needed by some member; it sits in slot r of the stage).
template <typename T, int CH = kPoolCh>
__device__ __forceinline__ void produce_window(const CUtensorMap* tmap, int use_tmap, const T* __restrict__ feats,
                                               int64_t row_base, int c, int ch0, uint32_t slice_bytes, int win,
                                               int n_patch, uint32_t u, int tile_min, T* dst, uint64_t* full_bar,
                                               int lane) {
  const bool tile = window_is_tile(win, u, n_patch, use_tmap, tile_min);
  if (lane == 0) {
    mbar_arrive_expect_tx(full_bar, tile ? uint32_t(kPoolRows) * CH * sizeof(T) : uint32_t(__popc(u)) * slice_bytes);
    if (tile) tma_load_2d(dst, tmap, ch0, int(row_base + 32 * win), full_bar);
  }
  if (!tile) {
    __syncwarp();
    if ((u >> lane) & 1u)
      bulk_g2s(dst + lane * CH, feats + (row_base + 32 * win + lane) * int64_t(c) + ch0, slice_bytes, full_bar);
  }
}

// slot of patch (win, r) inside a packed chunk whose earlier windows hold `base` rows
__device__ __forceinline__ int packed_slot(int base, uint32_t u, int r) { return base + __popc(u & ((1u << r) - 1u)); }

// The producer warp's part for one chunk (the caller has waited for the stage to be empty and, in the dense
// kernel, written the stage's member masks).  Issues the expect-tx arrive, then the copies.
template <typename T>
__device__ __forceinline__ void produce_chunk(const CUtensorMap* tmap, const T* __restrict__ feats, int64_t row_base,
                                              int c, int ch0, uint32_t slice_bytes, const PoolChunk& ch,
                                              const uint32_t* s_union, T* dst, uint64_t* full_bar, int lane) {
  if (lane == 0) {
    mbar_arrive_expect_tx(full_bar, ch.tile ? uint32_t(kPoolRows) * kPoolCh * sizeof(T) : uint32_t(ch.rows) * slice_bytes);
    if (ch.tile) tma_load_2d(dst, tmap, ch0, int(row_base + 32 * ch.w0), full_bar);
  }
  if (!ch.tile) {
    __syncwarp();
    int base = 0;
    for (int w = ch.w0; w < ch.w1; ++w) {
      const uint32_t u = s_union[w];
      if ((u >> lane) & 1u)
        bulk_g2s(dst + packed_slot(base, u, lane) * kPoolCh, feats + (row_base + 32 * w + lane) * int64_t(c) + ch0,
                 slice_bytes, full_bar);
      base += __popc(u);
    }
  }
}

__device__ __forceinline__ void add2(float2& acc, float2 v) {
  unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long b = *reinterpret_cast<unsigned long long*>(&v);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  acc = *reinterpret_cast<float2*>(&a);
}

// two adjacent channels of one staged row -> fp32 pair
template <typename T> struct Pair;
template <> struct Pair<float> {
  using Raw = float2;
  __device__ static float2 cvt(Raw r) { return r; }
};
template <> struct Pair<__nv_bfloat16> {
  using Raw = uint32_t;   // bf16 -> fp32 is a 16-bit shift
  __device__ static float2 cvt(Raw r) { return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u)); }
};
template <> struct Pair<__half> {
  using Raw = uint32_t;
  __device__ static float2 cvt(Raw r) { return __half22float2(*reinterpret_cast<const __half2*>(&r)); }
};

template <typename T, int OT>
__global__ void __launch_bounds__(kPoolThreads, sizeof(T) == 4 ? 4 : 8)
mask_pool_kernel(const __grid_constant__ CUtensorMap tmap, int use_tmap, const T* __restrict__ feats,
                 int n_patch, int c, int n_slices, const uint32_t* __restrict__ bits,
                 const int32_t* __restrict__ cnt, const int32_t* __restrict__ grp_row,
                 const int32_t* __restrict__ grp_off, const int32_t* __restrict__ grp_member, int tile_min,
                 float* __restrict__ pooled) {
  constexpr int S = kPoolStages, R = kPoolRows;
  using Raw = typename Pair<T>::Raw;
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  T* ring = reinterpret_cast<T*>(dyn_smem);                               // [S][R][kPoolCh]
  __shared__ uint32_t s_bits[8][UFV_BITS_WORDS];                           // members' patch bitmasks
  __shared__ uint32_t s_union[UFV_BITS_WORDS];
  __shared__ uint32_t s_chunk[UFV_BITS_WORDS];                             // chunk table of the frame
  __shared__ int s_n_chunks;
  __shared__ __align__(16) uint8_t s_omask[S][R];                          // member masks of the staged rows
  __shared__ __align__(8) uint64_t full_bar[S];
  __shared__ __align__(8) uint64_t empty_bar[S];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int g = blockIdx.x / n_slices;
  const int slice = blockIdx.x - g * n_slices;
  const int ch0 = slice * kPoolCh;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kPoolConsumers);
    }
    mbar_fence_init();
    if (use_tmap) tma_prefetch_desc(&tmap);
  }
  // the group's extent is plan data (uploaded before kernel 1 ran): fetched while kernel 1 drains
  const int m0 = grp_off[g];
  const int n_mem = grp_off[g + 1] - m0;
  const int row = grp_row[g];
  // 8 members x 24 words = 192 bitmask words for 96 threads: each thread's two member indices are plan data too
  constexpr int kBitLoads = (8 * UFV_BITS_WORDS + kPoolThreads - 1) / kPoolThreads;
  int member_row[kBitLoads];
#pragma unroll
  for (int j = 0; j < kBitLoads; ++j) {
    const int o = (tid + j * kPoolThreads) / UFV_BITS_WORDS;
    member_row[j] = o < n_mem && o < 8 ? grp_member[m0 + o] : -1;
  }
  __syncthreads();
  pdl_wait();                  // patch bitmasks and counts come from kernel 1
  pdl_launch_dependents();
#pragma unroll
  for (int j = 0; j < kBitLoads; ++j) {
    const int i = tid + j * kPoolThreads;
    if (i < 8 * UFV_BITS_WORDS) {
      const int o = i / UFV_BITS_WORDS, w = i - o * UFV_BITS_WORDS;
      s_bits[o][w] = member_row[j] >= 0 ? bits[size_t(member_row[j]) * UFV_BITS_WORDS + w] : 0u;
    }
  }
  __syncthreads();
  if (tid < UFV_BITS_WORDS) {
    uint32_t u = 0;
#pragma unroll
    for (int o = 0; o < 8; ++o) u |= s_bits[o][tid];
    s_union[tid] = u;
  }
  __syncthreads();
  const int n_win = (n_patch + R - 1) / R;
  if (tid == 0) s_n_chunks = plan_frame(s_union, n_win, n_patch, use_tmap, tile_min, s_chunk);
  __syncthreads();
  const int n_chunks = s_n_chunks;                             // 0: direct path (one stage per non-empty window)
  const int slice_ch = min(kPoolCh, c - ch0);

  if (warp == kPoolConsumers) {
    // ---------------- producer warp: windows -> TMA engine -> shared-memory ring ------------------
    const int64_t row_base = int64_t(row) * n_patch;
    const uint32_t slice_bytes = uint32_t(slice_ch) * sizeof(T);
    if (n_chunks == 0) {
      int k = 0;                                               // stages used so far (non-empty windows)
      for (int win = 0; win < n_win; ++win) {
        const uint32_t u = s_union[win];
        if (u == 0u) continue;                                 // no member needs any row of this window
        const int s = k % S;
        const uint32_t ph = (k / S) & 1;
        ++k;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint32_t m = 0;                                        // lane r: which members pool patch 32 win + r
#pragma unroll
        for (int o = 0; o < 8; ++o) m |= ((s_bits[o][win] >> lane) & 1u) << o;
        s_omask[s][lane] = static_cast<uint8_t>(m);
        __syncwarp();                                          // the masks precede lane 0's (releasing) arrive
        produce_window<T>(&tmap, use_tmap, feats, row_base, c, ch0, slice_bytes, win, n_patch, u, tile_min,
                          ring + size_t(s) * R * kPoolCh, &full_bar[s], lane);
      }
    }
    for (int k = 0; k < n_chunks; ++k) {
      const PoolChunk ch = unpack_chunk(s_chunk[k]);
      const int s = k % S;
      const uint32_t ph = (k / S) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      // member masks of the stage's slots: which of the <= 8 members pool the patch sitting in slot i
      if (ch.tile) {
        uint32_t m = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) m |= ((s_bits[o][ch.w0] >> lane) & 1u) << o;
        s_omask[s][lane] = static_cast<uint8_t>(m);
      } else {
        s_omask[s][lane] = 0;
        __syncwarp();
        int base = 0;
        for (int w = ch.w0; w < ch.w1; ++w) {
          const uint32_t u = s_union[w];
          if ((u >> lane) & 1u) {
            uint32_t m = 0;
#pragma unroll
            for (int o = 0; o < 8; ++o) m |= ((s_bits[o][w] >> lane) & 1u) << o;
            s_omask[s][packed_slot(base, u, lane)] = static_cast<uint8_t>(m);
          }
          base += __popc(u);
        }
      }
      __syncwarp();                                            // the masks precede lane 0's (releasing) arrive
      produce_chunk<T>(&tmap, feats, row_base, c, ch0, slice_bytes, ch, s_union, ring + size_t(s) * R * kPoolCh,
                       &full_bar[s], lane);
    }
  } else {
    // ---------------- consumer warps: ascending-patch accumulation ---------------------------------
    float2 acc[OT];
#pragma unroll
    for (int o = 0; o < OT; ++o) acc[o] = make_float2(0.f, 0.f);
    const int my_ch = warp * (kPoolCh / kPoolConsumers) + lane * 2;      // within the slice
    const bool live = my_ch < slice_ch;
    // output rows and denominators: requested now, needed only after the last stage
    int out_row[OT];
    float denorm[OT];
#pragma unroll
    for (int o = 0; o < OT; ++o) {
      out_row[o] = o < n_mem ? grp_member[m0 + o] : -1;
      denorm[o] = out_row[o] >= 0 ? __fadd_rn(float(cnt[out_row[o]]), 1e-8f) : 1.0f;   // layer.py:145
    }
    int n_stages = n_chunks;
    if (n_chunks == 0)
      for (int win = 0; win < n_win; ++win) n_stages += s_union[win] != 0u;
    for (int k = 0; k < n_stages; ++k) {                        // the stage's member masks say everything a consumer needs
      const int s = k % S;
      const uint32_t ph = (k / S) & 1;
      mbar_wait(&full_bar[s], ph);
      const T* src = ring + size_t(s) * R * kPoolCh + my_ch;
      uint32_t mk[R / 4];
      Raw v[R];
#pragma unroll
      for (int i = 0; i < R / 4; ++i) mk[i] = reinterpret_cast<const uint32_t*>(&s_omask[s][0])[i];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = *reinterpret_cast<const Raw*>(src + r * kPoolCh);   // unused slots: stale bytes, mask 0
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);   // stage is in registers: hand it back early
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t m = (mk[r >> 2] >> (8 * (r & 3))) & 0xffu;
        const float2 f = Pair<T>::cvt(v[r]);
#pragma unroll
        for (int o = 0; o < OT; ++o)
          if (m & (1u << o)) add2(acc[o], f);
      }
    }
    if (live) {
#pragma unroll
      for (int o = 0; o < OT; ++o) {
        if (out_row[o] >= 0) {
          float2 out;
          out.x = __fdiv_rn(acc[o].x, denorm[o]);
          out.y = __fdiv_rn(acc[o].y, denorm[o]);
          *reinterpret_cast<float2*>(pooled + size_t(out_row[o]) * c + ch0 + my_ch) = out;
        }
      }
    }
  }
}

// VEC adjacent channels of one staged row -> fp32 pairs
template <typename T, int VEC> struct Lanes;
template <> struct Lanes<float, 4> {
  __device__ static void load(const float* p, float2 (&f)[2]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    f[0] = make_float2(r.x, r.y);
    f[1] = make_float2(r.z, r.w);
  }
};
template <> struct Lanes<__nv_bfloat16, 4> {
  __device__ static void load(const __nv_bfloat16* p, float2 (&f)[2]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    f[0] = Pair<__nv_bfloat16>::cvt(r.x);
    f[1] = Pair<__nv_bfloat16>::cvt(r.y);
  }
};
template <> struct Lanes<__half, 4> {
  __device__ static void load(const __half* p, float2 (&f)[2]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    f[0] = Pair<__half>::cvt(r.x);
    f[1] = Pair<__half>::cvt(r.y);
  }
};
template <> struct Lanes<__nv_bfloat16, 8> {
  __device__ static void load(const __nv_bfloat16* p, float2 (&f)[4]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    f[0] = Pair<__nv_bfloat16>::cvt(r.x);
    f[1] = Pair<__nv_bfloat16>::cvt(r.y);
    f[2] = Pair<__nv_bfloat16>::cvt(r.z);
    f[3] = Pair<__nv_bfloat16>::cvt(r.w);
  }
};
template <> struct Lanes<__half, 8> {
  __device__ static void load(const __half* p, float2 (&f)[4]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    f[0] = Pair<__half>::cvt(r.x);
    f[1] = Pair<__half>::cvt(r.y);
    f[2] = Pair<__half>::cvt(r.z);
    f[3] = Pair<__half>::cvt(r.w);
  }
};

// ---- many objects on a frame: bit-iterating consumers ------------------------------------------------------
// MPW = members per consumer warp, NW = consumer warps (member j belongs to warp j % NW, its slot there is j / NW):
// 2 x 8, 4 x 8, 8 x 8 for groups of up to 16 / 32 / 64 members.  A warp visits, per staged 32-row chunk and per member it owns, exactly the rows
// that member pools (set bits of the member word, ascending) -- the accumulation order per (object, channel) is
// the same ascending-patch sequence as in the dense kernel and in the oracle.
// CH = channels per CTA slice, CH / 32 per lane.  The kernel is bound by its instruction count, and of the 15
// instructions one (member, row) costs at 4 channels per lane, 8 find the row and close the loop: with 16-bit features
// and up to 32 members the slices are 256 channels wide (8 per lane, one 16-byte shared load: 21 instructions per
// 8 channels), and a frame's odd 128 channels (1152 = 4 x 256 + 128) go to a CTA that runs the 128-wide body.
template <typename T, int MPW, int NW, int CH, int S>
__device__ __forceinline__ void pool_sparse_body(const CUtensorMap& tmap, int use_tmap, const T* __restrict__ feats,
                                                 int n_patch, int c, int g, int ch0, const uint32_t* __restrict__ bits,
                                                 const int32_t* __restrict__ cnt, const int32_t* __restrict__ grp_row,
                                                 const int32_t* __restrict__ grp_off,
                                                 const int32_t* __restrict__ grp_member, int tile_min,
                                                 float* __restrict__ pooled, T* ring, uint32_t (*s_bits)[UFV_BITS_WORDS],
                                                 uint32_t* s_union, uint64_t* full_bar, uint64_t* empty_bar) {
  constexpr int R = kPoolRows, PM = MPW * NW, kSparseThreads = SparseCfg<MPW, NW>::kThreads;
  constexpr int VEC = CH / 32;                                            // channels per lane
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NW);
    }
    mbar_fence_init();
    if (use_tmap) tma_prefetch_desc(&tmap);
  }
  const int m0 = grp_off[g];
  const int n_mem = grp_off[g + 1] - m0;
  const int row = grp_row[g];
  __syncthreads();
  pdl_wait();                  // patch bitmasks and counts come from kernel 1
  pdl_launch_dependents();
  for (int i = tid; i < PM * UFV_BITS_WORDS; i += kSparseThreads) {
    const int o = i / UFV_BITS_WORDS, w = i - o * UFV_BITS_WORDS;
    s_bits[o][w] = o < n_mem ? bits[size_t(grp_member[m0 + o]) * UFV_BITS_WORDS + w] : 0u;
  }
  __syncthreads();
  if (tid < UFV_BITS_WORDS) {
    uint32_t u = 0;
    for (int o = 0; o < PM; ++o) u |= s_bits[o][tid];
    s_union[tid] = u;
  }
  __syncthreads();
  // (No packed path here: with 9 .. 64 objects on a frame the union is rarely thin, and the packing code's registers
  // would push this instruction-bound kernel's hot loop into local memory at the 32 registers seven CTAs per SM allow.)
  const int n_win = (n_patch + R - 1) / R;
  const int slice_ch = min(CH, c - ch0);

  if (warp == NW) {
    // ---------------- producer warp ---------------------------------------------------------------------
    const int64_t row_base = int64_t(row) * n_patch;
    const uint32_t slice_bytes = uint32_t(slice_ch) * sizeof(T);
    int k = 0;
    for (int win = 0; win < n_win; ++win) {
      const uint32_t u = s_union[win];
      if (u == 0u) continue;
      const int s = k % S;
      const uint32_t ph = (k / S) & 1;
      ++k;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      produce_window<T, CH>(&tmap, use_tmap, feats, row_base, c, ch0, slice_bytes, win, n_patch, u, tile_min,
                            ring + size_t(s) * R * CH, &full_bar[s], lane);
    }
  } else {
    // ---------------- consumer warps: members warp, warp + 8, ..., all CH channels, VEC per lane ---------------
    float2 acc[MPW][VEC / 2];
#pragma unroll
    for (int mi = 0; mi < MPW; ++mi)
#pragma unroll
      for (int v = 0; v < VEC / 2; ++v) acc[mi][v] = make_float2(0.f, 0.f);
    const int my_ch = lane * VEC;
    const bool live = my_ch < slice_ch;
    // one stage per non-empty window, patch 32 win + r in slot r -- the tight loop (this kernel is bound by its
    // instruction count)
    int k = 0;
    for (int win = 0; win < n_win; ++win) {
      if (s_union[win] == 0u) continue;
      const int s = k % S;
      const uint32_t ph = (k / S) & 1;
      ++k;
      mbar_wait(&full_bar[s], ph);
      const T* src = ring + size_t(s) * R * CH + my_ch;
#pragma unroll
      for (int mi = 0; mi < MPW; ++mi) {
        uint32_t word = s_bits[mi * NW + warp][win];         // warp-uniform: the rows of this window the member pools
        while (word != 0u) {
          const int r = __ffs(word) - 1;
          word &= word - 1u;
          float2 f[VEC / 2];
          Lanes<T, VEC>::load(src + r * CH, f);
#pragma unroll
          for (int v = 0; v < VEC / 2; ++v) add2(acc[mi][v], f[v]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
#pragma unroll
    for (int mi = 0; mi < MPW; ++mi) {
      const int j = mi * NW + warp;
      if (j < n_mem && live) {
        const int out_row = grp_member[m0 + j];
        const float denorm = __fadd_rn(float(cnt[out_row]), 1e-8f);   // layer.py:145
        float* dst = pooled + size_t(out_row) * c + ch0 + my_ch;
#pragma unroll
        for (int v = 0; v < VEC / 4; ++v) {
          float4 out;
          out.x = __fdiv_rn(acc[mi][2 * v].x, denorm);
          out.y = __fdiv_rn(acc[mi][2 * v].y, denorm);
          out.z = __fdiv_rn(acc[mi][2 * v + 1].x, denorm);
          out.w = __fdiv_rn(acc[mi][2 * v + 1].y, denorm);
          reinterpret_cast<float4*>(dst)[v] = out;
        }
      }
    }
  }
}

// WIDE: 256-channel slices; the CTA of a frame's last, 128-channel-or-narrower slice runs the 128-wide body on it
// (tmap_narrow: 128-channel boxes).  Work item = (group, slice): blockIdx.x = g * n_slices + slice.
template <typename T, int MPW, int NW, bool WIDE>
__global__ void __launch_bounds__(SparseCfg<MPW, NW>::kThreads,
                                  sizeof(T) == 4 ? (NW <= 8 ? 4 : 1) : WIDE ? SparseCfg<MPW, NW>::kWideCtas : SparseCfg<MPW, NW>::kMinCtas)
mask_pool_sparse_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_narrow,
                        int use_tmap, const T* __restrict__ feats, int n_patch, int c, int n_slices,
                        const uint32_t* __restrict__ bits, const int32_t* __restrict__ cnt,
                        const int32_t* __restrict__ grp_row, const int32_t* __restrict__ grp_off,
                        const int32_t* __restrict__ grp_member, int tile_min, float* __restrict__ pooled) {
  constexpr int S = WIDE ? SparseCfg<MPW, NW>::kWideStages : SparseCfg<MPW, NW>::kStages, PM = MPW * NW;
  constexpr int CH = WIDE ? 2 * kPoolCh : kPoolCh;
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  T* ring = reinterpret_cast<T*>(dyn_smem);                               // [S][R][CH]
  __shared__ uint32_t s_bits[PM][UFV_BITS_WORDS];                          // members' patch bitmasks
  __shared__ uint32_t s_union[UFV_BITS_WORDS];
  __shared__ __align__(8) uint64_t full_bar[S];
  __shared__ __align__(8) uint64_t empty_bar[S];
  const int g = blockIdx.x / n_slices;
  const int slice = blockIdx.x - g * n_slices;
  const int ch0 = slice * CH;
  if (WIDE && c - ch0 <= kPoolCh)
    pool_sparse_body<T, MPW, NW, kPoolCh, S>(tmap_narrow, use_tmap, feats, n_patch, c, g, ch0, bits, cnt, grp_row, grp_off,
                                          grp_member, tile_min, pooled, ring, s_bits, s_union, full_bar, empty_bar);
  else
    pool_sparse_body<T, MPW, NW, CH, S>(tmap, use_tmap, feats, n_patch, c, g, ch0, bits, cnt, grp_row, grp_off, grp_member,
                                     tile_min, pooled, ring, s_bits, s_union, full_bar, empty_bar);
}

// ---- host -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* tensor_map_encoder() {
  static void* fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return p;
  }();
  return fn;
}

// L2 promotion of the tiled loads: 256 B by default; UFV_TMAP_L2PROMO=0|64|128|256 overrides (developer sweeps).
static CUtensorMapL2promotion l2_promotion() {
  static const CUtensorMapL2promotion v = [] {
    const char* e = getenv("UFV_TMAP_L2PROMO");
    const int n = e ? atoi(e) : 256;
    return n == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : n == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
           : n == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }();
  return v;
}

// 2-D row-major [rows, cols] tensor map with a [box_rows, box_cols] box.  Encoding costs about a
// microsecond of host time and the same few maps (features, tokens, weights) come back call after
// call, so the last few are kept in a small per-thread table keyed by every encode parameter.
int make_tensor_map_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                       uint32_t box_rows, uint32_t box_cols, int swizzle128) {
  struct Entry {
    const void* base; uint64_t rows, cols; uint32_t box_rows, box_cols; int dtype, swizzle; bool valid;
    CUtensorMap map;
  };
  constexpr int kEntries = 16;
  static thread_local Entry cache[kEntries] = {};
  const size_t slot = ((reinterpret_cast<uintptr_t>(base) >> 8) ^ (box_rows * 7u) ^ cols) % kEntries;
  Entry& e = cache[slot];
  if (e.valid && e.base == base && e.rows == rows && e.cols == cols && e.box_rows == box_rows &&
      e.box_cols == box_cols && e.dtype == dtype && e.swizzle == swizzle128) {
    *map = e.map;
    return 0;
  }
  auto encode = reinterpret_cast<EncodeTiledFn>(tensor_map_encoder());
  if (encode == nullptr) return fail(UFV_E_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  const CUtensorMapDataType dt = dtype == UFV_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == UFV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                     : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * uint64_t(dtype_bytes(dtype))};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UFV_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  e = Entry{base, rows, cols, box_rows, box_cols, dtype, swizzle128, true, *map};
  return 0;
}

struct PoolArgs {
  const void* feats; int n_patch; int c; const uint32_t* bits; const int32_t* cnt; const int32_t* grp_row;
  const int32_t* grp_off; const int32_t* grp_member; int n_groups; int tile_min; float* pooled;
  int feat_dtype; uint64_t tmap_rows;
};

template <typename T, int OT>
static int launch_pool(const CUtensorMap& tmap, int use_tmap, const PoolArgs& a, cudaStream_t stream) {
  const int n_slices = (a.c + kPoolCh - 1) / kPoolCh;
  const size_t smem = size_t(kPoolStages) * kPoolRows * kPoolCh * sizeof(T);
  auto kernel = mask_pool_kernel<T, OT>;
  static bool configured = false;   // idempotent attribute; a benign race sets it twice
  if (!configured) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    configured = true;
  }
  return check_launch(
      "ufv_mask_pool",
      launch_kernel(kernel, dim3(unsigned(a.n_groups) * n_slices), dim3(kPoolThreads), smem, stream, tmap,
                    use_tmap, static_cast<const T*>(a.feats), a.n_patch, a.c, n_slices, a.bits, a.cnt, a.grp_row,
                    a.grp_off, a.grp_member, a.tile_min, a.pooled));
}

template <typename T, int MPW, int NW, bool WIDE>
static int launch_pool_sparse(const CUtensorMap& tmap, int use_tmap, const PoolArgs& a, cudaStream_t stream) {
  constexpr int CH = WIDE ? 2 * kPoolCh : kPoolCh;
  const int n_slices = (a.c + CH - 1) / CH;
  const size_t smem = size_t(WIDE ? SparseCfg<MPW, NW>::kWideStages : SparseCfg<MPW, NW>::kStages) * kPoolRows * CH * sizeof(T);
  constexpr int kSparseThreads = SparseCfg<MPW, NW>::kThreads;
  auto kernel = mask_pool_sparse_kernel<T, MPW, NW, WIDE>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    configured = true;
  }
  CUtensorMap wide = tmap;                      // tmap: 128-channel boxes (also the narrow last slice of a wide launch)
  if (WIDE && use_tmap) {
    const int rc = make_tensor_map_2d(&wide, a.feats, a.feat_dtype, a.tmap_rows, uint64_t(a.c), kPoolRows, CH, 0);
    if (rc != 0) return rc;
  }
  return check_launch(
      "ufv_mask_pool (many objects per frame)",
      launch_kernel(kernel, dim3(unsigned(a.n_groups) * n_slices), dim3(kSparseThreads), smem, stream, wide, tmap,
                    use_tmap, static_cast<const T*>(a.feats), a.n_patch, a.c, n_slices, a.bits, a.cnt, a.grp_row,
                    a.grp_off, a.grp_member, a.tile_min, a.pooled));
}

template <typename T>
static int dispatch_group(int max_group, const CUtensorMap& tmap, int use_tmap, const PoolArgs& a,
                          cudaStream_t stream) {
  if (max_group <= 4) return launch_pool<T, 4>(tmap, use_tmap, a, stream);
  if (max_group <= 8) return launch_pool<T, 8>(tmap, use_tmap, a, stream);
  // 256-channel slices for 16-bit features and up to 32 members (UFV_POOL_NARROW=1: developer A/B knob)
  static const bool wide_ok = getenv("UFV_POOL_NARROW") == nullptr;
  constexpr bool kHalf = sizeof(T) == 2;
  if (kHalf && wide_ok && a.c > kPoolCh) {
    if (max_group <= 16) return launch_pool_sparse<T, 2, 8, kHalf>(tmap, use_tmap, a, stream);
    if (max_group <= 32) return launch_pool_sparse<T, 4, 8, kHalf>(tmap, use_tmap, a, stream);
  }
  if (max_group <= 16) return launch_pool_sparse<T, 2, 8, false>(tmap, use_tmap, a, stream);
  if (max_group <= 32) return launch_pool_sparse<T, 4, 8, false>(tmap, use_tmap, a, stream);
  return launch_pool_sparse<T, 8, 8, false>(tmap, use_tmap, a, stream);
}

// ---- adjoint of the mask pool (training) --------------------------------------------
// pooled[j] = sum over on patches of feats[row_j, p] / denorm_j   =>
// d_feats[row, p, :] = sum over object-frames j on that row with patch p on of w[j, :],  w = d_pooled / denorm.
// One CTA per (feature row, range of 81 patches): a thread owns 4 channels of every patch row, so a CTA
// writes whole 2304-byte rows back to back -- one contiguous 186 KB stretch of d_feats (the first version
// wrote 256-byte pieces 2304 bytes apart and reached 2.5 TB/s).  The w values of up to 8 object-frames
// live in registers; rows with more object-frames re-read w through L1.  Rows nobody pools from are
// written as zeros, so the output needs no memset.  The sum runs in plan order: deterministic.
constexpr int kBwdThreads = 288;       // x 4 channels = 1152
constexpr int kBwdSplit = 9;           // patch ranges per feature row
constexpr int kBwdRegMembers = 8;

template <typename T>
__device__ __forceinline__ void store_quad(T* dst, float4 v);
template <> __device__ __forceinline__ void store_quad<float>(float* dst, float4 v) {
  *reinterpret_cast<float4*>(dst) = v;
}
template <> __device__ __forceinline__ void store_quad<__nv_bfloat16>(__nv_bfloat16* dst, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
template <> __device__ __forceinline__ void store_quad<__half>(__half* dst, float4 v) {
  const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

template <typename T>
__global__ void __launch_bounds__(kBwdThreads)
mask_pool_backward_kernel(const float* __restrict__ w, const uint32_t* __restrict__ bits,
                          const int32_t* __restrict__ row_off, const int32_t* __restrict__ row_member,
                          int n_patch, int c, T* __restrict__ d_feats) {
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(dyn_smem);        // [n_mem][UFV_BITS_WORDS]
  int32_t* s_member = reinterpret_cast<int32_t*>(dyn_smem) + 0;    // placed after the bits below
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x / kBwdSplit;
  const int part = blockIdx.x - row * kBwdSplit;
  const int per_part = (n_patch + kBwdSplit - 1) / kBwdSplit;
  const int p_begin = part * per_part, p_end = min(n_patch, p_begin + per_part);
  const int m0 = row_off[row];
  const int n_mem = row_off[row + 1] - m0;
  s_member = reinterpret_cast<int32_t*>(s_bits + size_t(n_mem) * UFV_BITS_WORDS);
  const int tid = threadIdx.x;
  for (int i = tid; i < n_mem; i += kBwdThreads) s_member[i] = row_member[m0 + i];
  for (int i = tid; i < n_mem * UFV_BITS_WORDS; i += kBwdThreads)
    s_bits[i] = bits[size_t(row_member[m0 + i / UFV_BITS_WORDS]) * UFV_BITS_WORDS + i % UFV_BITS_WORDS];
  __syncthreads();
  for (int ch = tid * 4; ch < c; ch += kBwdThreads * 4) {          // one pass for c <= 1152
    T* out = d_feats + size_t(row) * n_patch * c + ch;
    if (n_mem <= kBwdRegMembers) {
      float4 wv[kBwdRegMembers];
#pragma unroll
      for (int m = 0; m < kBwdRegMembers; ++m)
        wv[m] = m < n_mem ? *reinterpret_cast<const float4*>(w + size_t(s_member[m]) * c + ch)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = p_begin; p < p_end; ++p) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int word = p >> 5;
        const uint32_t bit = 1u << (p & 31);
#pragma unroll
        for (int m = 0; m < kBwdRegMembers; ++m) {
          if (m < n_mem && (s_bits[m * UFV_BITS_WORDS + word] & bit)) {      // CTA-uniform
            acc.x += wv[m].x; acc.y += wv[m].y; acc.z += wv[m].z; acc.w += wv[m].w;
          }
        }
        store_quad<T>(out + size_t(p) * c, acc);
      }
    } else {
      for (int p = p_begin; p < p_end; ++p) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int word = p >> 5;
        const uint32_t bit = 1u << (p & 31);
        for (int m = 0; m < n_mem; ++m) {
          if (s_bits[m * UFV_BITS_WORDS + word] & bit) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(w + size_t(s_member[m]) * c + ch));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        }
        store_quad<T>(out + size_t(p) * c, acc);
      }
    }
  }
}

template <typename T>
static int launch_pool_backward(const float* w, const uint32_t* bits, const int32_t* row_off,
                                const int32_t* row_member, int64_t n_rows, int max_members, int n_patch, int c,
                                void* d_feats, cudaStream_t stream) {
  const size_t smem = size_t(max_members > 0 ? max_members : 1) * (UFV_BITS_WORDS * 4 + 4);
  auto kernel = mask_pool_backward_kernel<T>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  return check_launch("ufv_mask_pool_backward",
                      launch_kernel(kernel, dim3(unsigned(n_rows) * kBwdSplit), dim3(kBwdThreads), smem, stream, w,
                                    bits, row_off, row_member, n_patch, c, static_cast<T*>(d_feats)));
}

}  // namespace ufv

extern "C" int ufv_mask_pool_backward(const float* w, const uint32_t* bits, const int32_t* row_off,
                                      const int32_t* row_member, int64_t n_rows, int max_members, int n_patch,
                                      int c, void* d_feats, int feat_dtype, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(n_rows >= 0 && max_members >= 0, UFV_E_SHAPE, "ufv_mask_pool_backward: negative size");
  if (n_rows == 0) return 0;
  UFV_REQUIRE(w && bits && row_off && row_member && d_feats, UFV_E_NULL, "ufv_mask_pool_backward: null pointer");
  UFV_REQUIRE(n_patch >= 1 && n_patch <= UFV_MAX_PATCH_SIDE * UFV_MAX_PATCH_SIDE && c >= 8 && c % 8 == 0,
              UFV_E_SHAPE, "ufv_mask_pool_backward: n_patch=%d c=%d", n_patch, c);
  UFV_REQUIRE(max_members <= 2000, UFV_E_SHAPE, "ufv_mask_pool_backward: %d object-frames on one feature row",
              max_members);
  UFV_REQUIRE(aligned16(w) && aligned16(d_feats), UFV_E_ALIGN, "ufv_mask_pool_backward: unaligned buffer");
  UFV_REQUIRE(n_rows * kBwdSplit < (int64_t(1) << 31), UFV_E_SHAPE, "ufv_mask_pool_backward: grid too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (feat_dtype) {
    case UFV_F32: return launch_pool_backward<float>(w, bits, row_off, row_member, n_rows, max_members, n_patch, c, d_feats, st);
    case UFV_BF16: return launch_pool_backward<__nv_bfloat16>(w, bits, row_off, row_member, n_rows, max_members, n_patch, c, d_feats, st);
    case UFV_F16: return launch_pool_backward<__half>(w, bits, row_off, row_member, n_rows, max_members, n_patch, c, d_feats, st);
    default: return fail(UFV_E_DTYPE, "ufv_mask_pool_backward: unsupported dtype %d", feat_dtype);
  }
}

extern "C" int ufv_mask_pool(const void* feats, int feat_dtype, int64_t n_rows, int n_patch, int c,
                             const uint32_t* bits, const int32_t* cnt, const int32_t* grp_row,
                             const int32_t* grp_off, const int32_t* grp_member, int n_groups, int max_group,
                             float* pooled_out, void* stream) {
  using namespace ufv;
  UFV_REQUIRE(n_groups >= 0 && n_rows >= 0, UFV_E_SHAPE, "ufv_mask_pool: negative size");
  if (n_groups == 0) return 0;
  UFV_REQUIRE(feats && bits && cnt && grp_row && grp_off && grp_member && pooled_out, UFV_E_NULL,
              "ufv_mask_pool: null pointer");
  UFV_REQUIRE(n_patch >= 1 && n_patch <= UFV_MAX_PATCH_SIDE * UFV_MAX_PATCH_SIDE, UFV_E_SHAPE,
              "ufv_mask_pool: n_patch=%d out of range", n_patch);
  UFV_REQUIRE(max_group >= 1 && max_group <= UFV_MAX_GROUP, UFV_E_SHAPE,
              "ufv_mask_pool: max_group=%d not in [1, %d]", max_group, UFV_MAX_GROUP);
  UFV_REQUIRE(feat_dtype == UFV_F32 || feat_dtype == UFV_BF16 || feat_dtype == UFV_F16, UFV_E_DTYPE,
              "ufv_mask_pool: unsupported feature dtype %d", feat_dtype);
  UFV_REQUIRE(c >= 8 && c % 8 == 0, UFV_E_SHAPE, "ufv_mask_pool: c=%d must be a multiple of 8", c);
  UFV_REQUIRE(aligned16(feats) && aligned16(pooled_out), UFV_E_ALIGN,
              "ufv_mask_pool: feats / pooled_out must be 16-byte aligned");
  UFV_REQUIRE(n_rows * n_patch < (int64_t(1) << 31), UFV_E_SHAPE,
              "ufv_mask_pool: n_rows * n_patch exceeds 2^31");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmap;
  int rc = make_tensor_map_2d(&tmap, feats, feat_dtype, uint64_t(n_rows) * n_patch, uint64_t(c), kPoolRows,
                              kPoolCh, 0);
  if (rc != 0) return rc;
  static const int use_tile = getenv("UFV_POOL_NO_TILE") == nullptr;   // developer knobs (sweeps)
  static const int tile_min = [] {
    const char* e = getenv("UFV_POOL_TILE_MIN");
    const int v = e ? atoi(e) : 20;
    return v < 1 ? 1 : v > 33 ? 33 : v;
  }();
  const PoolArgs a{feats, n_patch, c, bits, cnt, grp_row, grp_off, grp_member, n_groups, tile_min, pooled_out,
                   feat_dtype, uint64_t(n_rows) * n_patch};
  switch (feat_dtype) {
    case UFV_F32:
      return dispatch_group<float>(max_group, tmap, use_tile, a, st);
    case UFV_BF16:
      return dispatch_group<__nv_bfloat16>(max_group, tmap, use_tile, a, st);
    default:
      return dispatch_group<__half>(max_group, tmap, use_tile, a, st);
  }
}
