// Kernel 2: segmented mask pool.
//
// Replaces, for every object-frame, the reference's gather feats[ann_index], the NCHW permute,
// the fp32 upcast (ufvideo/model/layer.py:98-104) and the masked mean
// (x * mask / denorm).sum(-1).sum(-1) (layer.py:145-147), which materialise three dense fp32
// [q, C, 27, 27] temporaries.  Here each needed feature row is read from HBM once per frame
// and shared by every object on that frame.
//
// Work item = (group, 128-channel slice).  A group is one feature row (frame) plus up to OT
// object-frames that are pooled from it.  Per CTA: a producer warp streams the slice of every
// patch row in the union of the group's masks through a ring of shared-memory stages with the
// TMA engine (one 2-D tiled tensor-map load when the stage's rows are consecutive patches,
// otherwise one 1-D bulk copy per row -- off patches are never fetched); a consumer warp owns
// 4 channels per lane and adds each staged row, in ascending patch order, into the fp32
// accumulators of the objects whose bit is set (warp-uniform test, packed f32x2 adds).
// Accumulation order per (object, channel) is the plain ascending-patch sequence, independent
// of any blocking, which is what oracle/restatement.py::mask_pool restates bit-for-bit.
//
// Roofline: HBM.  Algorithmic bytes per group = n_union_patches * C * sizeof(feat).
#include "common.cuh"

#include <cuda.h>

namespace ufv {

constexpr int kPoolCh = 128;        // channels per CTA slice (32 lanes x 4)
constexpr int kPoolThreads = 64;    // warp 0 = consumer, warp 1 = producer
constexpr int kPoolStages = 4;
constexpr int kMaxPatches = UFV_BITS_WORDS * 32;

__device__ __forceinline__ void add2(float2& acc, float2 v) {
  unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long b = *reinterpret_cast<unsigned long long*>(&v);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  acc = *reinterpret_cast<float2*>(&a);
}

template <typename T> __device__ __forceinline__ void load4(const T* p, float2& lo, float2& hi);
template <> __device__ __forceinline__ void load4<float>(const float* p, float2& lo, float2& hi) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  lo = make_float2(v.x, v.y);
  hi = make_float2(v.z, v.w);
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float2& lo, float2& hi) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);   // bf16 -> fp32 is a 16-bit shift
  lo = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u));
  hi = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
}
template <> __device__ __forceinline__ void load4<__half>(const __half* p, float2& lo, float2& hi) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
  hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
}

template <typename T, int OT, int R>
__global__ void __launch_bounds__(kPoolThreads)
mask_pool_kernel(const __grid_constant__ CUtensorMap tmap, int use_tmap, const T* __restrict__ feats,
                 int n_patch, int c, int n_slices, const uint32_t* __restrict__ bits,
                 const int32_t* __restrict__ cnt, const int32_t* __restrict__ grp_row,
                 const int32_t* __restrict__ grp_off, const int32_t* __restrict__ grp_member,
                 float* __restrict__ pooled) {
  constexpr int S = kPoolStages;
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  T* ring = reinterpret_cast<T*>(dyn_smem);                        // [S][R][kPoolCh]
  __shared__ __align__(8) uint64_t full_bar[S];
  __shared__ __align__(8) uint64_t empty_bar[S];
  __shared__ uint32_t s_bm[OT][UFV_BITS_WORDS];
  __shared__ uint32_t s_union[UFV_BITS_WORDS];
  __shared__ int32_t s_prefix[UFV_BITS_WORDS + 1];
  __shared__ int32_t s_member[OT];
  __shared__ uint16_t s_ulist[kMaxPatches];   // union patch indices, ascending
  __shared__ uint8_t s_omask[kMaxPatches];    // bit o = object o pools this union patch

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = blockIdx.x / n_slices;
  const int slice = blockIdx.x - g * n_slices;
  const int ch0 = slice * kPoolCh;
  const int m0 = grp_off[g];
  const int n_mem = grp_off[g + 1] - m0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_fence_init();
  }
  if (tid == 32 && use_tmap) tma_prefetch_desc(&tmap);
  if (tid < OT) s_member[tid] = tid < n_mem ? grp_member[m0 + tid] : -1;
  __syncthreads();
  for (int i = tid; i < OT * UFV_BITS_WORDS; i += kPoolThreads) {
    const int o = i / UFV_BITS_WORDS, w = i - o * UFV_BITS_WORDS;
    const int j = s_member[o];
    s_bm[o][w] = j >= 0 ? bits[size_t(j) * UFV_BITS_WORDS + w] : 0u;
  }
  __syncthreads();
  if (tid < 32) {
    uint32_t u = 0;
    if (tid < UFV_BITS_WORDS) {
#pragma unroll
      for (int o = 0; o < OT; ++o) u |= s_bm[o][tid];
      s_union[tid] = u;
    }
    const int mine = __popc(u);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, d);
      if (tid >= d) incl += up;
    }
    if (tid < UFV_BITS_WORDS) s_prefix[tid] = incl - mine;
    if (tid == UFV_BITS_WORDS - 1) s_prefix[UFV_BITS_WORDS] = incl;
  }
  __syncthreads();
  for (int p = tid; p < kMaxPatches; p += kPoolThreads) {
    const int w = p >> 5;
    const uint32_t bit = 1u << (p & 31);
    const uint32_t u = s_union[w];
    if (u & bit) {
      const int pos = s_prefix[w] + __popc(u & (bit - 1u));
      uint32_t m = 0;
#pragma unroll
      for (int o = 0; o < OT; ++o) m |= ((s_bm[o][w] >> (p & 31)) & 1u) << o;
      s_ulist[pos] = static_cast<uint16_t>(p);
      s_omask[pos] = static_cast<uint8_t>(m);
    }
  }
  __syncthreads();
  const int n_u = s_prefix[UFV_BITS_WORDS];
  const int n_chunks = (n_u + R - 1) / R;
  const int64_t row_base = int64_t(grp_row[g]) * n_patch;
  const int slice_ch = min(kPoolCh, c - ch0);

  if (tid >= 32) {
    // ---------------- producer warp: TMA engine -> shared-memory ring -------------------------
    const uint32_t slice_bytes = uint32_t(slice_ch) * sizeof(T);
    for (int it = 0; it < n_chunks; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const int rows = min(R, n_u - it * R);
      const int first = s_ulist[it * R];
      const bool tile = use_tmap && rows == R && (int(s_ulist[it * R + R - 1]) - first == R - 1);
      T* dst = ring + size_t(s) * R * kPoolCh;
      if (tile) {
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], uint32_t(R) * kPoolCh * sizeof(T));
          tma_load_2d(dst, &tmap, ch0, int(row_base + first), &full_bar[s]);
        }
      } else {
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], uint32_t(rows) * slice_bytes);
        __syncwarp();
        for (int r = lane; r < rows; r += 32) {
          const T* src = feats + (row_base + s_ulist[it * R + r]) * int64_t(c) + ch0;
          bulk_g2s(dst + r * kPoolCh, src, slice_bytes, &full_bar[s]);
        }
      }
    }
  } else {
    // ---------------- consumer warp: ascending-patch accumulation --------------------------------
    float2 acc[OT][2];
#pragma unroll
    for (int o = 0; o < OT; ++o) acc[o][0] = acc[o][1] = make_float2(0.f, 0.f);
    const bool live = lane * 4 < slice_ch;
    for (int it = 0; it < n_chunks; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait(&full_bar[s], ph);
      const int rows = min(R, n_u - it * R);
      const T* src = ring + size_t(s) * R * kPoolCh + lane * 4;
      const uint8_t* om = s_omask + it * R;
      if (live) {
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
          const uint32_t m = om[r];
          float2 lo, hi;
          load4<T>(src + r * kPoolCh, lo, hi);
#pragma unroll
          for (int o = 0; o < OT; ++o) {
            if (m & (1u << o)) {
              add2(acc[o][0], lo);
              add2(acc[o][1], hi);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    if (live) {
#pragma unroll
      for (int o = 0; o < OT; ++o) {
        const int j = s_member[o];
        if (j >= 0) {
          const float denorm = __fadd_rn(float(cnt[j]), 1e-8f);   // layer.py:145
          float4 out;
          out.x = __fdiv_rn(acc[o][0].x, denorm);
          out.y = __fdiv_rn(acc[o][0].y, denorm);
          out.z = __fdiv_rn(acc[o][1].x, denorm);
          out.w = __fdiv_rn(acc[o][1].y, denorm);
          *reinterpret_cast<float4*>(pooled + size_t(j) * c + ch0 + lane * 4) = out;
        }
      }
    }
  }
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

// 2-D row-major [rows, cols] tensor map with a [box_rows, box_cols] box.
int make_tensor_map_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                       uint32_t box_rows, uint32_t box_cols, int swizzle128) {
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
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UFV_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return 0;
}

template <typename T, int OT, int R>
static int launch_pool(const CUtensorMap& tmap, int use_tmap, const void* feats, int n_patch, int c,
                       const uint32_t* bits, const int32_t* cnt, const int32_t* grp_row,
                       const int32_t* grp_off, const int32_t* grp_member, int n_groups,
                       float* pooled, cudaStream_t stream) {
  const int n_slices = (c + kPoolCh - 1) / kPoolCh;
  const size_t smem = size_t(kPoolStages) * R * kPoolCh * sizeof(T);
  auto kernel = mask_pool_kernel<T, OT, R>;
  static bool configured = false;   // idempotent attribute; a benign race sets it twice
  if (!configured) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    configured = true;
  }
  kernel<<<unsigned(n_groups) * n_slices, kPoolThreads, smem, stream>>>(
      tmap, use_tmap, static_cast<const T*>(feats), n_patch, c, n_slices, bits, cnt, grp_row,
      grp_off, grp_member, pooled);
  return check_launch("ufv_mask_pool");
}

template <typename T, int R>
static int dispatch_group(int max_group, const CUtensorMap& tmap, int use_tmap, const void* feats,
                          int n_patch, int c, const uint32_t* bits, const int32_t* cnt,
                          const int32_t* grp_row, const int32_t* grp_off, const int32_t* grp_member,
                          int n_groups, float* pooled, cudaStream_t stream) {
  if (max_group <= 4)
    return launch_pool<T, 4, R>(tmap, use_tmap, feats, n_patch, c, bits, cnt, grp_row, grp_off,
                                grp_member, n_groups, pooled, stream);
  return launch_pool<T, 8, R>(tmap, use_tmap, feats, n_patch, c, bits, cnt, grp_row, grp_off,
                              grp_member, n_groups, pooled, stream);
}

}  // namespace ufv

extern "C" int ufv_mask_pool(const void* feats, int feat_dtype, int64_t n_rows, int n_patch, int c,
                             const uint32_t* bits, const int32_t* cnt, const int32_t* grp_row,
                             const int32_t* grp_off, const int32_t* grp_member, int n_groups,
                             int max_group, float* pooled_out, void* stream) {
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
  constexpr int R16 = 32, R32 = 16;   // 8 KiB per stage either way
  CUtensorMap tmap;
  const int r = feat_dtype == UFV_F32 ? R32 : R16;
  int rc = make_tensor_map_2d(&tmap, feats, feat_dtype, uint64_t(n_rows) * n_patch, uint64_t(c), r,
                              kPoolCh, 0);
  if (rc != 0) return rc;
  switch (feat_dtype) {
    case UFV_F32:
      return dispatch_group<float, R32>(max_group, tmap, 1, feats, n_patch, c, bits, cnt, grp_row,
                                        grp_off, grp_member, n_groups, pooled_out, st);
    case UFV_BF16:
      return dispatch_group<__nv_bfloat16, R16>(max_group, tmap, 1, feats, n_patch, c, bits, cnt,
                                                grp_row, grp_off, grp_member, n_groups, pooled_out, st);
    default:
      return dispatch_group<__half, R16>(max_group, tmap, 1, feats, n_patch, c, bits, cnt, grp_row,
                                         grp_off, grp_member, n_groups, pooled_out, st);
  }
}
