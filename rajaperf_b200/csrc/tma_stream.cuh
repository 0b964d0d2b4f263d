// tma_stream.cuh -- pieces shared by the TMA-staged streaming kernels (scan_tma.cu, indexlist_tma.cu):
// a 2-D tensor-map load, the swizzled 128-byte row read, and the driver entry point for cuTensorMapEncodeTiled.
#pragma once
#include "common.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace rpb_tma {

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// the same with an L2 eviction-priority policy (createpolicy) attached: streamed-once data should not linger in L2
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar,
                                                 unsigned long long policy)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               :: "r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_first()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// one compute thread: its row of 16 doubles out of the swizzled stage
__device__ __forceinline__ void load_row16(double (&v)[16], const double* stage, int row)
{
  const char* base = reinterpret_cast<const char*>(stage) + row * 128;
  const int sw = row & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double2 q = *reinterpret_cast<const double2*>(base + ((c ^ sw) << 4));
    v[2 * c] = q.x; v[2 * c + 1] = q.y;
  }
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline encode_fn_t get_encode()
{
  static encode_fn_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_fn_t>(p);
  }
  return fn;
}


// x viewed as rows of 16 doubles (128 bytes), boxes of 16 x box_rows, 128-byte swizzle
// L2 promotion of the tensor-map loads (experiment switch RPB200_TMA_L2PROMO = 0 none / 1 64 B / 2 128 B / 3 256 B)
inline CUtensorMapL2promotion l2_promotion()
{
  static int v = -1;
  if (v < 0) { const char* e = getenv("RPB200_TMA_L2PROMO"); v = e ? atoi(e) : 2; if (v < 0 || v > 3) v = 2; }      // 128 B: +1 % over 256 B on SCAN / INDEXLIST
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
       : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}
// experiment switch RPB200_TMA_EVICT_FIRST=1: attach an evict_first policy to the streamed tiles
inline int tma_evict_first()
{
  static int v = -1;
  if (v < 0) { const char* e = getenv("RPB200_TMA_EVICT_FIRST"); v = (e && atoi(e)) ? 1 : 0; }
  return v;
}

inline bool make_row_map(CUtensorMap* map, const double* x, int64_t rows, int box_rows)
{
  encode_fn_t encode = get_encode();
  if (!encode) return false;
  const cuuint64_t dims[2] = {16, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)(16 * sizeof(double))};
  const cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion(),
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace rpb_tma
