// common.cuh -- shared pieces of the Base_B200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rpb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "rajaperf_b200 kernels are written for sm_100a (B200) only"
#endif

#define RPB_CHECK(expr)                                   \
  do {                                                    \
    cudaError_t rpb_e_ = (expr);                          \
    if (rpb_e_ != cudaSuccess) return (int)rpb_e_;        \
  } while (0)

#define RPB_LAUNCH_CHECK() RPB_CHECK(cudaGetLastError())

// One launch tuning (reference analogue: the block_size tunings, GPUUtils.hpp:345-373).
struct rpb_tuning {
  int block_size;   // threads per CTA
  int ctas_per_sm;  // persistent grid = sm_count * ctas_per_sm; 0 = one tile per CTA
  int unroll;       // independent vector accesses in flight per thread
};

enum rpb_kernel_id {
  RPB_K_COPY = 0, RPB_K_MUL, RPB_K_ADD, RPB_K_TRIAD, RPB_K_DOT,
  RPB_K_REDUCE_SUM, RPB_K_SCAN, RPB_K_SORT, RPB_K_SORTPAIRS,
  RPB_K_MASS3DPA, RPB_K_DIFFUSION3DPA, RPB_K_CONVECTION3DPA, RPB_K_LTIMES,
  RPB_K_HALO_PACKING_FUSED, RPB_K_HALO_EXCHANGE_FUSED,
  RPB_K_INDEXLIST, RPB_K_POLYBENCH_GEMM,
  RPB_K_COUNT
};

#define RPB_MAX_PARTIALS 8192
#define RPB_MAX_STREAMS 64        // distinct streams one context can serve at the same time
#define RPB_PREALLOC_STREAMS 8    // scratch sets allocated by rpb200_create (no allocation on the first call of a stream)

// Everything a kernel writes besides the caller's arrays, ONE SET PER (context, stream): calls on different streams of one
// context never share a ticket, a partial, a look-back descriptor or an epoch (the reference gives every reducer its own
// scratch, GPUUtils.hpp:250-330).  All state is left re-armed by the kernel that used it, so nothing is cleared between calls,
// and the look-back epochs live in DEVICE memory (read at kernel start, committed by the last CTA to retire): a captured
// graph holding one scan replays correctly any number of times.
struct rpb_scratch {
  cudaStream_t  stream;            // the stream this set is attached to
  int           attached;
  void*         d_fixed;           // one allocation: partials | tickets | epochs | basis tables
  double*       d_partials;        // 2 x RPB_MAX_PARTIALS doubles (DOT, REDUCE_SUM)
  unsigned int* d_ticket;          // [0] DOT [1] REDUCE_SUM
  unsigned int* d_scan_ticket;     // [0..1] scan.cu [2..3] scan_tma.cu [4..5] indexlist.cu [6..7] indexlist_tma.cu
  unsigned long long* d_epoch;     // [0] scan [1] indexlist: epoch of the LAST COMPLETED call
  double*       d_basis_tables;    // 64 doubles: PA basis tables built per call by a 1-CTA prologue
  void*         d_scan_state;      // look-back descriptors, grown lazily (rpb200_scan_reserve pre-sizes)
  size_t        scan_state_bytes;
  void*         d_ilist_state;
  size_t        ilist_state_bytes;
};

struct rpb200_ctx {
  int device;
  int sm_count;
  rpb_tuning tune[RPB_K_COUNT];
  rpb_scratch slot[RPB_MAX_STREAMS];
  int         last_slot;           // the slot of the previous call (the common case: one stream)
  size_t      scan_reserve_bytes;  // rpb200_scan_reserve / rpb200_indexlist_reserve: minimum state size of every slot
  size_t      ilist_reserve_bytes;
  volatile int lock;               // slot table spin lock (the reference drives a context from one host thread)
};

// The scratch set of (ctx, stream); attaches a free set on the first call of a stream (allocating it if it is not one of
// the pre-allocated ones -- illegal while the stream is being captured: rpb200_stream_attach() it beforehand).
// Also checks that the calling thread's current device is the context's.  nullptr + *err on failure.
rpb_scratch* rpb_get_scratch(rpb200_ctx* ctx, cudaStream_t st, int* err);
int rpb_grow_state(void** d_state, size_t* bytes, size_t need, cudaStream_t st);

#define RPB_SCRATCH(sc, ctx, st)                                        \
  rpb_scratch* sc;                                                      \
  { int rpb_se_ = 0; sc = rpb_get_scratch((ctx), (st), &rpb_se_); if (!sc) return rpb_se_; }

// entry points that need no scratch still refuse to launch on another device than the context's
#define RPB_CHECK_DEVICE(ctx)                                           \
  do { int rpb_d_ = -1; RPB_CHECK(cudaGetDevice(&rpb_d_)); if (rpb_d_ != (ctx)->device) return RPB200_EDEVICE; } while (0)

// ---------------------------------------------------------------------------------
// 256-bit / 128-bit global accesses.  sm_100a has LDG/STG.256 (ld.global.v4.f64).
// Streaming data is touched once: bypass L1 allocation on loads and stores.
// ---------------------------------------------------------------------------------
struct __align__(32) dbl4 { double x, y, z, w; };

__device__ __forceinline__ dbl4 ldg256_stream(const double* p)
{
  dbl4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256_stream(double* p, const dbl4& v)
{
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ dbl4 ldg256(const double* p)
{
  dbl4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(double* p, const dbl4& v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ double2 ldg128_stream(const double* p)
{
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128_stream(double* p, const double2& v)
{
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};"
               :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------
// mbarrier + bulk-async (TMA) copies: the staging path that does not go through the LSU
// ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned int bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity)
{
  unsigned int ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16); completes tx bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned int bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

static inline bool rpb_aligned(const void* p, size_t a) { return (((uintptr_t)p) & (a - 1)) == 0; }

static inline cudaStream_t rpb_stream(rpb200_stream_t s) { return (cudaStream_t)s; }

// grid for a persistent tile loop
static inline int rpb_grid(const rpb200_ctx* ctx, const rpb_tuning& t, int64_t tiles)
{
  if (tiles < 1) tiles = 1;
  if (t.ctas_per_sm <= 0) return (int)(tiles > 0x7fffffff ? 0x7fffffff : tiles);
  int64_t g = (int64_t)ctx->sm_count * t.ctas_per_sm;
  return (int)(g < tiles ? g : tiles);
}
