// stream.cu -- Stream group (COPY, MUL, ADD, TRIAD, DOT) and Algorithm_REDUCE_SUM for sm_100a.
//
// Replaces stream/{COPY,MUL,ADD,TRIAD,DOT}-Cuda.cpp and algorithm/REDUCE_SUM-Cuda.cpp of the
// reference (one element per thread, 64-bit scalar accesses, tail predicate on every thread,
// smem tree + atomicAdd per block, host sync per rep).  Here:
//   * 256-bit LDG/STG (ld.global.v4.f64), L1 no-allocate: the data is touched once;
//   * each thread keeps UNROLL independent 32-byte loads per input in flight;
//   * tiles of BLOCK*UNROLL vectors are dealt round-robin to a (optionally persistent) grid,
//     full tiles run without any bounds test;
//   * MUL/TRIAD round the product and the sum separately (no FMA) => bit-identical to Base_Seq;
//   * DOT / REDUCE_SUM: per-thread vector accumulators -> warp shuffles -> one partial per CTA ->
//     the last CTA (ticket) folds the partials in a fixed order: deterministic, one launch,
//     result left on the device (no host round trip inside the rep loop).
#include "common.cuh"

enum { OP_COPY = 0, OP_MUL = 1, OP_ADD = 2, OP_TRIAD = 3 };

template <int OP>
__device__ __forceinline__ double ew_apply(double x, double y, double alpha)
{
  if (OP == OP_COPY) return x;
  if (OP == OP_MUL)  return __dmul_rn(alpha, x);
  if (OP == OP_ADD)  return __dadd_rn(x, y);
  return __dadd_rn(x, __dmul_rn(alpha, y));   // TRIAD: a = b + alpha*c, two roundings
}

template <int OP>
__device__ __forceinline__ dbl4 ew_apply4(const dbl4& x, const dbl4& y, double alpha)
{
  dbl4 r;
  r.x = ew_apply<OP>(x.x, y.x, alpha);
  r.y = ew_apply<OP>(x.y, y.y, alpha);
  r.z = ew_apply<OP>(x.z, y.z, alpha);
  r.w = ew_apply<OP>(x.w, y.w, alpha);
  return r;
}

// out[i] = f(in0[i], in1[i]);  nvec = n/4 vectors of 4 doubles; the n%4 tail is scalar.
template <int OP, int U>
__global__ void __launch_bounds__(512)
stream_ew_kernel(double* __restrict__ out, const double* __restrict__ in0,
                 const double* __restrict__ in1, double alpha, int64_t n)
{
  constexpr bool TWO = (OP == OP_ADD || OP == OP_TRIAD);
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)blockDim.x * U;          // vectors per tile
  const int64_t full_tiles = nvec / tile;

  for (int64_t t = blockIdx.x; t < full_tiles; t += gridDim.x) {
    const int64_t base = (t * tile + threadIdx.x) << 2;  // element index
    dbl4 x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = base + (((int64_t)u * blockDim.x) << 2);
      x[u] = ldg256_stream(in0 + e);
      if (TWO) y[u] = ldg256_stream(in1 + e);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = base + (((int64_t)u * blockDim.x) << 2);
      stg256_stream(out + e, ew_apply4<OP>(x[u], TWO ? y[u] : x[u], alpha));
    }
  }

  // ragged remainder: < one tile of vectors plus up to 3 scalars, owned by the CTA whose turn it is
  if ((int64_t)blockIdx.x == full_tiles % gridDim.x) {
    for (int64_t v = full_tiles * tile + threadIdx.x; v < nvec; v += blockDim.x) {
      const int64_t e = v << 2;
      dbl4 x = ldg256_stream(in0 + e), y = x;
      if (TWO) y = ldg256_stream(in1 + e);
      stg256_stream(out + e, ew_apply4<OP>(x, y, alpha));
    }
    const int64_t e = (nvec << 2) + threadIdx.x;
    if (e < n) out[e] = ew_apply<OP>(in0[e], TWO ? in1[e] : 0.0, alpha);
  }
}

// fallback for pointers that are not 32-byte aligned (e.g. an odd sub-range of an array)
template <int OP>
__global__ void __launch_bounds__(256)
stream_ew_scalar_kernel(double* __restrict__ out, const double* __restrict__ in0,
                        const double* __restrict__ in1, double alpha, int64_t n)
{
  constexpr bool TWO = (OP == OP_ADD || OP == OP_TRIAD);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ew_apply<OP>(in0[i], TWO ? in1[i] : 0.0, alpha);
}

template <int OP>
static int launch_ew(rpb200_ctx* ctx, int kid, double* out, const double* in0, const double* in1,
                     double alpha, int64_t n, rpb200_stream_t s)
{
  if (!ctx || n < 0 || (n > 0 && (!out || !in0))) return RPB200_EINVAL;
  if (n == 0) return 0;
  RPB_CHECK_DEVICE(ctx);
  cudaStream_t st = rpb_stream(s);
  constexpr bool two = (OP == OP_ADD || OP == OP_TRIAD);
  if (two && !in1) return RPB200_EINVAL;
  if (!rpb_aligned(out, 32) || !rpb_aligned(in0, 32) || (two && !rpb_aligned(in1, 32))) {
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 32;
    stream_ew_scalar_kernel<OP><<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(out, in0, in1, alpha, n);
    RPB_LAUNCH_CHECK();
    return 0;
  }
  rpb_tuning t = ctx->tune[kid];
  if (t.block_size > 512) t.block_size = 512;          // __launch_bounds__(512): 128 regs/thread
  int u = t.unroll >= 8 ? 8 : t.unroll >= 4 ? 4 : t.unroll >= 2 ? 2 : 1;
  if (two && u > 4) u = 4;                             // 2 inputs x 8 vectors would spill
  const int64_t tiles = ((n >> 2) + (int64_t)t.block_size * u - 1) / ((int64_t)t.block_size * u);
  const int grid = rpb_grid(ctx, t, tiles);
  switch (u) {
    case 8: stream_ew_kernel<OP, (two ? 4 : 8)><<<grid, t.block_size, 0, st>>>(out, in0, in1, alpha, n); break;
    case 4: stream_ew_kernel<OP, 4><<<grid, t.block_size, 0, st>>>(out, in0, in1, alpha, n); break;
    case 2: stream_ew_kernel<OP, 2><<<grid, t.block_size, 0, st>>>(out, in0, in1, alpha, n); break;
    default: stream_ew_kernel<OP, 1><<<grid, t.block_size, 0, st>>>(out, in0, in1, alpha, n); break;
  }
  RPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int rpb200_stream_copy(rpb200_ctx* ctx, double* c, const double* a, int64_t n, rpb200_stream_t s)
{ return launch_ew<OP_COPY>(ctx, RPB_K_COPY, c, a, nullptr, 0.0, n, s); }

extern "C" int rpb200_stream_mul(rpb200_ctx* ctx, double* b, const double* c, double alpha, int64_t n,
                                 rpb200_stream_t s)
{ return launch_ew<OP_MUL>(ctx, RPB_K_MUL, b, c, nullptr, alpha, n, s); }

extern "C" int rpb200_stream_add(rpb200_ctx* ctx, double* c, const double* a, const double* b, int64_t n,
                                 rpb200_stream_t s)
{ return launch_ew<OP_ADD>(ctx, RPB_K_ADD, c, a, b, 0.0, n, s); }

extern "C" int rpb200_stream_triad(rpb200_ctx* ctx, double* a, const double* b, const double* c,
                                   double alpha, int64_t n, rpb200_stream_t s)
{ return launch_ew<OP_TRIAD>(ctx, RPB_K_TRIAD, a, b, c, alpha, n, s); }

// Algorithm_MEMSET (algorithm/MEMSET-Cuda.cpp:27-76, MEMSET.hpp:27-28): x[i] = val -- a write-only stream, the
// in-suite ceiling for store bandwidth.  Algorithm_MEMCPY is rpb200_stream_copy.
__global__ void __launch_bounds__(512)
stream_set_kernel(double* __restrict__ out, double val, int64_t n, int64_t head)
{
  // `head` scalars bring the pointer to a 32-byte boundary; then whole vectors; then the tail
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (gtid < head) out[gtid] = val;
  double* __restrict__ body = out + head;
  const int64_t nvec = (n - head) >> 2;
  dbl4 v; v.x = v.y = v.z = v.w = val;
  for (int64_t i = gtid; i < nvec; i += stride) stg256_stream(body + (i << 2), v);
  const int64_t e = head + (nvec << 2) + gtid;
  if (e < n) out[e] = val;
}

extern "C" int rpb200_memset_f64(rpb200_ctx* ctx, double* x, double val, int64_t n, rpb200_stream_t s)
{
  if (!ctx || n < 0 || (n > 0 && !x)) return RPB200_EINVAL;
  if (n == 0) return 0;
  if (!rpb_aligned(x, 8)) return RPB200_EINVAL;
  int64_t head = (int64_t)((32 - ((uintptr_t)x & 31)) & 31) / 8;
  if (head > n) head = n;
  int64_t blocks = ((n >> 2) + 511) / 512;
  const int64_t cap = (int64_t)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  stream_set_kernel<<<(int)blocks, 512, 0, rpb_stream(s)>>>(x, val, n, head);
  RPB_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// Reductions: DOT (two inputs) and REDUCE_SUM (one input)
// =====================================================================================

// Folds a CTA's value into partials[blockIdx.x]; the last CTA to arrive (ticket) folds all
// partials in index order and writes the result.  Deterministic for a fixed grid.
template <int BLOCK_MAX_WARPS>
__device__ __forceinline__ void grid_fold(double v, double init, double* __restrict__ partials,
                                          unsigned int* __restrict__ ticket, double* __restrict__ out,
                                          int accumulate)
{
  __shared__ double s_warp[BLOCK_MAX_WARPS];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;

  v = warp_sum(v);
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double b = 0.0;
    for (int w = lane; w < nwarps; w += 32) b += s_warp[w];
    b = warp_sum(b);
    if (lane == 0) {
      partials[blockIdx.x] = b;
      __threadfence();
      const unsigned int done = atomicAdd(ticket, 1u);
      s_last = (done == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;

  __threadfence();
  // fixed-order fold of gridDim.x partials by this CTA
  double acc = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) acc += __ldcg(partials + i);
  acc = warp_sum(acc);
  __syncthreads();
  if (lane == 0) s_warp[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < nwarps; ++w) tot += s_warp[w];
    const double res = init + tot;
    *out = accumulate ? (*out + res) : res;
    *ticket = 0u;                    // ready for the next call on this stream
  }
}

template <int NIN, int U>
__global__ void __launch_bounds__(512)
reduce_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double init,
              double* __restrict__ partials, unsigned int* __restrict__ ticket,
              double* __restrict__ out, int accumulate)
{
  const int64_t nvec = n >> 2;
  const int64_t tile = (int64_t)blockDim.x * U;
  const int64_t full_tiles = nvec / tile;
  double acc[U][4];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.0;

  for (int64_t t = blockIdx.x; t < full_tiles; t += gridDim.x) {
    const int64_t base = (t * tile + threadIdx.x) << 2;
    dbl4 x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = base + (((int64_t)u * blockDim.x) << 2);
      x[u] = ldg256_stream(a + e);
      if (NIN == 2) y[u] = ldg256_stream(b + e);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (NIN == 2) {
        acc[u][0] = fma(x[u].x, y[u].x, acc[u][0]);
        acc[u][1] = fma(x[u].y, y[u].y, acc[u][1]);
        acc[u][2] = fma(x[u].z, y[u].z, acc[u][2]);
        acc[u][3] = fma(x[u].w, y[u].w, acc[u][3]);
      } else {
        acc[u][0] += x[u].x; acc[u][1] += x[u].y; acc[u][2] += x[u].z; acc[u][3] += x[u].w;
      }
    }
  }
  if ((int64_t)blockIdx.x == full_tiles % gridDim.x) {
    for (int64_t v = full_tiles * tile + threadIdx.x; v < nvec; v += blockDim.x) {
      const int64_t e = v << 2;
      dbl4 x = ldg256_stream(a + e);
      if (NIN == 2) {
        dbl4 y = ldg256_stream(b + e);
        acc[0][0] = fma(x.x, y.x, acc[0][0]); acc[0][1] = fma(x.y, y.y, acc[0][1]);
        acc[0][2] = fma(x.z, y.z, acc[0][2]); acc[0][3] = fma(x.w, y.w, acc[0][3]);
      } else {
        acc[0][0] += x.x; acc[0][1] += x.y; acc[0][2] += x.z; acc[0][3] += x.w;
      }
    }
    const int64_t e = (nvec << 2) + threadIdx.x;
    if (e < n) acc[0][0] += (NIN == 2) ? a[e] * b[e] : a[e];
  }
  double v = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u) v += (acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3]);
  grid_fold<32>(v, init, partials, ticket, out, accumulate);
}

template <int NIN>
__global__ void __launch_bounds__(256)
reduce_scalar_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double init,
                     double* __restrict__ partials, unsigned int* __restrict__ ticket,
                     double* __restrict__ out, int accumulate)
{
  double v = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    v += (NIN == 2) ? a[i] * b[i] : a[i];
  grid_fold<32>(v, init, partials, ticket, out, accumulate);
}

template <int NIN>
static int launch_reduce(rpb200_ctx* ctx, int kid, const double* a, const double* b, int64_t n,
                         double init, double* d_out, int accumulate, rpb200_stream_t s)
{
  if (!ctx || !d_out || n < 0 || (n > 0 && (!a || (NIN == 2 && !b)))) return RPB200_EINVAL;
  cudaStream_t st = rpb_stream(s);
  RPB_SCRATCH(sc, ctx, st);                    // this stream's partials + ticket: concurrent streams never share them
  unsigned int* ticket = sc->d_ticket + (NIN == 2 ? 0 : 1);
  double* partials = sc->d_partials + (NIN == 2 ? 0 : RPB_MAX_PARTIALS);
  rpb_tuning t = ctx->tune[kid];
  if (!rpb_aligned(a, 32) || (NIN == 2 && !rpb_aligned(b, 32))) {
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    int grid = (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
    reduce_scalar_kernel<NIN><<<grid, 256, 0, st>>>(a, b, n, init, partials, ticket, d_out, accumulate);
    RPB_LAUNCH_CHECK();
    return 0;
  }
  if (t.block_size > 512) t.block_size = 512;
  int u = t.unroll >= 8 ? 8 : t.unroll >= 4 ? 4 : t.unroll >= 2 ? 2 : 1;
  if (NIN == 2 && u > 4) u = 4;
  const int64_t tiles = ((n >> 2) + (int64_t)t.block_size * u - 1) / ((int64_t)t.block_size * u);
  int grid = rpb_grid(ctx, t, tiles);
  if (grid > RPB_MAX_PARTIALS) grid = RPB_MAX_PARTIALS;
  switch (u) {
    case 8: reduce_kernel<NIN, (NIN == 1 ? 8 : 4)><<<grid, t.block_size, 0, st>>>(a, b, n, init, partials, ticket, d_out, accumulate); break;
    case 4: reduce_kernel<NIN, 4><<<grid, t.block_size, 0, st>>>(a, b, n, init, partials, ticket, d_out, accumulate); break;
    case 2: reduce_kernel<NIN, 2><<<grid, t.block_size, 0, st>>>(a, b, n, init, partials, ticket, d_out, accumulate); break;
    default: reduce_kernel<NIN, 1><<<grid, t.block_size, 0, st>>>(a, b, n, init, partials, ticket, d_out, accumulate); break;
  }
  RPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int rpb200_stream_dot(rpb200_ctx* ctx, const double* a, const double* b, int64_t n, double init,
                                 double* d_out, int accumulate, rpb200_stream_t s)
{ return launch_reduce<2>(ctx, RPB_K_DOT, a, b, n, init, d_out, accumulate, s); }

extern "C" int rpb200_reduce_sum(rpb200_ctx* ctx, const double* x, int64_t n, double init, double* d_out,
                                 rpb200_stream_t s)
{ return launch_reduce<1>(ctx, RPB_K_REDUCE_SUM, x, nullptr, n, init, d_out, 0, s); }
