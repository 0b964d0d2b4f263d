// scan.cu -- Algorithm_SCAN: exclusive prefix sum of doubles, single pass, for sm_100a.
//
// Replaces algorithm/SCAN-Cuda.cpp:34-188 + common/CudaGridScan.hpp (striped loads, two
// cub::BlockExchange transposes, a per-rep cudaMemsetAsync of the ready flags, an sm_70 tuning
// table, reliance on in-order block scheduling).  Here:
//   * every thread owns VPT*4 CONTIGUOUS doubles (256-bit loads/stores): thread-local serial scan,
//     one warp-shuffle scan of the thread totals, one tiny smem scan of the warp totals -- no
//     transposes, 2 barriers per tile;
//   * tiles are handed out by an atomic ticket (forward progress does not depend on how the
//     hardware orders CTAs); the last tile to retire re-arms the ticket;
//   * decoupled look-back over 16-byte {epoch|status, value} descriptors read/written with single
//     128-bit accesses; the epoch tag makes stale descriptors from earlier calls read as
//     "not ready", so nothing is cleared between calls;
//   * algorithmic traffic only: 8 B read + 8 B written per element.
#include "common.cuh"

namespace {

constexpr unsigned long long ST_PARTIAL = 1ull, ST_INCLUSIVE = 2ull;

struct __align__(16) tile_desc { unsigned long long word; double value; };

__device__ __forceinline__ void desc_store(tile_desc* p, unsigned long long word, double v)
{
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};"
               :: "l"(p), "l"(word), "l"(__double_as_longlong(v)) : "memory");
}
__device__ __forceinline__ void desc_load(const tile_desc* p, unsigned long long& word, double& v)
{
  long long bits;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];"
               : "=l"(word), "=l"(bits) : "l"(p) : "memory");
  v = __longlong_as_double(bits);
}

template <int VPT>   // vectors (of 4 doubles) per thread
__global__ void __launch_bounds__(512)
scan_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n,
            tile_desc* __restrict__ desc, unsigned int* __restrict__ ticket,
            unsigned long long* d_epoch, unsigned int num_tiles, int vector_ok, int dstride)
{
  // the epoch of this call = 1 + the epoch of the last completed call, kept in DEVICE memory and committed by the last
  // CTA to retire: a captured launch replays with a fresh epoch every time (no host-side counter frozen into the graph)
  // (read by ONE thread before the first barrier and broadcast through shared memory, so that every read of the epoch
  // happens-before any CTA's terminating ticket, hence before the commit)
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long*)d_epoch + 1ull;
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  // dstride: distance between tile descriptors in 16-byte units (see scan_tma.cu)
  constexpr int IPT = VPT * 4;
  __shared__ double s_warp[32];
  __shared__ double s_prefix;
  __shared__ unsigned int s_tile;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tile_elems = (int64_t)blockDim.x * IPT;

  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&ticket[0], 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    if (tile >= num_tiles) {
      // every CTA draws exactly one terminating ticket; the last one to do so re-arms the counters
      // for the next call (no CTA can draw again after that)
      if (threadIdx.x == 0) {
        const unsigned int gone = atomicAdd(&ticket[1], 1u);
        if (gone == gridDim.x - 1) { ticket[0] = 0u; ticket[1] = 0u; *d_epoch = epoch; }
      }
      break;
    }

    const int64_t base = (int64_t)tile * tile_elems + (int64_t)threadIdx.x * IPT;
    const bool full = vector_ok && (int64_t)(tile + 1) * tile_elems <= n;

    double v[IPT];
    if (full) {
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        dbl4 q = ldg256_stream(x + base + 4 * k);
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k) v[k] = (base + k < n) ? x[base + k] : 0.0;
    }

    // thread-local exclusive scan; `run` ends as the thread total
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) { const double t = v[k]; v[k] = run; run += t; }

    // warp inclusive scan of thread totals
    double inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();

    // every warp scans the (<=16) warp totals redundantly
    double wt = (lane < nwarps) ? s_warp[lane] : 0.0;
    double winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += up;
    }
    const double tile_total = __shfl_sync(0xffffffffu, winc, nwarps - 1);
    const double warp_excl = __shfl_sync(0xffffffffu, winc - wt, warp);

    // warp 0: publish the aggregate, look back for the exclusive prefix of this tile
    if (warp == 0) {
      double prefix = 0.0;
      if (tile == 0) {
        if (lane == 0) desc_store(desc, (epoch << 2) | ST_INCLUSIVE, tile_total);      // tile 0 sits at index 0 for any spacing
      } else {
        if (lane == 0) desc_store(desc + (int64_t)tile * dstride, (epoch << 2) | ST_PARTIAL, tile_total);
        int64_t look = (int64_t)tile - 1;
        for (;;) {
          const int64_t idx = look - lane;
          unsigned long long word = (epoch << 2) | ST_INCLUSIVE;
          double val = 0.0;
          if (idx >= 0) {
            do { desc_load(desc + idx * dstride, word, val); } while ((word >> 2) != epoch || (word & 3ull) == 0ull);
          }
          const unsigned int incl_mask = __ballot_sync(0xffffffffu, (word & 3ull) == ST_INCLUSIVE);
          const int first = __ffs(incl_mask) - 1;            // nearest tile holding an inclusive prefix
          const double contrib = (first < 0 || lane <= first) ? val : 0.0;
          prefix += warp_sum(contrib);
          if (first >= 0) break;
          look -= 32;
        }
        if (lane == 0) desc_store(desc + (int64_t)tile * dstride, (epoch << 2) | ST_INCLUSIVE, prefix + tile_total);
      }
      if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();

    const double off = s_prefix + warp_excl + (inc - run);
    if (full) {
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        dbl4 q;
        q.x = off + v[4 * k]; q.y = off + v[4 * k + 1]; q.z = off + v[4 * k + 2]; q.w = off + v[4 * k + 3];
        stg256_stream(y + base + 4 * k, q);
      }
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k) if (base + k < n) y[base + k] = off + v[k];
    }
    __syncthreads();   // s_tile / s_prefix are rewritten by the next iteration
  }
}

}  // namespace

int rpb_scan_tma_try(rpb200_ctx* ctx, const double* x, double* y, int64_t n, void* d_desc, size_t desc_bytes,
                     unsigned int* d_ticket, unsigned long long* d_epoch, cudaStream_t st, int* handled);   // scan_tma.cu

extern "C" int rpb200_scan_reserve(rpb200_ctx* ctx, int64_t n)
{
  if (!ctx || n < 0) return RPB200_EINVAL;
  RPB_CHECK_DEVICE(ctx);
  // the smallest tile any tuning can select (32 threads x 4 doubles) bounds the tile count
  const size_t tiles = (size_t)((n + 127) / 128);
  const size_t need = 2 * sizeof(tile_desc) * tiles;
  if (need > ctx->scan_reserve_bytes) ctx->scan_reserve_bytes = need;
  // every scratch set that exists is grown now; sets attached later are sized by rpb200_stream_attach / on first use
  for (int i = 0; i < RPB_MAX_STREAMS; ++i) {
    rpb_scratch* sc = &ctx->slot[i];
    if (!sc->d_fixed) continue;
    const int rc = rpb_grow_state(&sc->d_scan_state, &sc->scan_state_bytes, ctx->scan_reserve_bytes, sc->attached ? sc->stream : nullptr);
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int rpb200_scan_exclusive(rpb200_ctx* ctx, const double* x, double* y, int64_t n,
                                     rpb200_stream_t s)
{
  if (!ctx || n < 0 || (n > 0 && (!x || !y))) return RPB200_EINVAL;
  if (n == 0) return 0;
  cudaStream_t st = rpb_stream(s);
  RPB_SCRATCH(sc, ctx, st);
  rpb_tuning t = ctx->tune[RPB_K_SCAN];
  if (t.block_size > 512) t.block_size = 512;
  if (t.block_size < 32) t.block_size = 32;
  const int vpt = t.unroll >= 4 ? 4 : (t.unroll >= 2 ? 2 : 1);
  const bool aligned = rpb_aligned(x, 32) && rpb_aligned(y, 32);
  const int64_t tile_elems = (int64_t)t.block_size * vpt * 4;
  const int64_t tiles64 = (n + tile_elems - 1) / tile_elems;
  if (tiles64 > 0x7ffffff0ll) return RPB200_EINVAL;
  const unsigned int tiles = (unsigned int)tiles64;

  // the TMA path spreads its descriptors one per 128-byte line (8192-element tiles)
  size_t need = 2 * sizeof(tile_desc) * (size_t)tiles;        // descriptors one per 32-byte sector
  { const size_t spread = 128 * (size_t)((n + 8191) / 8192 + 1); if (spread > need) need = spread; }
  if (need < ctx->scan_reserve_bytes) need = ctx->scan_reserve_bytes;
  { const int rc = rpb_grow_state(&sc->d_scan_state, &sc->scan_state_bytes, need, st); if (rc != 0) return rc; }
  unsigned long long* d_epoch = sc->d_epoch + 0;

  // large, aligned problems: the TMA-staged warp-specialised kernel (separate ticket pair: [2], [3])
  {
    int handled = 0;
    const int rc = rpb_scan_tma_try(ctx, x, y, n, sc->d_scan_state, sc->scan_state_bytes, sc->d_scan_ticket + 2, d_epoch, st, &handled);
    if (rc != 0) return rc;
    if (handled) return 0;
  }

  int grid = ctx->sm_count * (t.ctas_per_sm > 0 ? t.ctas_per_sm : 4);
  if ((unsigned int)grid > tiles) grid = (int)tiles;
  tile_desc* desc = (tile_desc*)sc->d_scan_state;
  const int vok = aligned ? 1 : 0;   // unaligned sub-ranges take the bounds-checked scalar path
  int ds = 2;                        // one descriptor per 32-byte sector when the state allows it
  while (ds > 1 && sizeof(tile_desc) * (size_t)tiles * ds > sc->scan_state_bytes) ds >>= 1;
  switch (vpt) {
    case 4: scan_kernel<4><<<grid, t.block_size, 0, st>>>(x, y, n, desc, sc->d_scan_ticket, d_epoch, tiles, vok, ds); break;
    case 2: scan_kernel<2><<<grid, t.block_size, 0, st>>>(x, y, n, desc, sc->d_scan_ticket, d_epoch, tiles, vok, ds); break;
    default: scan_kernel<1><<<grid, t.block_size, 0, st>>>(x, y, n, desc, sc->d_scan_ticket, d_epoch, tiles, vok, ds); break;
  }
  RPB_LAUNCH_CHECK();
  return 0;
}
