// indexlist.cu -- Basic_INDEXLIST / Basic_INDEXLIST_3LOOP: stream compaction of the indices whose
// value is negative, single pass, for sm_100a.
//
// Replaces basic/INDEXLIST-Cuda.cpp:33-265 (a hand-rolled grid scan over cub::BlockScan with a
// striped->blocked exchange and a per-rep memset of the ready flags) and the three kernels + the
// (N+1)-entry Index_type `counts` temporary of basic/INDEXLIST_3LOOP-Cuda.cpp:60-128.  Both reference
// kernels define the same result (INDEXLIST.hpp:17-25, INDEXLIST_3LOOP.hpp:17-24):
//     list[0..len) = ascending { i : x[i] < 0.0 },  len = their number,
// so one kernel serves both:
//   * every thread owns 16 CONTIGUOUS doubles (four 256-bit loads) -> a 16-bit flag mask + popcount;
//   * warp-shuffle scan of the thread counts, one smem scan of the 16 warp totals;
//   * decoupled look-back over 8-byte {epoch | status | count} descriptors (epoch-tagged: nothing is
//     cleared between calls), tiles dealt by atomic ticket;
//   * the tile's indices are compacted in shared memory and leave as one contiguous, coalesced run
//     list[prefix .. prefix + tile_count);
//   * algorithmic traffic only: 8 B read per element + 4 B written per selected element; no `counts`.
#include "common.cuh"

namespace {

constexpr int IL_BLOCK = 512;
constexpr int IL_IPT = 16;
constexpr int IL_TILE = IL_BLOCK * IL_IPT;     // 8192 elements = 64 KB of x per tile

// descriptor: [63:34] epoch (30 bits)  [33:32] status  [31:0] count
constexpr unsigned long long IL_PARTIAL = 1ull, IL_INCLUSIVE = 2ull;
constexpr size_t IL_DESC_BYTES = 128;          // state per tile: the TMA path keeps one descriptor per 128-byte line

__device__ __forceinline__ unsigned long long il_ld(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void il_st(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// 30-bit descriptor tag of a 64-bit call counter: never 0 (fresh, zeroed state reads "not ready"); a stale descriptor could
// only alias after 2^30 - 1 further calls that all left it untouched
__device__ __forceinline__ unsigned long long il_tag(unsigned long long epoch)
{ return (epoch % 0x3fffffffull + 1ull) << 34; }

__global__ void __launch_bounds__(IL_BLOCK)
indexlist_kernel(const double* __restrict__ x, int* __restrict__ list, int64_t n, long long* __restrict__ d_len,
                 unsigned long long* __restrict__ desc, unsigned int* __restrict__ ticket,
                 unsigned long long* d_epoch, unsigned int num_tiles, int vector_ok)
{
  __shared__ int s_out[IL_TILE];
  __shared__ unsigned int s_warp[IL_BLOCK / 32];
  __shared__ unsigned int s_prefix;
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_epoch;

  // this call's epoch = 1 + the last completed call's, in DEVICE memory (committed by the last CTA to retire): a captured
  // launch replays with a fresh tag; read once per CTA before the first barrier, so every read happens-before the commit
  if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long*)d_epoch + 1ull;
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long tag = il_tag(epoch);

  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&ticket[0], 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    if (tile >= num_tiles) {
      if (threadIdx.x == 0) {       // the last CTA to draw a terminating ticket re-arms the counters
        const unsigned int gone = atomicAdd(&ticket[1], 1u);
        if (gone == gridDim.x - 1) { ticket[0] = 0u; ticket[1] = 0u; *d_epoch = epoch; }
      }
      break;
    }

    const int64_t base = (int64_t)tile * IL_TILE + (int64_t)threadIdx.x * IL_IPT;
    const bool full = vector_ok && (int64_t)(tile + 1) * IL_TILE <= n;

    unsigned int mask = 0;          // bit k: x[base + k] < 0.0   (INDEXLIST_CONDITIONAL)
    if (full) {
#pragma unroll
      for (int k = 0; k < IL_IPT / 4; ++k) {
        const dbl4 q = ldg256_stream(x + base + 4 * k);
        mask |= (q.x < 0.0 ? 1u : 0u) << (4 * k);
        mask |= (q.y < 0.0 ? 1u : 0u) << (4 * k + 1);
        mask |= (q.z < 0.0 ? 1u : 0u) << (4 * k + 2);
        mask |= (q.w < 0.0 ? 1u : 0u) << (4 * k + 3);
      }
    } else {
#pragma unroll
      for (int k = 0; k < IL_IPT; ++k)
        if (base + k < n && x[base + k] < 0.0) mask |= 1u << k;
    }
    const unsigned int cnt = __popc(mask);

    unsigned int inc = cnt;         // warp inclusive scan of the thread counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();

    const unsigned int wt = (lane < IL_BLOCK / 32) ? s_warp[lane] : 0u;
    unsigned int winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += up;
    }
    const unsigned int tile_total = __shfl_sync(0xffffffffu, winc, IL_BLOCK / 32 - 1);
    const unsigned int warp_excl = __shfl_sync(0xffffffffu, winc - wt, warp);

    // warp 0: publish the tile count, look back for the number of selected indices before this tile
    if (warp == 0) {
      unsigned int prefix = 0;
      if (tile == 0) {
        if (lane == 0) il_st(desc, tag | (IL_INCLUSIVE << 32) | tile_total);
      } else {
        if (lane == 0) il_st(desc + tile, tag | (IL_PARTIAL << 32) | tile_total);
        int64_t look = (int64_t)tile - 1;
        for (;;) {
          const int64_t idx = look - lane;
          unsigned long long w = tag | (IL_INCLUSIVE << 32);
          if (idx >= 0) {
            do { w = il_ld(desc + idx); } while ((w >> 34) != (tag >> 34) || ((w >> 32) & 3ull) == 0ull);
          }
          const unsigned int incl = __ballot_sync(0xffffffffu, ((w >> 32) & 3ull) == IL_INCLUSIVE);
          const int first = __ffs(incl) - 1;
          unsigned int c = (first < 0 || lane <= first) ? (unsigned int)w : 0u;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
          prefix += c;
          if (first >= 0) break;
          look -= 32;
        }
        if (lane == 0) il_st(desc + tile, tag | (IL_INCLUSIVE << 32) | (unsigned long long)(prefix + tile_total));
      }
      if (lane == 0) {
        s_prefix = prefix;
        if (tile == num_tiles - 1) *d_len = (long long)prefix + (long long)tile_total;   // m_len
      }
    }

    // compact this tile's indices in shared memory (ascending: thread order, then bit order)
    unsigned int at = warp_excl + (inc - cnt);
    const int i0 = (int)base;
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      s_out[at++] = i0 + k;
    }
    __syncthreads();

    int* __restrict__ out = list + s_prefix;
    for (unsigned int k = threadIdx.x; k < tile_total; k += IL_BLOCK) out[k] = s_out[k];
    __syncthreads();     // s_out / s_tile / s_prefix are rewritten by the next iteration
  }
}

}  // namespace

int rpb_indexlist_tma_try(rpb200_ctx* ctx, const double* x, int* list, int64_t n, long long* d_len, unsigned long long* d_desc,
                          size_t desc_bytes, unsigned int* d_ticket, unsigned long long* d_epoch, cudaStream_t st, int* handled);   // indexlist_tma.cu

extern "C" int rpb200_indexlist_reserve(rpb200_ctx* ctx, int64_t n)
{
  if (!ctx || n < 0) return RPB200_EINVAL;
  RPB_CHECK_DEVICE(ctx);
  const size_t need = IL_DESC_BYTES * (size_t)((n + IL_TILE - 1) / IL_TILE + 1);
  if (need > ctx->ilist_reserve_bytes) ctx->ilist_reserve_bytes = need;
  for (int i = 0; i < RPB_MAX_STREAMS; ++i) {      // every existing scratch set; later ones: rpb200_stream_attach / first use
    rpb_scratch* sc = &ctx->slot[i];
    if (!sc->d_fixed) continue;
    const int rc = rpb_grow_state(&sc->d_ilist_state, &sc->ilist_state_bytes, ctx->ilist_reserve_bytes, sc->attached ? sc->stream : nullptr);
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int rpb200_indexlist(rpb200_ctx* ctx, const double* x, int* list, int64_t n, int64_t* d_len,
                                rpb200_stream_t s)
{
  if (!ctx || n < 0 || !d_len || (n > 0 && (!x || !list))) return RPB200_EINVAL;
  if (n > 0x7fffffffll) return RPB200_EINVAL;          // Int_type index list (RPTypes.hpp:81)
  cudaStream_t st = rpb_stream(s);
  RPB_SCRATCH(sc, ctx, st);
  if (n == 0) { RPB_CHECK(cudaMemsetAsync(d_len, 0, sizeof(int64_t), st)); return 0; }
  const unsigned int tiles = (unsigned int)((n + IL_TILE - 1) / IL_TILE);
  {
    size_t need = IL_DESC_BYTES * ((size_t)tiles + 1);
    if (need < ctx->ilist_reserve_bytes) need = ctx->ilist_reserve_bytes;
    const int rc = rpb_grow_state(&sc->d_ilist_state, &sc->ilist_state_bytes, need, st);
    if (rc != 0) return rc;
  }
  unsigned long long* d_epoch = sc->d_epoch + 1;
  static_assert(sizeof(long long) == sizeof(int64_t), "Index_type");
  {   // large, aligned problems: the TMA-staged warp-specialised kernel (separate ticket pair: [6], [7])
    int handled = 0;
    const int rc = rpb_indexlist_tma_try(ctx, x, list, n, (long long*)d_len, (unsigned long long*)sc->d_ilist_state,
                                         sc->ilist_state_bytes, sc->d_scan_ticket + 6, d_epoch, st, &handled);
    if (rc != 0) return rc;
    if (handled) return 0;
  }
  const int cps = ctx->tune[RPB_K_INDEXLIST].ctas_per_sm > 0 ? ctx->tune[RPB_K_INDEXLIST].ctas_per_sm : 2;
  int grid = ctx->sm_count * cps;
  if ((unsigned int)grid > tiles) grid = (int)tiles;
  indexlist_kernel<<<grid, IL_BLOCK, 0, st>>>(x, list, n, (long long*)d_len, (unsigned long long*)sc->d_ilist_state,
                                              sc->d_scan_ticket + 4, d_epoch, tiles, rpb_aligned(x, 32) ? 1 : 0);
  RPB_LAUNCH_CHECK();
  return 0;
}
