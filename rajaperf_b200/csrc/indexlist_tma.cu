// indexlist_tma.cu -- Basic_INDEXLIST, large-n path: persistent, warp-specialised, TMA-staged compaction.
//
// Same machine as scan_tma.cu (one persistent CTA per SM, x streamed through a 3-stage 128-byte-swizzled
// shared-memory ring by cp.async.bulk.tensor, dedicated look-back warps, mbarrier hand-offs only), with
// the per-tile work of the index list.  A tile is only 64 KB of reads (1.5 us of one SM's share of HBM),
// less than one look-back round trip, so the roles are split further than in the scan: warp 0 only
// produces (full/empty ring), THREE look-back warps take every third tile, and the compute warps run two
// tiles ahead of the one they are waiting for:
//   * A(k): a compute thread reads its 16 contiguous doubles from the ring and keeps ONE register of them,
//     the 16-bit mask of x[i] < 0.0; warp-shuffle scan of the popcounts; the stage is free again as soon
//     as the 16 warps have done this (the look-back never holds a stage);
//   * a look-back warp resolves the tile's exclusive prefix by decoupled look-back over 8-byte
//     {epoch | status | count} descriptors and hands every warp its output offset;
//   * B(k), software-pipelined behind A(k+1) and A(k+2): each warp compacts its (<= 512) selected indices in its own
//     2 KB shared-memory slice (__syncwarp only) and writes them as one contiguous, coalesced run.
// Traffic stays algorithmic: 8 B read per element + 4 B written per selected element.
#include "common.cuh"
#include "tma_stream.cuh"

#include <stdlib.h>

namespace {

constexpr int IT_WARPS = 16;
constexpr int IT_SLOTS = 3;                        // tiles a CTA may have between A (count) and B (write): look-back depth
constexpr int IT_THREADS = (IT_WARPS + 1 + IT_SLOTS) * 32;    // + producer warp 0 + look-back warps 17..19
constexpr int IT_IPT = 16;
constexpr int IT_ROWS = IT_WARPS * 32;            // 512 rows of 16 doubles
constexpr int IT_TILE = IT_ROWS * IT_IPT;         // 8192 elements = 64 KiB
constexpr int IT_STAGES = 3;
constexpr int IT_BOX_ROWS = 256;
constexpr unsigned int IT_INVALID = 0xffffffffu;
constexpr unsigned long long IT_PARTIAL = 1ull, IT_INCLUSIVE = 2ull;   // descriptor: [63:34] epoch [33:32] status [31:0] count

__device__ __forceinline__ unsigned long long it_ld(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void it_st(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

struct il_smem {
  alignas(1024) double tile[IT_STAGES][IT_TILE];
  int slice[IT_WARPS][32 * IT_IPT];                  // per-warp compaction buffer
  unsigned long long full[IT_STAGES];
  unsigned long long empty[IT_STAGES];               // 16 compute warps are done reading the stage
  unsigned long long agg_ready[IT_SLOTS];
  unsigned long long prefix_ready[IT_SLOTS];
  unsigned int wtot[IT_SLOTS][IT_WARPS];
  unsigned int woff[IT_SLOTS][IT_WARPS];
  unsigned int tile_id[IT_STAGES];
  unsigned int lb_tile[IT_SLOTS];                           // tile of the sequence number a look-back warp is handed
  unsigned int arrived[IT_SLOTS];
  unsigned long long epoch;                                 // this call's epoch, read once per CTA
};

// IT_LBW = descriptors per lane per look-back round; dstride = distance between descriptors in 8-byte words
// (16 = one descriptor per 128-byte line, so the polls of a round spread over 32 L2 slices instead of 2 lines)
template <int IT_LBW>
__global__ void __launch_bounds__(IT_THREADS, 1)
indexlist_tma_kernel(const __grid_constant__ CUtensorMap x_map, int* __restrict__ list, long long rows,
                     unsigned long long* __restrict__ desc, unsigned int* __restrict__ ticket, unsigned long long* d_epoch,
                     unsigned int num_tiles, int dstride, unsigned int backoff_ns, int num_lb, int evict_first)
{
  extern __shared__ unsigned char smem_raw[];
  il_smem& S = *reinterpret_cast<il_smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this call's epoch = 1 + the last COMPLETED call's, kept in device memory and committed by the last producer to retire
  // (a captured launch replays with a fresh tag).  Thread 0 reads it before the barrier below and broadcasts it through
  // shared memory: no producer draws a ticket before every warp of its CTA can see S.epoch, so every read of *d_epoch
  // happens-before the commit.
  if (threadIdx.x == 0) S.epoch = *(volatile unsigned long long*)d_epoch + 1ull;

  if (threadIdx.x == 0) {
    for (int s = 0; s < IT_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], IT_WARPS); }
    for (int s = 0; s < IT_SLOTS; ++s) { mbar_init(&S.agg_ready[s], IT_WARPS); mbar_init(&S.prefix_ready[s], 1); S.arrived[s] = 0u; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long epoch = S.epoch;
  const unsigned long long tag = (epoch % 0x3fffffffull + 1ull) << 34;      // 30-bit tag, never 0 (indexlist.cu: il_tag)

  if (warp == 0) {
    // ---------------------------------------------------------------- producer (one lane)
    if (lane == 0) {
      int invalid_left = 1;             // every CTA draws exactly one terminating ticket
      for (int seq = 0; invalid_left > 0; ++seq) {
        const int st = seq % IT_STAGES;
        if (seq >= IT_STAGES) mbar_wait(&S.empty[st], ((seq / IT_STAGES) - 1) & 1);
        unsigned int t = atomicAdd(&ticket[0], 1u);
        if (t >= num_tiles) t = IT_INVALID;
        S.tile_id[st] = t;
        if (t != IT_INVALID) {
          const long long row0 = (long long)t * IT_ROWS;
          const bool second = row0 + IT_BOX_ROWS < rows;
          mbar_arrive_expect_tx(&S.full[st], (second ? 2u : 1u) * IT_BOX_ROWS * 128u);
          if (evict_first) {
            const unsigned long long pol = rpb_tma::policy_evict_first();
            rpb_tma::tma_load_2d_hint(&S.tile[st][0], &x_map, 0, (int)row0, &S.full[st], pol);
            if (second) rpb_tma::tma_load_2d_hint(&S.tile[st][IT_BOX_ROWS * IT_IPT], &x_map, 0, (int)(row0 + IT_BOX_ROWS), &S.full[st], pol);
          } else {
            rpb_tma::tma_load_2d(&S.tile[st][0], &x_map, 0, (int)row0, &S.full[st]);
            if (second) rpb_tma::tma_load_2d(&S.tile[st][IT_BOX_ROWS * IT_IPT], &x_map, 0, (int)(row0 + IT_BOX_ROWS), &S.full[st]);
          }
        } else {
          --invalid_left;
          mbar_arrive(&S.full[st]);
        }
      }
      const unsigned int gone = atomicAdd(&ticket[1], 1u);     // the last CTA to retire re-arms the ticket
      if (gone == gridDim.x - 1) { ticket[0] = 0u; ticket[1] = 0u; *d_epoch = epoch; }
    }
    return;
  }

  if (warp > IT_WARPS) {
    // ---------------------------------------------------------------- look-back warps: tiles k = slot, slot + 2, ...
    // (they never touch the stage ring: the compute warps hand them the tile number with agg_ready, and
    //  cannot run two sequence numbers ahead of a look-back warp, so the barrier phases cannot alias)
    // num_lb == 1: warp 17 alone takes every tile (experiment switch, RPB200_IL_NLB); else one warp per slot
    if (num_lb == 1 && warp != IT_WARPS + 1) return;
    for (int k = warp - IT_WARPS - 1;; k += num_lb) {
      const int slot = k % IT_SLOTS;
      mbar_wait(&S.agg_ready[slot], (k / IT_SLOTS) & 1);
      const unsigned int tile = S.lb_tile[slot];
      if (tile == IT_INVALID) break;
      const unsigned int wt = (lane < IT_WARPS) ? S.wtot[slot][lane] : 0u;
      unsigned int winc = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += up;
      }
      const unsigned int tile_total = __shfl_sync(0xffffffffu, winc, IT_WARPS - 1);

      unsigned int prefix = 0;
      if (tile != 0) {                  // the aggregate was already published by the last compute warp of A(k)
        // At full HBM speed ~56 tiles finish per microsecond across the GPU, more than one 32-descriptor round
        // trip can cover: every lane keeps IT_LBW independent descriptor loads in flight (128 tiles per round).
        long long look = (long long)tile - 1;
        for (;;) {
          unsigned long long w[IT_LBW];
#pragma unroll
          for (int j = 0; j < IT_LBW; ++j) {
            const long long idx = look - 32 * j - lane;
            w[j] = tag | (IT_INCLUSIVE << 32);                 // before tile 0: an inclusive prefix of 0
            if (idx >= 0) w[j] = it_ld(desc + idx * dstride);
          }
          bool found = false;
#pragma unroll
          for (int j = 0; j < IT_LBW; ++j) {
            if (!found) {                                       // warp-uniform
              const long long idx = look - 32 * j - lane;
              if (idx >= 0) {
                while ((w[j] >> 34) != (tag >> 34) || ((w[j] >> 32) & 3ull) == 0ull) {
                  if (backoff_ns) __nanosleep(backoff_ns);
                  w[j] = it_ld(desc + idx * dstride);
                }
              }
              const unsigned int incl = __ballot_sync(0xffffffffu, ((w[j] >> 32) & 3ull) == IT_INCLUSIVE);
              const int first = __ffs(incl) - 1;
              unsigned int c = (first < 0 || lane <= first) ? (unsigned int)w[j] : 0u;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
              prefix += c;
              found = first >= 0;
            }
          }
          if (found) break;
          look -= 32 * IT_LBW;
        }
        if (lane == 0) it_st(desc + (long long)tile * dstride, tag | (IT_INCLUSIVE << 32) | (unsigned long long)(prefix + tile_total));
      }
      if (lane < IT_WARPS) S.woff[slot][lane] = prefix + (winc - wt);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.prefix_ready[slot]);
    }
    return;
  }

  // ------------------------------------------------------------------------ compute warps 1..16
  const int cw = warp - 1;
  const int row = cw * 32 + lane;
  unsigned int mask = 0, nmask = 0, lane_excl = 0, n_lane_excl = 0, wtotal = 0, n_wtotal = 0;
  unsigned int tile = IT_INVALID, ntile = IT_INVALID;

  auto stage_A = [&](int k, unsigned int& m, unsigned int& lexcl, unsigned int& wsum, unsigned int& t) {
    const int st = k % IT_STAGES, slot = k % IT_SLOTS;
    mbar_wait(&S.full[st], (k / IT_STAGES) & 1);
    t = S.tile_id[st];
    if (t == IT_INVALID) {              // end of the stream: release the look-back warp of this sequence number
      if (lane == 0) { S.lb_tile[slot] = IT_INVALID; mbar_arrive(&S.agg_ready[slot]); }
      return;
    }
    m = 0;
    if ((long long)t * IT_ROWS + row < rows) {           // rows past the end: zero-filled or stale, never selected
      double r[IT_IPT];
      rpb_tma::load_row16(r, &S.tile[st][0], row);
#pragma unroll
      for (int i = 0; i < IT_IPT; ++i) m |= (r[i] < 0.0 ? 1u : 0u) << i;       // INDEXLIST_CONDITIONAL
    }
    const unsigned int cnt = __popc(m);
    unsigned int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    lexcl = inc - cnt;
    wsum = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 31) S.wtot[slot][cw] = inc;
    __syncwarp();
    unsigned int prev = 0;
    if (lane == 0) { __threadfence_block(); prev = atomicAdd(&S.arrived[slot], 1u); }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == IT_WARPS - 1) {         // last warp of the tile: publish the aggregate as early as possible
      __threadfence_block();
      unsigned int tot = lane < IT_WARPS ? S.wtot[slot][lane] : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane == 0) {
        S.arrived[slot] = 0u;
        it_st(desc + (long long)t * dstride, tag | ((t == 0 ? IT_INCLUSIVE : IT_PARTIAL) << 32) | (unsigned long long)tot);
      }
      __syncwarp();
    }
    if (lane == 0) { S.lb_tile[slot] = t; mbar_arrive(&S.empty[st]); mbar_arrive(&S.agg_ready[slot]); }
  };

  // Software pipeline of depth 2: the counts of tiles k+1 and k+2 are taken (and their look-backs started) before
  // this warp waits for the prefix of tile k, so a look-back has two tile times (~3.6 us) to resolve.
  int* __restrict__ slice = &S.slice[cw][0];
  unsigned int mask2 = 0, lane_excl2 = 0, wtotal2 = 0, tile2 = IT_INVALID;
  stage_A(0, mask, lane_excl, wtotal, tile);
  if (tile != IT_INVALID) stage_A(1, nmask, n_lane_excl, n_wtotal, ntile);
  int k = 0;
  for (; tile != IT_INVALID; ++k) {
    tile2 = IT_INVALID;
    if (ntile != IT_INVALID) stage_A(k + 2, mask2, lane_excl2, wtotal2, tile2);      // never past the terminating sequence number
    const int slot = k % IT_SLOTS;
    // compact this warp's indices into its slice while the prefix is being resolved
    {
      unsigned int at = lane_excl, m = mask;
      const int i0 = (int)(((long long)tile * IT_ROWS + row) * IT_IPT);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        slice[at++] = i0 + b;
      }
    }
    __syncwarp();
    mbar_wait(&S.prefix_ready[slot], (k / IT_SLOTS) & 1);
    int* __restrict__ out = list + S.woff[slot][cw];
    for (unsigned int q = lane; q < wtotal; q += 32) out[q] = slice[q];
    __syncwarp();                                                // the slice is rewritten by the next tile
    mask = nmask; lane_excl = n_lane_excl; wtotal = n_wtotal; tile = ntile;
    nmask = mask2; n_lane_excl = lane_excl2; n_wtotal = wtotal2; ntile = tile2;
  }
  // sequence number k was the terminating one (the look-back warp of its slot is released); release the others too
  if (lane == 0)
    for (int d = 1; d < IT_SLOTS; ++d) { S.lb_tile[(k + d) % IT_SLOTS] = IT_INVALID; mbar_arrive(&S.agg_ready[(k + d) % IT_SLOTS]); }
}

// the n % 16 elements after the last full row + the length (m_len): count so far = inclusive prefix of the last tile
__global__ void indexlist_tail_kernel(const double* __restrict__ x, int* __restrict__ list, long long first, int count,
                                      const unsigned long long* __restrict__ last_desc, long long* __restrict__ d_len)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long run = (long long)(unsigned int)(*last_desc);
  for (int i = 0; i < count; ++i)
    if (x[first + i] < 0.0) list[run++] = (int)(first + i);
  *d_len = run;
}

}  // namespace

// handled = 1 if the call was served here, 0 if the caller should use the register-staged kernel
int rpb_indexlist_tma_try(rpb200_ctx* ctx, const double* x, int* list, int64_t n, long long* d_len, unsigned long long* d_desc,
                          size_t desc_bytes, unsigned int* d_ticket, unsigned long long* d_epoch, cudaStream_t st, int* handled)
{
  *handled = 0;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("RPB200_INDEXLIST_NO_TMA"); disabled = (e && atoi(e)) ? 1 : 0; }
  if (disabled) return 0;
  const int64_t rows = n / IT_IPT;
  if (rows < (int64_t)IT_ROWS * ctx->sm_count * 2) return 0;
  if (!rpb_aligned(x, 16) || rows > 0x7fffffffll) return 0;
  const int64_t tiles = (rows + IT_ROWS - 1) / IT_ROWS;
  static int lbw = -1, dstride = -1, backoff = 0, nlb = IT_SLOTS;
  if (lbw < 0) {
    const char* e = getenv("RPB200_IL_LBW"); lbw = e ? atoi(e) : 1;
    e = getenv("RPB200_IL_DSTRIDE"); dstride = e ? atoi(e) : 16;
    e = getenv("RPB200_IL_BACKOFF"); backoff = e ? atoi(e) : 0;
    e = getenv("RPB200_IL_NLB"); nlb = (e && atoi(e) == 1) ? 1 : IT_SLOTS;
    if (lbw != 1 && lbw != 2 && lbw != 4) lbw = 1;
    if (dstride < 1) dstride = 1;
  }
  int ds = dstride;
  while (ds > 1 && sizeof(unsigned long long) * (size_t)tiles * ds > desc_bytes) ds >>= 1;     // denser if the state is small
  if (sizeof(unsigned long long) * (size_t)tiles * ds > desc_bytes) return 0;
  CUtensorMap map;
  if (!rpb_tma::make_row_map(&map, x, rows, IT_BOX_ROWS)) return 0;

  const size_t smem = sizeof(il_smem) + 1024;
  // per call: the attribute belongs to the current device's context, and a process may hold several contexts
  if (lbw == 4) RPB_CHECK(cudaFuncSetAttribute(indexlist_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (lbw == 2) RPB_CHECK(cudaFuncSetAttribute(indexlist_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else RPB_CHECK(cudaFuncSetAttribute(indexlist_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ctx->sm_count;
  if (grid > tiles) grid = (int)tiles;
  if (lbw == 4) indexlist_tma_kernel<4><<<grid, IT_THREADS, smem, st>>>(map, list, (long long)rows, d_desc, d_ticket, d_epoch, (unsigned int)tiles, ds, (unsigned int)backoff, nlb, rpb_tma::tma_evict_first());
  else if (lbw == 2) indexlist_tma_kernel<2><<<grid, IT_THREADS, smem, st>>>(map, list, (long long)rows, d_desc, d_ticket, d_epoch, (unsigned int)tiles, ds, (unsigned int)backoff, nlb, rpb_tma::tma_evict_first());
  else indexlist_tma_kernel<1><<<grid, IT_THREADS, smem, st>>>(map, list, (long long)rows, d_desc, d_ticket, d_epoch, (unsigned int)tiles, ds, (unsigned int)backoff, nlb, rpb_tma::tma_evict_first());
  RPB_LAUNCH_CHECK();
  indexlist_tail_kernel<<<1, 32, 0, st>>>(x, list, (long long)rows * IT_IPT, (int)(n - rows * IT_IPT), d_desc + (tiles - 1) * ds, d_len);
  RPB_LAUNCH_CHECK();
  *handled = 1;
  return 0;
}
