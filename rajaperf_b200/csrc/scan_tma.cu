// scan_tma.cu -- Algorithm_SCAN, large-n path: persistent, warp-specialised, TMA-staged single-pass scan.
//
// Why: in the register-staged kernel (scan.cu) a CTA serialises  load -> scan -> look-back -> store ;
// ncu put 42 % of the warp time on the barrier behind the look-back, and prefetching the next tile
// through the LSU does not help (the descriptor polls queue behind the prefetch in the in-order L1
// pipeline).  Here, per SM, ONE CTA of 17 warps:
//   * warp 0, lane 0  = producer: draws tile tickets and issues cp.async.bulk.tensor (TMA) loads of
//     whole 8192-element tiles into a 3-stage shared-memory ring -- no registers, no LSU queue, always
//     two tiles ahead of the compute warps;
//   * warp 0          = look-back warp: turns the 16 warp totals of a tile into the tile aggregate,
//     publishes it, resolves the tile's exclusive prefix by decoupled look-back and hands the per-warp
//     offsets back through shared memory;
//   * warps 1..16     = compute warps: read their 16 contiguous doubles per thread from the ring
//     (the tensor map's 128-byte swizzle makes the 128-bit reads bank-conflict free), scan them in
//     registers, and -- software-pipelined -- scan tile k+1 BEFORE waiting for the prefix of tile k, so
//     the look-back latency of one tile is covered by the work on the next; results go out with
//     256-bit stores straight from registers.
// All hand-offs are mbarriers (no block-wide barrier in the steady state).  Traffic stays algorithmic:
// 8 B read + 8 B written per element.
#include "common.cuh"
#include "tma_stream.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int ST_WARPS = 16;                       // compute warps
constexpr int ST_THREADS = (ST_WARPS + 1) * 32;    // + warp 0 (producer / look-back)
constexpr int ST_IPT = 16;                         // doubles per compute thread = one 128-byte row
constexpr int ST_ROWS = ST_WARPS * 32;             // 512 rows per tile
constexpr int ST_TILE = ST_ROWS * ST_IPT;          // 8192 elements = 64 KiB
constexpr int ST_STAGES = 3;
constexpr int ST_BOX_ROWS = 256;                   // TMA box: 16 x 256 doubles = 32 KiB
constexpr unsigned int ST_INVALID = 0xffffffffu;
constexpr unsigned long long ST_PARTIAL = 1ull, ST_INCLUSIVE = 2ull;

struct __align__(16) tile_desc { unsigned long long word; double value; };

__device__ __forceinline__ void desc_store(tile_desc* p, unsigned long long word, double v)
{
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(word), "l"(__double_as_longlong(v)) : "memory");
}
__device__ __forceinline__ void desc_load(const tile_desc* p, unsigned long long& word, double& v)
{
  long long bits;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(word), "=l"(bits) : "l"(p) : "memory");
  v = __longlong_as_double(bits);
}

struct scan_smem {
  alignas(1024) double tile[ST_STAGES][ST_TILE];          // 3 x 64 KiB, 128-byte-swizzled rows
  unsigned long long full[ST_STAGES];                     // TMA landed (tx bytes)
  unsigned long long agg_ready[2];                        // 16 compute warps -> look-back warp
  unsigned long long prefix_ready[2];                     // look-back warp -> compute warps
  double wtot[2][ST_WARPS];                               // warp totals of a tile
  double woff[2][ST_WARPS];                               // exclusive prefix of every warp's segment
  unsigned int tile_id[ST_STAGES];
  unsigned int arrived[2];                                // compute warps done with A(k): the last one publishes
  unsigned long long epoch;                               // this call's epoch (*d_epoch + 1), read once per CTA
};

// (A line-major variant of the result stores -- a 4 x 4 piece transpose inside every group of four lanes, then whole-line
// stores -- was written in round 1 and first run in round 2: 6074 GB/s against 6210 for the row-per-thread stores below, and
// its transpose was wrong (profiles/r02_a_optin.log).  Deleted: the 32 extra shuffles per tile cost more than the store
// pattern gains once the loads come through TMA.)
template <int LB, int BACKOFF>
__global__ void __launch_bounds__(ST_THREADS, 1)
scan_tma_kernel(const __grid_constant__ CUtensorMap x_map, double* __restrict__ y, long long rows,
                tile_desc* __restrict__ desc, unsigned int* __restrict__ ticket, unsigned long long* d_epoch,
                unsigned int num_tiles, unsigned long long* __restrict__ dbg, int dstride)
{
  // epoch of this call = 1 + the last COMPLETED call's, in device memory (committed below by the last CTA to retire), so a
  // captured launch draws a fresh epoch at every replay
  // (read by thread 0 before the first barrier, broadcast through shared memory: every read happens-before the commit)
  // dstride: distance between tile descriptors in 16-byte units (2 = one per 32-byte sector: the polls
  // of a look-back round spread over more L2 lines/slices; 5860 -> 6146 GB/s at 2^27, profiles/r01_widened.md)
  // dbg (optional, RPB200_SCAN_DEBUG=1): per CTA {tiles, clk waiting for TMA, clk waiting for agg_ready,
  // clk in look-back, look-back rounds, descriptor polls, total clk}
  unsigned long long d_tiles = 0, d_full = 0, d_agg = 0, d_lb = 0, d_rounds = 0, d_polls = 0;
  const long long t_begin = clock64();
  extern __shared__ unsigned char smem_raw[];
  scan_smem& S = *reinterpret_cast<scan_smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_STAGES; ++s) mbar_init(&S.full[s], 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&S.agg_ready[s], ST_WARPS); mbar_init(&S.prefix_ready[s], 1); S.arrived[s] = 0u; }
    S.epoch = *(volatile unsigned long long*)d_epoch + 1ull;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long epoch = S.epoch;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer + look-back warp
    auto produce = [&](int seq) {        // lane 0 only: ticket -> TMA into stage seq % ST_STAGES
      const int st = seq % ST_STAGES;
      const unsigned int t = atomicAdd(&ticket[0], 1u);
      if (t < num_tiles) {
        S.tile_id[st] = t;
        const long long row0 = (long long)t * ST_ROWS;
        const bool second = row0 + ST_BOX_ROWS < rows;
        mbar_arrive_expect_tx(&S.full[st], (second ? 2u : 1u) * ST_BOX_ROWS * 128u);
        rpb_tma::tma_load_2d(&S.tile[st][0], &x_map, 0, (int)row0, &S.full[st]);
        if (second) rpb_tma::tma_load_2d(&S.tile[st][ST_BOX_ROWS * ST_IPT], &x_map, 0, (int)(row0 + ST_BOX_ROWS), &S.full[st]);
      } else {
        S.tile_id[st] = ST_INVALID;      // every CTA draws exactly one terminating ticket
        mbar_arrive(&S.full[st]);
      }
      return t < num_tiles;
    };
    bool more = true;
    if (lane == 0) {
      for (int s = 0; s < ST_STAGES && more; ++s) more = produce(s);
    }
    more = __shfl_sync(0xffffffffu, more ? 1 : 0, 0) != 0;

    for (int k = 0;; ++k) {
      const int st = k % ST_STAGES, slot = k & 1;
      // the stage used by sequence k-1 is free once every compute warp arrived on agg_ready(k-1),
      // which this warp waited for in the previous iteration
      if (k >= 1 && more) {
        if (lane == 0) more = produce(k - 1 + ST_STAGES);
        more = __shfl_sync(0xffffffffu, more ? 1 : 0, 0) != 0;
      }
      long long t0 = clock64();
      mbar_wait(&S.full[st], (k / ST_STAGES) & 1);
      const unsigned int tile = S.tile_id[st];
      if (tile == ST_INVALID) break;
      long long t1 = clock64();
      mbar_wait(&S.agg_ready[slot], (k >> 1) & 1);
      long long t2 = clock64();
      d_full += t1 - t0; d_agg += t2 - t1; d_tiles++;
      const double wt = (lane < ST_WARPS) ? S.wtot[slot][lane] : 0.0;
      double winc = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += up;
      }
      const double tile_total = __shfl_sync(0xffffffffu, winc, ST_WARPS - 1);

      // the tile aggregate was already published (PARTIAL; INCLUSIVE for tile 0) by the last compute warp
      // to finish A(k) -- it does not have to wait behind this warp's previous look-back
      double prefix = 0.0;
      if (tile != 0) {
        long long look = (long long)tile - 1;
        for (;;) {
          // lane l examines tiles look-l, look-32-l, ...: LB independent descriptor loads per lane per round
          unsigned long long word[LB];
          double val[LB];
#pragma unroll
          for (int j = 0; j < LB; ++j) {
            const long long idx = look - 32 * j - lane;
            word[j] = (epoch << 2) | ST_INCLUSIVE; val[j] = 0.0;
            if (idx >= 0) desc_load(desc + idx * dstride, word[j], val[j]);
          }
          bool found = false;
#pragma unroll
          for (int j = 0; j < LB; ++j) {
            if (!found) {                                   // warp-uniform
              const long long idx = look - 32 * j - lane;
              if (idx >= 0) {
                while ((word[j] >> 2) != epoch || (word[j] & 3ull) == 0ull) {
                  d_polls++;
                  if (BACKOFF > 0) __nanosleep(BACKOFF);
                  desc_load(desc + idx * dstride, word[j], val[j]);
                }
              }
              const unsigned int incl_mask = __ballot_sync(0xffffffffu, (word[j] & 3ull) == ST_INCLUSIVE);
              const int first = __ffs(incl_mask) - 1;
              const double contrib = (first < 0 || lane <= first) ? val[j] : 0.0;
              prefix += warp_sum(contrib);
              found = first >= 0;
            }
          }
          d_rounds++;
          if (found) break;
          look -= 32 * LB;
        }
        if (lane == 0) desc_store(desc + (long long)tile * dstride, (epoch << 2) | ST_INCLUSIVE, prefix + tile_total);
      }
      d_lb += clock64() - t2;
      if (lane < ST_WARPS) S.woff[slot][lane] = prefix + (winc - wt);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.prefix_ready[slot]);
    }
    // the last CTA to retire re-arms the ticket for the next call (no CTA can draw after that)
    if (dbg) {
      d_polls = __reduce_max_sync(0xffffffffu, (unsigned int)d_polls);
      if (lane == 0) {
        unsigned long long* o = dbg + 8 * blockIdx.x;
        o[0] = d_tiles; o[1] = d_full; o[2] = d_agg; o[3] = d_lb; o[4] = d_rounds; o[5] = d_polls; o[6] = clock64() - t_begin;
      }
    }
    if (lane == 0) {
      const unsigned int gone = atomicAdd(&ticket[1], 1u);
      if (gone == gridDim.x - 1) { ticket[0] = 0u; ticket[1] = 0u; *d_epoch = epoch; }
    }
    return;
  }

  // ------------------------------------------------------------------------ compute warps 1..16
  const int cw = warp - 1;                      // 0..15
  const int row = cw * 32 + lane;               // this thread's row of the tile
  double v[ST_IPT], nv[ST_IPT];
  double lane_excl = 0.0, n_lane_excl = 0.0;
  unsigned int tile = ST_INVALID, ntile = ST_INVALID;

  // A(k): wait for the stage, scan the row in registers, publish the warp total
  auto stage_A = [&](int k, double (&r)[ST_IPT], double& lexcl, unsigned int& t) {
    const int st = k % ST_STAGES, slot = k & 1;
    mbar_wait(&S.full[st], (k / ST_STAGES) & 1);
    t = S.tile_id[st];
    if (t == ST_INVALID) return;
    rpb_tma::load_row16(r, &S.tile[st][0], row);
    if ((long long)t * ST_ROWS + row >= rows) {          // rows past the end of the array (zero-filled or stale)
#pragma unroll
      for (int i = 0; i < ST_IPT; ++i) r[i] = 0.0;
    }
    double run = 0.0;
#pragma unroll
    for (int i = 0; i < ST_IPT; ++i) { const double e = r[i]; r[i] = run; run += e; }
    double inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    lexcl = inc - run;
    if (lane == 31) S.wtot[slot][cw] = inc;
    __syncwarp();
    unsigned int prev = 0;
    if (lane == 0) { __threadfence_block(); prev = atomicAdd(&S.arrived[slot], 1u); }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == ST_WARPS - 1) {
      // last warp of the tile: publish the aggregate NOW, so successors' look-backs see it as early as possible
      __threadfence_block();
      const double tot = warp_sum(lane < ST_WARPS ? S.wtot[slot][lane] : 0.0);
      if (lane == 0) {
        S.arrived[slot] = 0u;
        desc_store(desc + (long long)t * dstride, (epoch << 2) | (t == 0 ? ST_INCLUSIVE : ST_PARTIAL), tot);
      }
      __syncwarp();
    }
    if (lane == 0) mbar_arrive(&S.agg_ready[slot]);     // also: this warp is done reading the stage
  };

  stage_A(0, v, lane_excl, tile);
  for (int k = 0; tile != ST_INVALID; ++k) {
    stage_A(k + 1, nv, n_lane_excl, ntile);              // overlaps the look-back of tile k
    const int slot = k & 1;
    mbar_wait(&S.prefix_ready[slot], (k >> 1) & 1);
    const double off = S.woff[slot][cw] + lane_excl;
    const long long grow = (long long)tile * ST_ROWS + row;
    if (grow < rows) {
      double* yp = y + grow * ST_IPT;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        dbl4 o;
        o.x = off + v[4 * q]; o.y = off + v[4 * q + 1]; o.z = off + v[4 * q + 2]; o.w = off + v[4 * q + 3];
        stg256_stream(yp + 4 * q, o);
      }
    }
#pragma unroll
    for (int i = 0; i < ST_IPT; ++i) v[i] = nv[i];
    lane_excl = n_lane_excl;
    tile = ntile;
  }
}

// the n % 16 elements after the last full row: prefix = inclusive prefix of the last tile
__global__ void scan_tail_kernel(const double* __restrict__ x, double* __restrict__ y, long long first, int count,
                                 const tile_desc* __restrict__ last_desc)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double run = last_desc->value;
  for (int i = 0; i < count; ++i) { y[first + i] = run; run += x[first + i]; }
}

}  // namespace

// Returns 1 if the call was handled here, 0 if the caller should use the register-staged kernel, < 0 / cudaError on failure.
int rpb_scan_tma_try(rpb200_ctx* ctx, const double* x, double* y, int64_t n, void* d_desc, size_t desc_bytes,
                     unsigned int* d_ticket, unsigned long long* d_epoch, cudaStream_t st, int* handled)
{
  *handled = 0;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("RPB200_SCAN_NO_TMA"); disabled = (e && atoi(e)) ? 1 : 0; }
  if (disabled) return 0;
  const int64_t rows = n / ST_IPT;
  if (rows < (int64_t)ST_ROWS * ctx->sm_count * 2) return 0;          // small problems: the simpler kernel
  if (!rpb_aligned(x, 16) || !rpb_aligned(y, 32) || rows > 0x7fffffffll) return 0;
  rpb_tma::encode_fn_t encode = rpb_tma::get_encode();
  if (!encode) return 0;
  const int64_t tiles = (rows + ST_ROWS - 1) / ST_ROWS;
  static int dstride = -1;
  if (dstride < 0) { const char* e = getenv("RPB200_SCAN_DSTRIDE"); dstride = e ? atoi(e) : 2; if (dstride < 1) dstride = 1; }
  int ds = dstride;
  while (ds > 1 && sizeof(tile_desc) * (size_t)tiles * ds > desc_bytes) ds >>= 1;      // denser if the state is small
  if (sizeof(tile_desc) * (size_t)tiles * ds > desc_bytes) return 0;

  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)ST_IPT, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)(ST_IPT * sizeof(double))};
  const cuuint32_t box[2] = {(cuuint32_t)ST_IPT, (cuuint32_t)ST_BOX_ROWS};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, rpb_tma::l2_promotion(),
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;

  const size_t smem = sizeof(scan_smem) + 1024;
  int grid = ctx->sm_count;
  if (grid > tiles) grid = (int)tiles;
  static unsigned long long* dbg = nullptr;
  static int want_dbg = -1;
  if (want_dbg < 0) {
    const char* e = getenv("RPB200_SCAN_DEBUG");
    want_dbg = (e && atoi(e)) ? 1 : 0;
    if (want_dbg) { RPB_CHECK(cudaMalloc(&dbg, 8 * sizeof(unsigned long long) * 1024)); RPB_CHECK(cudaMemset(dbg, 0, 8 * 8 * 1024)); }
  }
  RPB_CHECK(cudaFuncSetAttribute(scan_tma_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scan_tma_kernel<1, 0><<<grid, ST_THREADS, smem, st>>>(map, y, (long long)rows, (tile_desc*)d_desc, d_ticket, d_epoch,
                                                        (unsigned int)tiles, dbg, ds);
  RPB_LAUNCH_CHECK();
  if (want_dbg) {
    static int printed = 0;
    if (printed++ == 3) {       // the 4th call: warm
      unsigned long long h[8 * 160];
      RPB_CHECK(cudaMemcpyAsync(h, dbg, sizeof(unsigned long long) * 8 * grid, cudaMemcpyDeviceToHost, st));
      RPB_CHECK(cudaStreamSynchronize(st));
      double a[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < grid; ++c) for (int i = 0; i < 7; ++i) a[i] += (double)h[8 * c + i];
      fprintf(stderr, "[scan_tma dbg] per tile: wait_tma %.0f clk, wait_agg %.0f clk, lookback %.0f clk, rounds %.2f, polls(max lane) %.1f; per CTA total %.0f clk, tiles %.1f\n",
              a[1] / a[0], a[2] / a[0], a[3] / a[0], a[4] / a[0], a[5] / a[0], a[6] / grid, a[0] / grid);
    }
  }
  const int rem = (int)(n - rows * ST_IPT);
  if (rem > 0) {
    scan_tail_kernel<<<1, 32, 0, st>>>(x, y, (long long)rows * ST_IPT, rem, (const tile_desc*)d_desc + (tiles - 1) * ds);
    RPB_LAUNCH_CHECK();
  }
  *handled = 1;
  return 0;
}
