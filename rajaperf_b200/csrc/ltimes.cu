// ltimes.cu -- Apps_LTIMES: phi[z][g][m] += sum_d ell[m][d] * psi[z][g][d]   (FP64)
//
// Replaces apps/LTIMES-Cuda.cpp:44-102 (one thread per (z,g,m), a serial 64-long d loop doing a
// read-modify-write of phi in global memory per step, 25 of 32 x-lanes active).
//
// (z,g) flattens to one row index r, so this is the skinny GEMM  Phi[R x 25] += Psi[R x 64] * Ell^T.
// At 912 B and 3200 flop per row it sits at ~70% of the FP64 ridge: it is HBM-bound only if the
// FP64 issue slots are spent on nothing but math.  So for the suite's shape (num_d=64, num_m=25):
//   * the contraction runs on the FP64 tensor path, mma.sync.m8n8k4 (DMMA): one warp owns 8 rows;
//     columns 0..23 are three 8-wide DMMA tiles, column 24 is a 16-FMA dot per lane + 2 shuffles --
//     exactly 1600 FMA-equivalents per row, nothing padded;
//   * the A fragments (psi) are loaded STRAIGHT from global memory in fragment order with 256-bit
//     loads: lane (i,kk) owns the 32-byte piece kk of each of the four 128-byte lines of row r0+i
//     (d = 16j + 4kk + e), so one warp load instruction covers 8 whole lines; k-step s = 4j + e pairs
//     element s of every lane's pieces, which is a permutation of d -- legal because the sum over d
//     is order-free -- so no shared-memory staging or transposition of psi exists at all (owning the
//     contiguous 128 bytes 16kk .. 16kk+15 instead spreads every load over 32 lines: 3 % slower);
//   * the matching B fragments (ell) are loop-invariant and live in registers for the whole kernel;
//   * phi tiles (8 rows x 25 = 1600 contiguous bytes) are updated with coalesced 256-bit
//     read-modify-writes through a per-warp shared staging slab; the next tile's psi is prefetched
//     into registers while the current one is in the tensor pipe.
// Any other (num_d, num_m) takes a plain FMA kernel.
#include "common.cuh"

namespace {

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int LT_D = 64, LT_M = 25, LT_ROWS = 8;
constexpr int LT_WARPS = 4;

// LINE = false: lane (i,kk) owns the 128 contiguous bytes psi[row][16kk .. 16kk+15]; each of its four 256-bit loads
//   touches a different 128-byte line than the other three lanes of the row (32 lines per warp instruction).
// LINE = true:  lane (i,kk) owns the 32-byte piece kk of each of the row's four 128-byte lines, d = 16j + 4kk + e; one warp
//   instruction then covers 8 whole lines (one per row) -- a quarter of the L1 line requests for the same bytes.
// Either way k-step s pairs element s of the four lanes of a row: one more permutation of d, applied to ell alike.
template <bool LINE>
__device__ __forceinline__ int lt_d_of(int kk, int s) { return LINE ? 16 * (s >> 2) + 4 * kk + (s & 3) : 16 * kk + s; }

template <bool LINE>
__device__ __forceinline__ void lt_load_a(double (&a)[16], const double* __restrict__ psi, int64_t row, int kk)
{
  const double* p = psi + row * LT_D + (LINE ? kk * 4 : kk * 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const dbl4 v = ldg256_stream(p + (LINE ? 16 : 4) * j);
    a[4 * j] = v.x; a[4 * j + 1] = v.y; a[4 * j + 2] = v.z; a[4 * j + 3] = v.w;
  }
}

template <bool LINE>
__global__ void __launch_bounds__(LT_WARPS * 32, 2)
ltimes_dmma_kernel(double* __restrict__ phi, const double* __restrict__ ell,
                   const double* __restrict__ psi, int64_t ntiles)
{
  __shared__ __align__(32) double s_c[LT_WARPS][LT_ROWS * LT_M];   // 1600 B per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = lane >> 2, kk = lane & 3;
  double* sc = s_c[warp];

  // loop-invariant B fragments: b[t][s] = ell[m = 8t + i][d(kk, s)]; e24[s] = ell[24][d(kk, s)]
  double b[3][16], e24[16];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int s = 0; s < 16; ++s) b[t][s] = __ldg(ell + (8 * t + i) * LT_D + lt_d_of<LINE>(kk, s));
#pragma unroll
  for (int s = 0; s < 16; ++s) e24[s] = __ldg(ell + 24 * LT_D + lt_d_of<LINE>(kk, s));

  const int64_t wstride = (int64_t)gridDim.x * LT_WARPS;
  int64_t tile = (int64_t)blockIdx.x * LT_WARPS + warp;
  double a[16], an[16];
  if (tile < ntiles) lt_load_a<LINE>(a, psi, tile * LT_ROWS + i, kk);

  for (; tile < ntiles; tile += wstride) {
    const int64_t nxt = tile + wstride;
    if (nxt < ntiles) lt_load_a<LINE>(an, psi, nxt * LT_ROWS + i, kk);

    // old phi of this tile: 200 contiguous doubles = 50 vectors of 4
    double* ptile = phi + tile * (LT_ROWS * LT_M);
    const dbl4 old0 = ldg256(ptile + 4 * lane);
    dbl4 old1 = {0.0, 0.0, 0.0, 0.0};
    if (lane < 18) old1 = ldg256(ptile + 4 * (lane + 32));

    double c[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    double c24 = 0.0;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      dmma_884(c[0][0], c[0][1], a[s], b[0][s]);
      dmma_884(c[1][0], c[1][1], a[s], b[1][s]);
      dmma_884(c[2][0], c[2][1], a[s], b[2][s]);
      c24 = fma(a[s], e24[s], c24);
    }
    c24 += __shfl_xor_sync(0xffffffffu, c24, 1);
    c24 += __shfl_xor_sync(0xffffffffu, c24, 2);

    // C fragment: row i, columns 8t + 2kk + {0,1}
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      sc[i * LT_M + 8 * t + 2 * kk] = c[t][0];
      sc[i * LT_M + 8 * t + 2 * kk + 1] = c[t][1];
    }
    if (kk == 0) sc[i * LT_M + 24] = c24;
    __syncwarp();
    {
      const dbl4 add = *reinterpret_cast<const dbl4*>(sc + 4 * lane);
      dbl4 o; o.x = old0.x + add.x; o.y = old0.y + add.y; o.z = old0.z + add.z; o.w = old0.w + add.w;
      stg256(ptile + 4 * lane, o);
    }
    if (lane < 18) {
      const dbl4 add = *reinterpret_cast<const dbl4*>(sc + 4 * (lane + 32));
      dbl4 o; o.x = old1.x + add.x; o.y = old1.y + add.y; o.z = old1.z + add.z; o.w = old1.w + add.w;
      stg256(ptile + 4 * (lane + 32), o);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 16; ++s) a[s] = an[s];
  }
}


// ---- the same contraction with psi staged by bulk-async copies (opt-in tunings: measured SLOWER) ---------
// ncu on the kernel above: 8 warps per SM (242 registers), `long_scoreboard` the top stall, ~45 KB of loads in
// flight per SM -- so a staged variant was built to test whether memory-level parallelism was the limit.  It is not:
// 3, 4 or 6 stages all land at 5190-5230 GB/s against 6070 for the register prefetch (the wait -> LDS -> DMMA chain
// per 4 KB tile costs more than the prefetch registers).  Kept selectable (`unroll` 5..8) and covered by the tests.
// Every warp owns a 3-stage ring of
// psi tiles in shared memory (8 rows x 512 B, 576-byte pitch) filled two tiles ahead by cp.async.bulk (one row per
// lane 0..7, completion on a per-warp mbarrier): no prefetch registers, 64+ KB in flight per SM.  A lane reads its
// 128 bytes as eight 16-byte chunks starting at chunk kk (conflict-free with the 64-byte row padding); that is one
// more permutation of d, applied identically to the resident ell fragments.  phi of the NEXT tile is prefetched
// into the registers the psi prefetch no longer needs.
constexpr int LR_PITCH = 72;                    // doubles per staged row: 64 + 8 pad = 576 bytes (ONE_COPY: 64, no pad)

template <int LR_STAGES, bool ONE_COPY>
struct lt_ring_smem {
  alignas(128) double psi[LT_WARPS][LR_STAGES][LT_ROWS * (ONE_COPY ? LT_D : LR_PITCH)];
  alignas(32) double c[LT_WARPS][LT_ROWS * LT_M];
  unsigned long long full[LT_WARPS][LR_STAGES];
};

template <int LR_STAGES, bool ONE_COPY>
__global__ void __launch_bounds__(LT_WARPS * 32, 2)
ltimes_dmma_ring_kernel(double* __restrict__ phi, const double* __restrict__ ell,
                        const double* __restrict__ psi, int64_t ntiles)
{
  extern __shared__ __align__(128) unsigned char lt_smem_raw[];
  using smem_t = lt_ring_smem<LR_STAGES, ONE_COPY>;
  constexpr int PITCH = ONE_COPY ? LT_D : LR_PITCH;
  smem_t& S = *reinterpret_cast<smem_t*>(lt_smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = lane >> 2, kk = lane & 3;
  double* sc = S.c[warp];
  unsigned long long* full = S.full[warp];
  if (lane == 0) {
    for (int s = 0; s < LR_STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  // element e of this lane's chunk list is psi[row][16 kk + perm(e)], perm(2c + h) = 2 ((c + kk) & 7) + h
  double b[3][16], e24[16];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int d = 16 * kk + 2 * ((c + kk) & 7) + h;
#pragma unroll
      for (int t = 0; t < 3; ++t) b[t][2 * c + h] = __ldg(ell + (8 * t + i) * LT_D + d);
      e24[2 * c + h] = __ldg(ell + 24 * LT_D + d);
    }

  const int64_t wstride = (int64_t)gridDim.x * LT_WARPS;
  const int64_t tile0 = (int64_t)blockIdx.x * LT_WARPS + warp;
  auto issue = [&](int64_t tile, int stage) {         // lanes 0..7: one 512-byte row each
    if (lane == 0) mbar_arrive_expect_tx(&full[stage], LT_ROWS * LT_D * 8u);
    __syncwarp();
    if (ONE_COPY) {
      if (lane == 0) bulk_g2s(&S.psi[warp][stage][0], psi + tile * (LT_ROWS * LT_D), LT_ROWS * LT_D * 8u, &full[stage]);
    } else if (lane < LT_ROWS) {
      bulk_g2s(&S.psi[warp][stage][lane * PITCH], psi + (tile * LT_ROWS + lane) * LT_D, LT_D * 8u, &full[stage]);
    }
  };
#pragma unroll
  for (int s = 0; s < LR_STAGES - 1; ++s)
    if (tile0 + s * wstride < ntiles) issue(tile0 + s * wstride, s);

  dbl4 old0 = {0.0, 0.0, 0.0, 0.0}, old1 = {0.0, 0.0, 0.0, 0.0};
  if (tile0 < ntiles) {
    const double* ptile = phi + tile0 * (LT_ROWS * LT_M);
    old0 = ldg256(ptile + 4 * lane);
    if (lane < 18) old1 = ldg256(ptile + 4 * (lane + 32));
  }
  int it = 0;
  for (int64_t tile = tile0; tile < ntiles; tile += wstride, ++it) {
    const int stage = it % LR_STAGES;
    // the stage read in the previous iteration is free (every lane passed the __syncwarp that ended it)
    const int64_t ahead = tile + (int64_t)(LR_STAGES - 1) * wstride;
    if (ahead < ntiles) issue(ahead, (it + LR_STAGES - 1) % LR_STAGES);
    // phi of the next tile: lands while this one is in the tensor pipe
    const int64_t nxt = tile + wstride;
    dbl4 nold0 = {0.0, 0.0, 0.0, 0.0}, nold1 = {0.0, 0.0, 0.0, 0.0};
    if (nxt < ntiles) {
      const double* pn = phi + nxt * (LT_ROWS * LT_M);
      nold0 = ldg256(pn + 4 * lane);
      if (lane < 18) nold1 = ldg256(pn + 4 * (lane + 32));
    }

    mbar_wait(&full[stage], (it / LR_STAGES) & 1);
    double a[16];
    {
      const double* rowp = &S.psi[warp][stage][i * PITCH + 16 * kk];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double2 q = *reinterpret_cast<const double2*>(rowp + 2 * ((c + kk) & 7));
        a[2 * c] = q.x; a[2 * c + 1] = q.y;
      }
    }

    double c[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    double c24 = 0.0;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      dmma_884(c[0][0], c[0][1], a[s], b[0][s]);
      dmma_884(c[1][0], c[1][1], a[s], b[1][s]);
      dmma_884(c[2][0], c[2][1], a[s], b[2][s]);
      c24 = fma(a[s], e24[s], c24);
    }
    c24 += __shfl_xor_sync(0xffffffffu, c24, 1);
    c24 += __shfl_xor_sync(0xffffffffu, c24, 2);

    double* ptile = phi + tile * (LT_ROWS * LT_M);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      sc[i * LT_M + 8 * t + 2 * kk] = c[t][0];
      sc[i * LT_M + 8 * t + 2 * kk + 1] = c[t][1];
    }
    if (kk == 0) sc[i * LT_M + 24] = c24;
    __syncwarp();
    {
      const dbl4 add = *reinterpret_cast<const dbl4*>(sc + 4 * lane);
      dbl4 o; o.x = old0.x + add.x; o.y = old0.y + add.y; o.z = old0.z + add.z; o.w = old0.w + add.w;
      stg256(ptile + 4 * lane, o);
    }
    if (lane < 18) {
      const dbl4 add = *reinterpret_cast<const dbl4*>(sc + 4 * (lane + 32));
      dbl4 o; o.x = old1.x + add.x; o.y = old1.y + add.y; o.z = old1.z + add.z; o.w = old1.w + add.w;
      stg256(ptile + 4 * (lane + 32), o);
    }
    __syncwarp();
    old0 = nold0; old1 = nold1;
  }
}

// any shape: one thread per (row, m), FMA chain over d in the reference's order
__global__ void __launch_bounds__(256)
ltimes_generic_kernel(double* __restrict__ phi, const double* __restrict__ ell,
                      const double* __restrict__ psi, int64_t nd, int64_t nm, int64_t nrows)
{
  const int64_t total = nrows * nm;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / nm, m = idx - r * nm;
    const double* e = ell + m * nd;
    const double* p = psi + r * nd;
    double acc = phi[idx];
    for (int64_t d = 0; d < nd; ++d) acc = fma(__ldg(e + d), __ldg(p + d), acc);
    phi[idx] = acc;
  }
}

}  // namespace

extern "C" int rpb200_ltimes(rpb200_ctx* ctx, double* phi, const double* ell, const double* psi,
                             int64_t num_d, int64_t num_g, int64_t num_m, int64_t num_z,
                             rpb200_stream_t s)
{
  if (!ctx || num_d <= 0 || num_g <= 0 || num_m <= 0 || num_z < 0) return RPB200_EINVAL;
  RPB_CHECK_DEVICE(ctx);
  if (num_z == 0) return 0;
  if (!phi || !ell || !psi) return RPB200_EINVAL;
  cudaStream_t st = rpb_stream(s);
  const int64_t rows = num_z * num_g;
  if (num_d == LT_D && num_m == LT_M && rows % LT_ROWS == 0 && rpb_aligned(phi, 32) && rpb_aligned(psi, 32)) {
    const int64_t ntiles = rows / LT_ROWS;
    const int cps = ctx->tune[RPB_K_LTIMES].ctas_per_sm > 0 ? ctx->tune[RPB_K_LTIMES].ctas_per_sm : 2;
    int64_t grid = (int64_t)ctx->sm_count * cps;
    const int64_t need = (ntiles + LT_WARPS - 1) / LT_WARPS;
    if (need < grid) grid = need;
    const int variant = ctx->tune[RPB_K_LTIMES].unroll;
    if (variant < 5 || variant > 8) {                   // default: the register-prefetch kernel with line-major A fragments
                                                        // (6125-6260 GB/s; row chunks, variant 10: 5905-6066 in the same run; the
                                                        // staged variants below measure 4790-5230 GB/s, profiles/r01_widened.md)
      if (variant == 10) ltimes_dmma_kernel<false><<<(int)grid, LT_WARPS * 32, 0, st>>>(phi, ell, psi, ntiles);
      else ltimes_dmma_kernel<true><<<(int)grid, LT_WARPS * 32, 0, st>>>(phi, ell, psi, ntiles);
    } else {
#define RPB_LT_RING(S, O)                                                                                                   \
  do {                                                                                                                      \
    RPB_CHECK(cudaFuncSetAttribute(ltimes_dmma_ring_kernel<S, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lt_ring_smem<S, O>))); \
    ltimes_dmma_ring_kernel<S, O><<<(int)grid, LT_WARPS * 32, sizeof(lt_ring_smem<S, O>), st>>>(phi, ell, psi, ntiles);      \
  } while (0)
      switch (variant) {
        case 5: RPB_LT_RING(3, true); break;
        case 6: RPB_LT_RING(4, true); break;
        case 7: RPB_LT_RING(6, true); break;
        default: RPB_LT_RING(3, false); break;       // 8
      }
#undef RPB_LT_RING
    }
  } else {
    const int64_t total = rows * num_m;
    int64_t grid = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (grid > cap) grid = cap;
    ltimes_generic_kernel<<<(int)grid, 256, 0, st>>>(phi, ell, psi, num_d, num_m, rows);
  }
  RPB_LAUNCH_CHECK();
  return 0;
}
