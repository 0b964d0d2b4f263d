// sort.cu -- Algorithm_SORT / Algorithm_SORTPAIRS: LSD radix sort of doubles for sm_100a.
//
// Replaces algorithm/SORT-Cuda.cpp:35-43 and SORTPAIRS-Cuda.cpp:35-43, i.e. RAJA::sort /
// RAJA::sort_pairs -> cub::DeviceRadixSort (tpl/RAJA/include/RAJA/policy/cuda/sort.hpp:86-146,
// 337-409) with its per-call pool malloc/free and tail cudaMemcpyAsync.  Written from scratch:
//   * keys are mapped to order-preserving uint64 (sign flip / full flip) on the way in and back on
//     the way out, fused into the first and last pass;
//   * one histogram kernel reads the keys once and builds all 8 digit histograms;
//   * 8 "onesweep" passes (8-bit digits): each tile ranks its keys with warp match-any, publishes
//     its 256 digit counts, and resolves the counts of all earlier tiles by decoupled look-back, so
//     every pass reads and writes each key exactly once (16 B/key/pass, 32 B for pairs);
//   * keys are first reordered by digit in shared memory, so the global scatter is made of
//     contiguous runs;
//   * UNIFORM TILES: a tile whose 8192 keys all carry the same digit in this pass (the suite's keys are rand()/RAND_MAX in
//     [0, 1]: their top byte is 0x3f for 99.997 % of them, so ~78 % of the tiles of the last pass; any pass of clustered
//     or already sorted data) needs no ranking at all -- its keys keep their order and move as one block to
//     base[digit] + prefix.  A warp vote per round and one shared-memory word per warp detect it; the tile then skips the
//     8-ballot match (45 % of a pass's instructions), the warp-count scan and the shared-memory reorder, and streams
//     its keys out of the registers they were loaded into;
//   * look-back descriptors carry a pass-parity code, so they are never cleared between passes or
//     between calls of the same size; tiles are dealt by atomic ticket (no reliance on CTA order);
//   * caller-provided scratch, no allocation, no host synchronisation; stable (pairs well-defined).
#include "common.cuh"

#include <type_traits>

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int NUM_PASSES = 64 / RADIX_BITS;

constexpr int SORT_WARPS = 16;                 // 512 threads
constexpr int SORT_IPT = 16;                   // keys per thread
constexpr int SORT_BLOCK = SORT_WARPS * 32;
constexpr int SORT_TILE = SORT_BLOCK * SORT_IPT;   // 8192 keys

// double -> uint64 whose unsigned order is the IEEE total order (-0 < +0, no NaN handling needed)
__device__ __forceinline__ unsigned long long key_encode(unsigned long long b)
{
  const unsigned long long m = (unsigned long long)((long long)b >> 63) | 0x8000000000000000ull;
  return b ^ m;
}
__device__ __forceinline__ unsigned long long key_decode(unsigned long long k)
{
  const unsigned long long m = (unsigned long long)((long long)(~k) >> 63) | 0x8000000000000000ull;
  return k ^ m;
}

// Lanes holding the same 8-bit digit.  Built from 8 warp ballots: the hardware MATCH.ANY instruction
// is an order of magnitude slower on sm_100 (profiles/r01_sort_ncu.md).  Four instructions per bit: test the bit into a
// predicate (one LOP3), ballot, turn the predicate into an all-ones / all-zeros word (SEL), and fold
// peers &= ~(vote ^ own) with ONE three-input LOP3 (0x90 = a & ~(b ^ c)).  The round-1 form -- (d >> bit) & 1, then
// `one ? vote : ~vote` -- compiled to six (SHF, LOP3, ISETP, VOTE, SEL, LOP3): the ranking loop is 45 % of a pass's
// instructions and the pass is issue-bound, so the two instructions per bit are ~10 % of the sort.
__device__ __forceinline__ unsigned int match_digit(unsigned int d)
{
  unsigned int peers = 0xffffffffu;
#pragma unroll
  for (int bit = 0; bit < RADIX_BITS; ++bit) {
    // written in PTX: nvcc rewrites the C form back into shift / and / compare / negate (6 SASS instructions per bit)
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 ".reg .b32 t, v;\n\t"
                 "and.b32 t, %1, %2;\n\t"
                 "setp.ne.u32 p, t, 0;\n\t"
                 "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
                 "selp.b32 t, 0xffffffff, 0, p;\n\t"
                 "lop3.b32 %0, %0, v, t, 0x90;\n\t"
                 "}" : "+r"(peers) : "r"(d), "r"(1u << bit));
  }
  return peers;
}

struct sort_scratch_layout {
  size_t off_alt_keys, off_alt_vals, off_hist, off_ctrs, off_desc, total;
  int64_t tiles;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline sort_scratch_layout make_layout(int64_t n, int pairs)
{
  sort_scratch_layout L;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  L.tiles = (int64_t)((nn + SORT_TILE - 1) / SORT_TILE);
  size_t o = 0;
  L.off_alt_keys = o; o = align_up(o + nn * 8, 256);
  L.off_alt_vals = o; if (pairs) o = align_up(o + nn * 8, 256);
  L.off_hist = o;     o = align_up(o + sizeof(unsigned long long) * NUM_PASSES * RADIX, 256);
  L.off_ctrs = o;     o = align_up(o + sizeof(unsigned int) * 2 * NUM_PASSES + 64, 256);
  L.off_desc = o;     o = align_up(o + sizeof(unsigned long long) * RADIX * (size_t)L.tiles, 256);
  L.total = o;
  return L;
}

// ----------------------------------------------------------------------------------------------
// histogram of all 8 digits in one read of the keys
// ----------------------------------------------------------------------------------------------
// one key into the 8 shared-memory histograms.  The top byte (sign + high exponent bits) of similar-magnitude keys
// usually hits one bin for the whole warp: it is counted with one atomic per warp instead of a 32-way conflict.
__device__ __forceinline__ void hist_one(unsigned int (*s_hist)[RADIX], unsigned long long k, bool ok, int lane)
{
  const unsigned int lo = (unsigned int)k, hi = (unsigned int)(k >> 32);
  if (ok) {
    atomicAdd(&s_hist[0][lo & 0xffu], 1u);
    atomicAdd(&s_hist[1][(lo >> 8) & 0xffu], 1u);
    atomicAdd(&s_hist[2][(lo >> 16) & 0xffu], 1u);
    atomicAdd(&s_hist[3][lo >> 24], 1u);
    atomicAdd(&s_hist[4][hi & 0xffu], 1u);
    atomicAdd(&s_hist[5][(hi >> 8) & 0xffu], 1u);
    atomicAdd(&s_hist[6][(hi >> 16) & 0xffu], 1u);
  }
  const unsigned int d7 = hi >> 24;                        // sign + high exponent bits
  const unsigned int d7_0 = __shfl_sync(0xffffffffu, d7, 0);
  if (__all_sync(0xffffffffu, ok && d7 == d7_0)) {
    if (lane == 0) atomicAdd(&s_hist[7][d7], 32u);
  } else if (ok) {
    atomicAdd(&s_hist[7][d7], 1u);
  }
}

// 256-bit loads, two independent vectors per thread per trip (the kernel is a pure stream of the keys)
__global__ void __launch_bounds__(512)
sort_hist_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                 unsigned long long* __restrict__ g_hist, int vector_ok)
{
  __shared__ unsigned int s_hist[NUM_PASSES][RADIX];
  for (int i = threadIdx.x; i < NUM_PASSES * RADIX; i += blockDim.x) (&s_hist[0][0])[i] = 0u;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nv = vector_ok ? n / 4 : 0;                // whole 4-key vectors
  const int64_t nv_round = (nv + 31) & ~31ll;              // whole warps iterate together
  for (int64_t v = gtid; v < nv_round; v += 2 * stride) {
    const int64_t v1 = v + stride;
    const bool ok0 = v < nv, ok1 = v1 < nv;
    dbl4 q0, q1;
    if (ok0) q0 = ldg256_stream(reinterpret_cast<const double*>(keys) + 4 * v);
    if (ok1) q1 = ldg256_stream(reinterpret_cast<const double*>(keys) + 4 * v1);
    const double e0[4] = {q0.x, q0.y, q0.z, q0.w}, e1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) hist_one(s_hist, ok0 ? key_encode((unsigned long long)__double_as_longlong(e0[j])) : 0ull, ok0, lane);
    if (v1 < nv_round) {                                    // warp-uniform
#pragma unroll
      for (int j = 0; j < 4; ++j) hist_one(s_hist, ok1 ? key_encode((unsigned long long)__double_as_longlong(e1[j])) : 0ull, ok1, lane);
    }
  }
  const int64_t r0 = 4 * nv, nr_round = (n - r0 + 31) & ~31ll;     // scalar remainder (everything if unaligned)
  for (int64_t i = gtid; i < nr_round; i += stride) {
    const bool ok = r0 + i < n;
    hist_one(s_hist, ok ? key_encode(keys[r0 + i]) : 0ull, ok, lane);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NUM_PASSES * RADIX; i += blockDim.x) {
    const unsigned int c = (&s_hist[0][0])[i];
    if (c) atomicAdd(&g_hist[i], (unsigned long long)c);
  }
}

// ---- the same histograms with LANE-PRIVATE counters (the default since round 2: bit-exact at 2^27 / 1000003 / 4097 keys and
// 7.818 ms against 7.863 ms for the whole keys sort at 2^27, profiles/r02_a_optin.log; tuning `unroll` 8 selects the
// shared-bin kernel above) ---------
// sort_hist_kernel is bound by its shared-memory atomics: 32 random bins per warp instruction fall on ~3.5 addresses of the
// busiest bank (437 us for 2^27 keys against 165 us of DRAM time).  Here lane l of every warp counts in its own column: bin b
// of digit d lives in the 16-bit half (b & 1) of word col[d][(b >> 1) * 32 + l], so the 32 atomics of a warp instruction hit
// 32 different banks whatever the digits are.  A column is shared by the 32 warps of the CTA, hence a 16-bit counter may
// receive 32 increments per "round" of one key per thread: the columns are folded into the 64-bit global histograms (and
// cleared) every 2040 keys per thread, and at the end.  The top byte keeps the warp-uniform shortcut of hist_one.
constexpr int HL_WARPS = 32, HL_BLOCK = HL_WARPS * 32, HL_DIGITS = NUM_PASSES - 1;
constexpr int HL_COL_WORDS = (RADIX / 2) * 32;                       // words per digit: 128 bin pairs x 32 lanes
constexpr int HL_FLUSH_KEYS = 2040;                                  // 32 warps x 2040 < 65536
constexpr size_t HL_SMEM = sizeof(unsigned int) * (HL_DIGITS * HL_COL_WORDS + RADIX);

__device__ __forceinline__ void hist_lane_one(unsigned int* col, unsigned int* s_top, unsigned long long k, bool ok, int lane)
{
  const unsigned int lo = (unsigned int)k, hi = (unsigned int)(k >> 32);
  if (ok) {
#pragma unroll
    for (int d = 0; d < HL_DIGITS; ++d) {
      const unsigned int b = d < 4 ? (lo >> (8 * d)) & 0xffu : (hi >> (8 * (d - 4))) & 0xffu;
      atomicAdd(&col[d * HL_COL_WORDS + ((b >> 1) << 5) + lane], 1u << ((b & 1u) << 4));
    }
  }
  const unsigned int d7 = hi >> 24;
  const unsigned int d7_0 = __shfl_sync(0xffffffffu, d7, 0);
  if (__all_sync(0xffffffffu, ok && d7 == d7_0)) {
    if (lane == 0) atomicAdd(&s_top[d7], 32u);
  } else if (ok) {
    atomicAdd(&s_top[d7], 1u);
  }
}

// fold the lane columns into the global histograms and clear them (whole CTA, between two barriers)
__device__ __forceinline__ void hist_lane_flush(unsigned int* col, unsigned long long* __restrict__ g_hist)
{
  __syncthreads();
  for (int e = threadIdx.x; e < HL_DIGITS * RADIX; e += HL_BLOCK) {
    const int d = e >> 8, b = e & 0xff;
    const unsigned int* row = col + d * HL_COL_WORDS + ((b >> 1) << 5);
    const int sh = (b & 1) << 4;
    unsigned int c = 0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) c += (row[(l + threadIdx.x) & 31] >> sh) & 0xffffu;     // rotated: no common bank
    if (c) atomicAdd(&g_hist[e], (unsigned long long)c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HL_DIGITS * HL_COL_WORDS; i += HL_BLOCK) col[i] = 0u;
  __syncthreads();
}

__global__ void __launch_bounds__(HL_BLOCK, 1)
sort_hist_lanes_kernel(const unsigned long long* __restrict__ keys, int64_t n, unsigned long long* __restrict__ g_hist,
                       int vector_ok)
{
  extern __shared__ __align__(16) unsigned int hl_smem[];
  unsigned int* col = hl_smem;                                        // [HL_DIGITS][128][32]
  unsigned int* s_top = hl_smem + HL_DIGITS * HL_COL_WORDS;           // [RADIX]
  for (int i = threadIdx.x; i < HL_DIGITS * HL_COL_WORDS + RADIX; i += HL_BLOCK) hl_smem[i] = 0u;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nv = vector_ok ? n / 4 : 0;                          // whole 4-key vectors
  // every thread of the CTA runs the same number of trips (the flush is a CTA barrier): round the trip space up per CTA
  const int64_t trips = (nv + 2 * stride - 1) / (2 * stride);
  int since_flush = 0;
  for (int64_t it = 0; it < trips; ++it) {
    const int64_t v0 = gtid + 2 * it * stride, v1 = v0 + stride;
    const bool ok0 = v0 < nv, ok1 = v1 < nv;
    dbl4 q0 = {0.0, 0.0, 0.0, 0.0}, q1 = {0.0, 0.0, 0.0, 0.0};
    if (ok0) q0 = ldg256_stream(reinterpret_cast<const double*>(keys) + 4 * v0);
    if (ok1) q1 = ldg256_stream(reinterpret_cast<const double*>(keys) + 4 * v1);
    const double e0[4] = {q0.x, q0.y, q0.z, q0.w}, e1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) hist_lane_one(col, s_top, key_encode((unsigned long long)__double_as_longlong(e0[j])), ok0, lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) hist_lane_one(col, s_top, key_encode((unsigned long long)__double_as_longlong(e1[j])), ok1, lane);
    since_flush += 8;
    if (since_flush + 8 > HL_FLUSH_KEYS) { hist_lane_flush(col, g_hist); since_flush = 0; }     // CTA-uniform
  }
  const int64_t r0 = 4 * nv;                                          // scalar remainder (everything if unaligned)
  const int64_t rtrips = (n - r0 + stride - 1) / stride;
  for (int64_t it = 0; it < rtrips; ++it) {
    const int64_t i = r0 + gtid + it * stride;
    const bool ok = i < n;
    hist_lane_one(col, s_top, ok ? key_encode(keys[i]) : 0ull, ok, lane);
    if (++since_flush + 8 > HL_FLUSH_KEYS) { hist_lane_flush(col, g_hist); since_flush = 0; }
  }
  hist_lane_flush(col, g_hist);
  for (int i = threadIdx.x; i < RADIX; i += HL_BLOCK) {
    const unsigned int c = s_top[i];
    if (c) atomicAdd(&g_hist[HL_DIGITS * RADIX + i], (unsigned long long)c);
  }
}

// exclusive scan of each digit histogram in place: g_hist[p][b] = #keys with digit_p < b; skewed[p] = 1 if one digit of pass p
// holds at least a quarter of the keys (then uniform tiles are likely and the pass tests for them)
__global__ void __launch_bounds__(RADIX)
sort_hist_scan_kernel(unsigned long long* __restrict__ g_hist, unsigned int* __restrict__ skewed, unsigned long long n)
{
  __shared__ unsigned long long s[RADIX];
  const int p = blockIdx.x, b = threadIdx.x;
  const unsigned long long c = g_hist[p * RADIX + b];
  if (4ull * c >= n) skewed[p] = 1u;                 // cleared per call with the histograms
  s[b] = c;
  __syncthreads();
  for (int o = 1; o < RADIX; o <<= 1) {
    const unsigned long long t = (b >= o) ? s[b - o] : 0ull;
    __syncthreads();
    s[b] += t;
    __syncthreads();
  }
  g_hist[p * RADIX + b] = s[b] - c;
}

// ----------------------------------------------------------------------------------------------
// one onesweep pass
// ----------------------------------------------------------------------------------------------
// descriptor word = (count << 2) | code; code = 2*parity + (0: tile aggregate, 1: inclusive prefix).
// A word whose parity differs from the running pass is a leftover of the previous pass: not ready.
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// digit of a pass: the shift is a multiple of 8, so the byte lives in one 32-bit half of the key
__device__ __forceinline__ unsigned int key_digit(unsigned long long k, int shift)
{
  const unsigned int w = (shift & 32) ? (unsigned int)(k >> 32) : (unsigned int)k;
  return (w >> (shift & 31)) & (RADIX - 1);
}

template <bool PAIRS, bool FIRST, bool LAST, int SORT_LB>     // SORT_LB: tile descriptors in flight per look-back step
__global__ void __launch_bounds__(SORT_BLOCK, 2)
sort_onesweep_kernel(const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
                     const unsigned long long* __restrict__ vals_in, unsigned long long* __restrict__ vals_out,
                     int64_t n, int shift, const unsigned long long* __restrict__ g_base /*[RADIX]*/,
                     unsigned long long* __restrict__ desc, unsigned int* __restrict__ ticket,
                     unsigned int num_tiles, unsigned int parity, const unsigned int* __restrict__ skewed)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw);            // [TILE]
  unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_keys + SORT_TILE);               // [WARPS][RADIX]
  unsigned int* s_bin_start = s_cnt + SORT_WARPS * RADIX;                                  // [RADIX]
  long long* s_delta = reinterpret_cast<long long*>(s_bin_start + RADIX);                  // [RADIX]
  unsigned char* s_digit = reinterpret_cast<unsigned char*>(s_delta + RADIX);              // [TILE] (pairs)
  __shared__ unsigned int s_tile;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int lt_mask = (1u << lane) - 1u;

  if (threadIdx.x == 0) s_tile = atomicAdd(&ticket[0], 1u);
  for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_BLOCK) s_cnt[i] = 0u;
  __syncthreads();
  const unsigned int tile = s_tile;
  const int64_t tile_base = (int64_t)tile * SORT_TILE;
  const int valid = (int)((n - tile_base) < (int64_t)SORT_TILE ? (n - tile_base) : (int64_t)SORT_TILE);

  // ---- load: warp w owns the contiguous chunk [w*32*IPT, (w+1)*32*IPT); round r, lane l
  unsigned long long key[SORT_IPT];
  const int chunk = warp * 32 * SORT_IPT + lane;
  const bool full = (valid == SORT_TILE);                     // every tile but the last: no bounds tests
  if (full) {
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
      unsigned long long k = keys_in[tile_base + chunk + r * 32];
      key[r] = FIRST ? key_encode(k) : k;
    }
  } else {
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
      const int p = chunk + r * 32;
      unsigned long long k = 0xffffffffffffffffull;           // padding sorts last, never written
      if (p < valid) {
        k = keys_in[tile_base + p];
        if (FIRST) k = key_encode(k);
      }
      key[r] = k;
    }
  }

  // ---- uniform tile?  every key of the tile has the digit of the tile's first key (full tiles only: the padding of the
  // last tile carries digit 0xff).  Tested only in passes whose histogram is skewed enough for such tiles to be likely
  // (`skewed`: some digit holds at least a quarter of the keys, sort_hist_scan_kernel), and cheaply: OR the XOR of every key
  // with the thread's first key (one LOP3 per key on the 32-bit half that holds the digit), look at the digit's bits once,
  // one vote per warp, one word per warp in shared memory, one barrier.
  __shared__ unsigned int s_wdigit[SORT_WARPS];
  bool uniform = false;                                    // CTA-uniform
  if (skewed[0] != 0u) {
    const unsigned int d0 = __shfl_sync(0xffffffffu, key_digit(key[0], shift), 0);
    const bool hi = (shift & 32) != 0;
    const unsigned int w0 = hi ? (unsigned int)(key[0] >> 32) : (unsigned int)key[0];
    unsigned int diff = 0;
#pragma unroll
    for (int r = 1; r < SORT_IPT; ++r) diff |= (hi ? (unsigned int)(key[r] >> 32) : (unsigned int)key[r]) ^ w0;
    bool same = full && ((diff >> (shift & 31)) & (RADIX - 1)) == 0u && key_digit(key[0], shift) == d0;
    same = __all_sync(0xffffffffu, same);
    if (lane == 0) s_wdigit[warp] = same ? d0 : 0xffffffffu;
    __syncthreads();
    unsigned int udig = s_wdigit[0];
#pragma unroll
    for (int w = 1; w < SORT_WARPS; ++w) udig = (s_wdigit[w] == udig) ? udig : 0xffffffffu;
    uniform = udig != 0xffffffffu;
  }
  const unsigned int udigit = uniform ? s_wdigit[0] : 0u;

  // ---- rank inside the warp chunk (stable): match-any on the digit.  The pairs kernel keeps two 16-bit ranks (and, below,
  // positions) per register: its value loads need the registers (no spills: 80-88 bytes before); the keys kernel is bound
  // by instruction issue and keeps one per register (the packing costs ~4 instructions per key: 3 % of a pass, measured).
  constexpr int RK = PAIRS ? SORT_IPT / 2 : SORT_IPT;
  unsigned int rank2[RK];
  auto rank_get = [&](int r) { return PAIRS ? ((rank2[r >> 1] >> ((r & 1) * 16)) & 0xffffu) : rank2[r]; };
  unsigned int* my_cnt = s_cnt + warp * RADIX;
  if (!uniform) {
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
      const unsigned int d = key_digit(key[r], shift);
      const unsigned int peers = match_digit(d);
      const int leader = __ffs(peers) - 1;
      unsigned int before = 0;
      if (lane == leader) { before = my_cnt[d]; my_cnt[d] = before + __popc(peers); }
      before = __shfl_sync(0xffffffffu, before, leader);
      const unsigned int rk = before + __popc(peers & lt_mask);
      if (PAIRS) rank2[r >> 1] = (r & 1) ? (rank2[r >> 1] | (rk << 16)) : rk;
      else rank2[r] = rk;
      __syncwarp();
    }
    __syncthreads();
  }

  // ---- per digit: scan the warp counts, publish the tile count, look back
  unsigned int my_total = 0;
  if (threadIdx.x < RADIX) {
    const int b = threadIdx.x;
    if (uniform) {
      my_total = ((unsigned int)b == udigit) ? (unsigned int)SORT_TILE : 0u;
    } else {
      unsigned int run = 0;
#pragma unroll
      for (int w = 0; w < SORT_WARPS; ++w) { const unsigned int c = s_cnt[w * RADIX + b]; s_cnt[w * RADIX + b] = run; run += c; }
      my_total = run;
    }
    const unsigned long long code_agg = 2ull * parity, code_inc = 2ull * parity + 1ull;
    if (tile + 1 < num_tiles)     // nobody looks at the last tile
      st_desc(desc + (size_t)tile * RADIX + b, ((unsigned long long)my_total << 2) | (tile == 0 ? code_inc : code_agg));
  }
  // tile-local exclusive scan of the 256 digit totals (positions in the reordered tile; a uniform tile is not reordered:
  // every bin starts at 0, only the one digit is used)
  if (uniform) {
    if (threadIdx.x < RADIX) s_bin_start[threadIdx.x] = 0u;
  } else {
    unsigned int inc = my_total;            // threads >= RADIX hold 0
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += up; }
    __shared__ unsigned int s_wsum[SORT_WARPS];
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (threadIdx.x < RADIX) {
      unsigned int wbase = 0;
      for (int w = 0; w < (RADIX / 32); ++w) if (w < warp) wbase += s_wsum[w];
      s_bin_start[threadIdx.x] = wbase + inc - my_total;
    }
  }
  if (threadIdx.x < RADIX) {
    const int b = threadIdx.x;
    unsigned long long excl = 0;
    if (tile > 0) {
      // look back SORT_LB tiles at a time: the descriptor loads of a batch are independent, so a walk of d tiles costs
      // ~d / SORT_LB L2 round trips instead of d (the walk is on the tile's critical path: every other thread waits for it)
      int64_t look = (int64_t)tile - 1;
      bool done = false;
      if constexpr (SORT_LB == 1) {
        for (;;) {
          unsigned long long w;
          do { w = ld_desc(desc + (size_t)look * RADIX + b); } while (((w >> 1) & 1ull) != parity);
          excl += (w >> 2);
          if (w & 1ull) break;          // inclusive prefix of that tile: done
          --look;
        }
        done = true;
      }
      while (!done) {
        unsigned long long w[SORT_LB];
#pragma unroll
        for (int j = 0; j < SORT_LB; ++j)
          w[j] = (look - j >= 0) ? ld_desc(desc + (size_t)(look - j) * RADIX + b) : (2ull * parity + 1ull);   // before tile 0: inclusive 0
#pragma unroll
        for (int j = 0; j < SORT_LB; ++j) {
          if (!done) {
            if (look - j >= 0)
              while (((w[j] >> 1) & 1ull) != parity) w[j] = ld_desc(desc + (size_t)(look - j) * RADIX + b);   // not published yet
            excl += (w[j] >> 2);
            if (w[j] & 1ull) done = true;       // inclusive prefix of that tile
          }
        }
        look -= SORT_LB;
      }
      if (tile + 1 < num_tiles)
        st_desc(desc + (size_t)tile * RADIX + b, ((excl + my_total) << 2) | (2ull * parity + 1ull));
    }
    s_delta[b] = (long long)(g_base[b] + excl) - (long long)s_bin_start[b];
  }
  __syncthreads();

  if (uniform) {
    // ---- the tile moves as one block, in order, straight from the registers it was loaded into
    const long long delta = s_delta[udigit];
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
      const int p = chunk + r * 32;
      keys_out[p + delta] = LAST ? key_decode(key[r]) : key[r];
    }
    if (PAIRS) {
      unsigned long long v[SORT_IPT];
#pragma unroll
      for (int r = 0; r < SORT_IPT; ++r)
        asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v[r]) : "l"(vals_in + tile_base + chunk + r * 32));
#pragma unroll
      for (int r = 0; r < SORT_IPT; ++r) vals_out[chunk + r * 32 + delta] = v[r];
    }
    return;
  }

  // ---- reorder by digit in shared memory (pairs: the positions are kept, two 16-bit values per register)
  unsigned int pos2[PAIRS ? SORT_IPT / 2 : 1];
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    const unsigned int d = key_digit(key[r], shift);
    const unsigned int ps = s_bin_start[d] + my_cnt[d] + rank_get(r);
    if (PAIRS) pos2[r >> 1] = (r & 1) ? (pos2[r >> 1] | (ps << 16)) : ps;
    s_keys[ps] = key[r];
    if (PAIRS) s_digit[ps] = (unsigned char)d;
  }
  // pairs: the values are requested only now, when the key registers are dead (the kernel stays at
  // 64 registers = 2 CTAs per SM); their latency hides behind the key write-out
  unsigned long long val[PAIRS ? SORT_IPT : 1];
  if (PAIRS) {
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
      const int p = chunk + r * 32;
      val[r] = 0ull;
      if (p < valid)
        asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(val[r]) : "l"(vals_in + tile_base + p));
    }
  }
  __syncthreads();

  // ---- write out: contiguous runs per digit
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    const int p = threadIdx.x + r * SORT_BLOCK;
    if (full || p < valid) {
      unsigned long long k = s_keys[p];
      const unsigned int d = key_digit(k, shift);
      if (LAST) k = key_decode(k);
      keys_out[p + s_delta[d]] = k;
    }
  }
  if (PAIRS) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) s_keys[(pos2[r >> 1] >> ((r & 1) * 16)) & 0xffffu] = val[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
      const int p = threadIdx.x + r * SORT_BLOCK;
      if (full || p < valid) vals_out[p + s_delta[s_digit[p]]] = s_keys[p];
    }
  }
}

template <bool PAIRS>
int sort_impl(rpb200_ctx* ctx, double* keys, double* vals, int64_t n, void* scratch, size_t scratch_bytes,
              rpb200_stream_t s)
{
  if (!ctx || n < 0 || (n > 0 && (!keys || (PAIRS && !vals)))) return RPB200_EINVAL;
  if (n <= 1) return 0;
  RPB_CHECK_DEVICE(ctx);
  const sort_scratch_layout L = make_layout(n, PAIRS ? 1 : 0);
  if (!scratch || scratch_bytes < L.total || !rpb_aligned(scratch, 256)) return RPB200_EINVAL;
  if (L.tiles > 0x7ffffff0ll) return RPB200_EINVAL;
  cudaStream_t st = rpb_stream(s);
  unsigned char* base = (unsigned char*)scratch;
  unsigned long long* alt_keys = (unsigned long long*)(base + L.off_alt_keys);
  unsigned long long* alt_vals = (unsigned long long*)(base + L.off_alt_vals);
  unsigned long long* hist = (unsigned long long*)(base + L.off_hist);
  unsigned int* ctrs = (unsigned int*)(base + L.off_ctrs);
  unsigned long long* desc = (unsigned long long*)(base + L.off_desc);
  const unsigned int tiles = (unsigned int)L.tiles;

  // histograms + tickets are cleared per call (18 KB); descriptors only need parity 1 before pass 0,
  // which is what pass 7 of a previous same-shape sort leaves behind -- but scratch is caller-owned
  // and may be fresh, so it is set once per call too (8*RADIX*tiles bytes, ~0.2% of the key traffic)
  RPB_CHECK(cudaMemsetAsync(base + L.off_hist, 0, L.off_desc - L.off_hist, st));
  RPB_CHECK(cudaMemsetAsync(desc, 0xff, sizeof(unsigned long long) * RADIX * (size_t)tiles, st));

  {
    int grid = ctx->sm_count * 4;
    int64_t need = (n + 511) / 512;
    if (need < grid) grid = (int)need;
    if (ctx->tune[PAIRS ? RPB_K_SORTPAIRS : RPB_K_SORT].unroll != 8) {          // default: lane-private counters, one CTA per SM (tuning `unroll` 8: shared bins)
      int g1 = ctx->sm_count;
      const int64_t need1 = (n + HL_BLOCK - 1) / HL_BLOCK;
      if (need1 < g1) g1 = (int)need1;
      RPB_CHECK(cudaFuncSetAttribute(sort_hist_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM));
      sort_hist_lanes_kernel<<<g1, HL_BLOCK, HL_SMEM, st>>>((const unsigned long long*)keys, n, hist, rpb_aligned(keys, 32) ? 1 : 0);
    } else {
      sort_hist_kernel<<<grid, 512, 0, st>>>((const unsigned long long*)keys, n, hist, rpb_aligned(keys, 32) ? 1 : 0);
    }
    RPB_LAUNCH_CHECK();
    sort_hist_scan_kernel<<<NUM_PASSES, RADIX, 0, st>>>(hist, ctrs + 2 * NUM_PASSES, (unsigned long long)n);
    RPB_LAUNCH_CHECK();
  }

  const size_t smem = sizeof(unsigned long long) * SORT_TILE + sizeof(unsigned int) * SORT_WARPS * RADIX +
                      sizeof(unsigned int) * RADIX + sizeof(long long) * RADIX + (PAIRS ? SORT_TILE : 0);
  // skew[p] != 0: pass p tests its tiles for uniformity; skew[NUM_PASSES] is never set (tuning `unroll` 7: no pass tests)
  const unsigned int* skew = ctrs + 2 * NUM_PASSES;
  const bool no_uniform = ctx->tune[PAIRS ? RPB_K_SORTPAIRS : RPB_K_SORT].unroll == 7;
  const int lbt = ctx->tune[PAIRS ? RPB_K_SORTPAIRS : RPB_K_SORT].ctas_per_sm;
  const int lb = (lbt == 1 || lbt == 2) ? lbt : 4;         // tuning `ctas_per_sm` 1 / 2: look back one / two tiles at a time (default 4)
  unsigned long long* kin = (unsigned long long*)keys; unsigned long long* kout = alt_keys;
  unsigned long long* vin = (unsigned long long*)vals; unsigned long long* vout = alt_vals;
  for (int p = 0; p < NUM_PASSES; ++p) {
    const unsigned int parity = (unsigned int)(p & 1);
    const int shift = p * RADIX_BITS;
    const unsigned int* sk = skew + (no_uniform ? NUM_PASSES : p);
    // the attribute is set per call: it belongs to the current device's context, and a process may hold several contexts
#define RPB_SORT_PASS(F, L, LB)                                                                                                  \
    do {                                                                                                                         \
      RPB_CHECK(cudaFuncSetAttribute(sort_onesweep_kernel<PAIRS, F, L, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      sort_onesweep_kernel<PAIRS, F, L, LB><<<tiles, SORT_BLOCK, smem, st>>>(kin, kout, vin, vout, n, shift, hist + p * RADIX, desc, \
                                                                             ctrs + 2 * p, tiles, parity, sk);                  \
    } while (0)
#define RPB_SORT_LB(F, L) do { if (lb == 1) RPB_SORT_PASS(F, L, 1); else if (lb == 2) RPB_SORT_PASS(F, L, 2); else RPB_SORT_PASS(F, L, 4); } while (0)
    if (p == 0) RPB_SORT_LB(true, false);
    else if (p == NUM_PASSES - 1) RPB_SORT_LB(false, true);
    else RPB_SORT_LB(false, false);
#undef RPB_SORT_LB
#undef RPB_SORT_PASS
    RPB_LAUNCH_CHECK();
    unsigned long long* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
  // 8 passes: the result is back in the caller's arrays
  return 0;
}

}  // namespace

extern "C" size_t rpb200_sort_scratch_bytes(int64_t n, int pairs)
{
  if (n < 0) return 0;
  return make_layout(n, pairs).total;
}

extern "C" int rpb200_sort_keys_f64(rpb200_ctx* ctx, double* keys, int64_t n, void* scratch,
                                    size_t scratch_bytes, rpb200_stream_t s)
{ return sort_impl<false>(ctx, keys, nullptr, n, scratch, scratch_bytes, s); }

extern "C" int rpb200_sort_pairs_f64(rpb200_ctx* ctx, double* keys, double* vals, int64_t n, void* scratch,
                                     size_t scratch_bytes, rpb200_stream_t s)
{ return sort_impl<true>(ctx, keys, vals, n, scratch, scratch_bytes, s); }
