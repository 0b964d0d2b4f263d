// halo.cu -- Comm group: HALO_base index lists, HALO_PACKING_FUSED, HALO_EXCHANGE_FUSED for sm_100a.
//
// Replaces comm/HALO_PACKING_FUSED-Cuda.cpp:52-197 and HALO_EXCHANGE_FUSED-Cuda.cpp:52-206:
//   reference                                         here
//   ------------------------------------------------  ----------------------------------------------
//   (buffer,list,var,len) tuples in pinned HOST        tuples + a chunk->tuple map in DEVICE memory
//   memory, read by every CTA over PCIe                (one 4-byte and one 32-byte load per chunk)
//   grid (ceil(avg_len/1024), 78): CTAs of short       the work is cut into 2048-element chunks; a grid of
//   tuples idle, CTAs of long ones loop                4 CTAs per SM takes equal contiguous chunk ranges
//   1 element per thread per loop trip                 8 independent index loads, then 8 gathers (or
//                                                      scatters) in flight per thread, all coalesced
//   cudaStreamSynchronize after pack and after unpack  no host synchronisation at all
//   MPI_Isend/Irecv through pinned host buffers        pack stores straight into the peer GPU's
//                                                      receive buffer over NVLink; per-message
//                                                      release/acquire flags replace MPI_Waitall
#include "common.cuh"

#include <algorithm>
#include <math.h>
#include <new>
#include <stdlib.h>
#include <vector>

namespace {

constexpr int HALO_BLOCK = 256;
constexpr int HALO_CHUNK = 2048;      // elements per CTA: 8 per thread
constexpr int NNB = RPB200_HALO_NEIGHBORS;

// receive buffers are written by a peer GPU while this kernel may be resident: never through L1
__device__ __forceinline__ double ld_cg(const double* p)
{
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// L2 eviction priorities.  A face whose cells are strided in the variable (every list with an x
// offset: stride nx+2h) touches ONE 8-byte cell per 32-byte sector, and the ghost cell the unpack
// writes shares that sector with the owned cell the pack reads.  Those sectors (2 x (ny+2h)(nz+2h) x
// 32 B per variable: 51 MB for 3 variables at 512^3) fit the 126 MB L2, so they are kept with
// evict_last while everything that streams (contiguous faces, buffers, lists) goes evict_first:
// from the second rep on the strided faces are L2 hits and the partial-sector ghost writes never
// become DRAM read-modify-writes.
// (scalar ld/st take the priority as a createpolicy operand: the .L2::evict_* qualifiers are only
// accepted on 256-bit accesses for sm_100)
__device__ __forceinline__ unsigned long long policy_keep()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_once()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ld_hint(const double* p, unsigned long long pol)
{
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_hint(double* p, double v, unsigned long long pol)
{
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ int ld_hint_i32(const int* p, unsigned long long pol)
{
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

constexpr int SEG_STRIDED = 1;       // rpb200_halo_seg.flags bit 0 (set by the library)
constexpr int SEG_IDENTITY = 2;      // bit 1: no index list, element i of the segment is var[i] (HALO_SENDRECV puts)

// per-message signalling state of an exchange (device)
struct halo_msg {
  unsigned long long* remote_flag;   // pack: flag of the receive slot in the DESTINATION window
  const unsigned long long* my_flag; // unpack: flag of this slot in MY window
  unsigned int chunks;               // pack chunks of this message (all variables)
  unsigned int pad;
};

// MODE 0: plain;  MODE 1: pack + signal;  MODE 2: wait + unpack
//
// Every CTA owns a CONTIGUOUS range of chunks (so it walks the tuples in message order, stays inside
// a few pages, and in MODE 1 needs ONE system-scope fence for everything it wrote).  Inside a chunk
// lane t handles elements t, t+256, ...: the index loads, the buffer side and -- for every face whose
// cells are contiguous -- the variable side are all fully coalesced (a thread owning 4 consecutive
// elements would turn each of its 4 scatter instructions into 32 partial-sector writes).
template <bool PACK, int MODE, bool HINT, bool STRIDED>
__global__ void __launch_bounds__(HALO_BLOCK)
halo_kernel(const rpb200_halo_seg* __restrict__ segs_g0, const rpb200_halo_seg* __restrict__ segs_g1,
            const int* __restrict__ chunk_seg, const long long* __restrict__ seg_first_chunk, int total_chunks,
            const halo_msg* __restrict__ msgs, unsigned int* __restrict__ msg_done,
            unsigned long long* __restrict__ d_epoch, unsigned int* __restrict__ unpack_done,
            int* __restrict__ error, int chunk_lo, int commit, int reverse, unsigned long long timeout_ns)
{
  __shared__ int s_ok;
  // reverse (pack launches): walk the chunks from the end of the work list, so that the strided x faces -- first in the
  // list -- are packed LAST and are the most recently used L2 lines when the unpack, which walks forward, starts with
  // the ghost cells that share those lines (LIFO reuse across the two launches).
  // The launch covers chunks [chunk_lo, total_chunks) of the work list (the whole list for the fused kernels, one
  // (neighbour, variable) tuple for the unfused HALO_EXCHANGE); `commit` = this is the last unpack launch of the rep.
  constexpr int EPT = HALO_CHUNK / HALO_BLOCK;     // 8 elements per thread per chunk
  // The exchange epoch lives in device memory (so a CUDA graph of many reps replays correctly): the
  // pack and the unpack of one rep both see *d_epoch + 1; the last unpack CTA to retire commits it.
  // Its parity selects the receive-buffer generation.
  unsigned long long epoch = 0;
  if (MODE != 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(epoch) : "l"(d_epoch) : "memory");
  epoch += 1;
  const rpb200_halo_seg* __restrict__ segs = (MODE != 0 && (epoch & 1ull)) ? segs_g1 : segs_g0;
  const int per = (total_chunks - chunk_lo + gridDim.x - 1) / gridDim.x;
  const int c_begin = chunk_lo + (STRIDED ? (int)blockIdx.x : blockIdx.x * per);
  const int c_end = STRIDED ? total_chunks : min(c_begin + per, total_chunks);
  const int c_step = STRIDED ? (int)gridDim.x : 1;
  int waited_msg = -1;
  bool timed_out = false;
  const unsigned long long pol_keep = HINT ? policy_keep() : 0ull, pol_once = HINT ? policy_once() : 0ull;

  const int c_flip = chunk_lo + total_chunks - 1;          // reverse: chunk c stands for chunk c_flip - c
  for (int cc = c_begin; cc < c_end; cc += c_step) {
    const int c = reverse ? c_flip - cc : cc;
    const int s = __ldg(chunk_seg + c);
    const rpb200_halo_seg seg = segs[s];
    const int64_t i0 = ((int64_t)c - __ldg(seg_first_chunk + s)) * HALO_CHUNK;
    const int cnt = (int)((seg.len - i0) < HALO_CHUNK ? (seg.len - i0) : HALO_CHUNK);

    if (MODE == 2 && seg.msg != waited_msg) {      // first chunk of a message this CTA touches: acquire it
      if (threadIdx.x == 0) {
        const unsigned long long* f = msgs[seg.msg].my_flag;
        unsigned long long t0 = 0, t1 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        int ok = 1;
        while (ld_acquire_sys(f) < epoch) {
          __nanosleep(40);
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > timeout_ns) { atomicExch(error, RPB200_ETIMEDOUT); ok = 0; break; }   // wall-clock bound: give up loudly
        }
        s_ok = ok;
      }
      __syncthreads();
      timed_out = timed_out || (s_ok == 0);
      __syncthreads();
      waited_msg = seg.msg;
    }
    if (MODE == 2 && timed_out) continue;          // a message that did not arrive is NOT unpacked (and the epoch not committed)

    const int* __restrict__ list = seg.list + i0;
    double* __restrict__ buf = seg.buffer + i0;
    double* __restrict__ var = seg.var;
    int idx[EPT];
    double v[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int i = k * HALO_BLOCK + threadIdx.x;
      idx[k] = (i < cnt) ? ((seg.flags & SEG_IDENTITY) ? (int)(i0 + i) : (HINT ? ld_hint_i32(list + i, pol_once) : __ldg(list + i))) : -1;
    }
    const bool keep = HINT && (seg.flags & SEG_STRIDED);      // uniform over the chunk
    if (PACK) {
      if (keep) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) v[k] = ld_hint(var + idx[k], pol_keep);
      } else if (HINT) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) v[k] = ld_hint(var + idx[k], pol_once);
      } else {
        // .cg: a gather through L1 fetches TWO sectors from L2 per miss (64-byte L1 fill): for a strided face that is 64
        // bytes of DRAM traffic per 8-byte cell (ncu, profiles/r02_d/: 6.56 M L2 read sectors for 2.98 M L1 miss sectors)
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) v[k] = ld_cg(var + idx[k]);
      }
      // MODE 1: the buffer is the peer's receive slot, read back by its unpack right away: default policy
      if (HINT && MODE == 0) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) st_hint(buf + k * HALO_BLOCK + threadIdx.x, v[k], pol_once);
      } else {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) buf[k * HALO_BLOCK + threadIdx.x] = v[k];
      }
    } else {
      if (HINT && MODE == 0) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) v[k] = ld_hint(buf + k * HALO_BLOCK + threadIdx.x, pol_once);
      } else {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) v[k] = ld_cg(buf + k * HALO_BLOCK + threadIdx.x);
      }
      if (keep) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) st_hint(var + idx[k], v[k], pol_keep);
      } else if (HINT) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) st_hint(var + idx[k], v[k], pol_once);
      } else {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (idx[k] >= 0) var[idx[k]] = v[k];
      }
    }
  }

  if (MODE == 1 && c_begin < c_end) {   // (STRIDED: blockIdx.x < total_chunks always holds, the grid is clamped)
    // all stores of this CTA -> barrier -> ONE system fence -> credit every message it touched; whoever
    // completes a message publishes the epoch to the destination's flag (release at system scope)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      int c = c_begin;
      while (c < c_end) {
        const int m = segs[__ldg(chunk_seg + (reverse ? c_flip - c : c))].msg;
        int run = 1;
        while (c + run * c_step < c_end && segs[__ldg(chunk_seg + (reverse ? c_flip - (c + run * c_step) : c + run * c_step))].msg == m) ++run;
        const halo_msg hm = msgs[m];
        const unsigned int prev = atomicAdd(msg_done + m, (unsigned int)run);
        if (prev + run == hm.chunks) {
          msg_done[m] = 0u;              // re-armed for the next rep (stream-ordered launches)
          st_release_sys(hm.remote_flag, epoch);
        }
        c += run * c_step;
      }
    }
  }
  if (MODE == 2 && commit) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int prev = atomicAdd(unpack_done, 1u);
      if (prev == gridDim.x - 1) {       // every CTA has read the epoch and finished: commit it (unless a message timed out)
        *unpack_done = 0u;
        __threadfence();
        int err = 0;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(err) : "l"(error) : "memory");
        if (err == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(d_epoch), "l"(epoch) : "memory");
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The ITEM-LIST kernel: HALO_PACKING_FUSED as ONE launch (pack and unpack items interleaved), and HALO_EXCHANGE_FUSED as ONE
// launch (all pack items, signal, then wait + unpack items).
//
// Why.  ncu on the two-launch form (profiles/r01_halo_l2_forward_vs_backward.csv): 320 MB of DRAM traffic per rep for 190 MB
// of algorithmic bytes, at half of the DRAM bandwidth.  The waste is all in the +-x faces.  A cell of the -x / +x face is one
// 8-byte word per (j, k) row of the variable, 4112 bytes from the next one: every access is its own DRAM burst (64 bytes
// fetched for 8 used), and the pack launch (owned cells i = 1, i = nx) and the unpack launch (ghost cells i = 0, i = nx + 1)
// each pay it, the unpack as a read-modify-write of a partial sector.  But in memory the four cells
//      (nx, j, k)  (nx+1, j, k)  (0, j+1, k)  (1, j+1, k)      =  +x owned, +x ghost, -x ghost, -x owned
// are CONSECUTIVE (the end of row j and the start of row j + 1): 32 bytes, one sector when j is even, two adjacent sectors
// when j is odd.  So element n of the four tuples  pack(-x), pack(+x), unpack(-x), unpack(+x)  of one variable touch the same
// one or two sectors.  The item list therefore groups chunk c of those four tuples into one UNIT, and a unit is the grain of
// dealing: ONE CTA moves its four chunks back to back, so thread t touches (1, j_t, k), (nx, j_t, k), (0, j_t, k),
// (nx+1, j_t, k) within a microsecond -- the second to fourth access find the sector in this SM's L1 or in L2, and the
// strided faces cost one DRAM burst read + one write-back per row instead of four bursts (one a read-modify-write).
// (Round 2, call b measured the same four chunks dealt to four DIFFERENT CTAs: 108 us against 96 us for two launches at 512^3
// -- adjacency in the list is not locality on the chip; profiles/r02_b/.)  Every other chunk is a unit of its own.
// Pack and unpack of HALO_PACKING_FUSED touch disjoint cells (owned / ghost) and disjoint buffers
// (HALO_PACKING_FUSED-Seq.cpp:43-97), so any interleaving computes the reference's result.
//
// Dealing is round-robin over units (unit u -> CTA u mod grid), which spreads the slow strided units and the streaming ones
// evenly over every CTA.  A thread keeps the index loads of its NEXT item in flight while it gathers / scatters the current one.
// seg: bits 0-7 first tuple, bits 8-15 number of CONSECUTIVE tuples that share one index list (the variables of a message:
// the indices of a chunk are loaded once and used for every variable), bit 30 set = unpack side
struct halo_item { int seg; int chunk; };
constexpr int ITEM_UNPACK = 1 << 30;
__host__ __device__ inline int item_first(int seg) { return seg & 0xff; }
__host__ __device__ inline int item_count(int seg) { return (seg >> 8) & 0xff; }
inline int item_make(int first, int count, int bit) { return first | (count << 8) | bit; }
constexpr int ITEMS_MAX_SEGS = 128;                  // tuples per side kept in shared memory (26 neighbours x <= 4 variables); < 256

struct halo_items_args {
  const rpb200_halo_seg* psegs[2];                   // pack tuples, generation 0 / 1 (the same array for HALO_PACKING_FUSED)
  const rpb200_halo_seg* usegs[2];
  const halo_item* items;
  const int* unit_first;                             // unit u = items [unit_first[u], unit_first[u + 1])
  unsigned int* ticket;                              // [0] phase-1 unit ticket, [1] phase-2 unit ticket, [2] CTAs done (re-armed by the last)
  int n_units, n_pack_units, npsegs, nusegs;         // XCHG: units [0, n_pack_units) are the pack phase
  int progressive;                                   // XCHG: ONE ticket over pack and unpack units, messages signalled unit by unit
  const halo_msg* pmsgs; const halo_msg* umsgs;
  unsigned int* msg_done; unsigned long long* d_epoch; unsigned int* unpack_done; int* error;
  unsigned long long timeout_ns;
};

template <bool XCHG>
__global__ void __launch_bounds__(HALO_BLOCK, 4)
halo_items_kernel(const __grid_constant__ halo_items_args A)
{
  constexpr int EPT = HALO_CHUNK / HALO_BLOCK;
  __shared__ rpb200_halo_seg s_seg[2][ITEMS_MAX_SEGS];
  __shared__ unsigned int s_credit[NNB];
  __shared__ int s_ok;
  unsigned long long epoch = 0;
  if (XCHG) {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(epoch) : "l"(A.d_epoch) : "memory");
    epoch += 1;
  }
  const int gen = XCHG ? (int)(epoch & 1ull) : 0;
  {
    const rpb200_halo_seg* __restrict__ ps = A.psegs[gen];
    const rpb200_halo_seg* __restrict__ us = A.usegs[gen];
    for (int i = threadIdx.x; i < A.npsegs; i += HALO_BLOCK) s_seg[0][i] = ps[i];
    for (int i = threadIdx.x; i < A.nusegs; i += HALO_BLOCK) s_seg[1][i] = us[i];
    if (threadIdx.x < NNB) s_credit[threadIdx.x] = 0u;
    if (threadIdx.x == 0) s_ok = 1;
  }
  __syncthreads();

  int idx[EPT], nidx[EPT];
  // the index loads of item `it` (nothing when it < 0)
  auto fetch = [&](int it, int (&dst)[EPT], halo_item& h) {
    if (it >= 0) {
      h = A.items[it];
      const rpb200_halo_seg& seg = s_seg[(h.seg & ITEM_UNPACK) ? 1 : 0][item_first(h.seg)];
      const int64_t i0 = (int64_t)h.chunk * HALO_CHUNK;
      const int cnt = (int)((seg.len - i0) < HALO_CHUNK ? (seg.len - i0) : HALO_CHUNK);
      const int* __restrict__ list = seg.list + i0;
#pragma unroll
      for (int k = 0; k < EPT; ++k) { const int i = k * HALO_BLOCK + threadIdx.x; dst[k] = (i < cnt) ? __ldg(list + i) : -1; }
    }
  };
  // one chunk of every variable of the item, with the SAME indices.  Gathers bypass L1 (.cg): through L1 a strided 8-byte
  // load fetches two sectors from L2.
  auto move = [&](const halo_item h, const int (&ix)[EPT]) {
    const bool unpack = (h.seg & ITEM_UNPACK) != 0;
    const int first = item_first(h.seg), count = item_count(h.seg);
    for (int j = 0; j < count; ++j) {
      const rpb200_halo_seg& seg = s_seg[unpack ? 1 : 0][first + j];
      double* __restrict__ buf = seg.buffer + (int64_t)h.chunk * HALO_CHUNK;
      double* __restrict__ var = seg.var;
      double v[EPT];
      if (!unpack) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (ix[k] >= 0) v[k] = ld_cg(var + ix[k]);
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (ix[k] >= 0) buf[k * HALO_BLOCK + threadIdx.x] = v[k];
      } else {
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (ix[k] >= 0) v[k] = ld_cg(buf + k * HALO_BLOCK + threadIdx.x);
#pragma unroll
        for (int k = 0; k < EPT; ++k) if (ix[k] >= 0) var[ix[k]] = v[k];
      }
    }
  };

  // Units are drawn from an atomic ticket, two ahead: while unit u is being moved the next unit is already known (its first
  // index loads are issued during u's last item) and the ticket after that is in flight.  Heavy units (the x faces) and light
  // ones therefore balance themselves, whatever the grid.  `tk` is re-armed by the last CTA of the launch (below).
  __shared__ unsigned int s_tk[2];
  // credit the messages this CTA packed since the last call (thread 0, after a CTA barrier that follows the stores): ONE
  // system fence, then whoever completes a message publishes the epoch to the destination's flag (release at system scope)
  auto flush_credits = [&]() {
    __threadfence_system();
    for (int m = 0; m < NNB; ++m) {
      const unsigned int mine = s_credit[m];
      if (mine == 0u) continue;
      s_credit[m] = 0u;
      const halo_msg hm = A.pmsgs[m];
      const unsigned int prev = atomicAdd(A.msg_done + m, mine);
      if (prev + mine == hm.chunks) {
        A.msg_done[m] = 0u;              // re-armed for the next rep (stream-ordered launches)
        st_release_sys(hm.remote_flag, epoch);
      }
    }
  };
  // mode 0: pack units (never waits); 1: unpack units (acquire each message's flag once per CTA); 2: both kinds from ONE
  // ticket, pack units first -- a pack unit's messages are credited as soon as the unit is stored (thread 0, while the other
  // warps go on), so flags are released message by message and the unpack of early messages overlaps the packing (and the
  // NVLink transfer) of late ones.  No deadlock: tickets increase, so every pack unit is held by a CTA that has not drawn an
  // unpack unit yet, pack units never wait, and the grid is co-resident.
  auto run_phase = [&](unsigned int* tk, int u_lo, int u_end, int mode, unsigned int& acquired, bool& failed) {
    if (threadIdx.x == 0) { s_tk[0] = atomicAdd(tk, 1u); s_tk[1] = atomicAdd(tk, 1u); }
    __syncthreads();
    int u_cur = u_lo + (int)s_tk[0], u_nxt = u_lo + (int)s_tk[1];
    __syncthreads();                                         // both slots are read before slot 0 is rewritten
    halo_item h, nh;
    int it = -1, end_it = 0;
    if (u_cur < u_end) { it = __ldg(A.unit_first + u_cur); end_it = __ldg(A.unit_first + u_cur + 1); }
    fetch(u_cur < u_end ? it : -1, idx, h);
    for (int k = 0; u_cur < u_end; ++k) {
      if (threadIdx.x == 0) s_tk[k & 1] = atomicAdd(tk, 1u);          // the unit after u_nxt
      int n_it = -1, n_end = 0;
      if (u_nxt < u_end) { n_it = __ldg(A.unit_first + u_nxt); n_end = __ldg(A.unit_first + u_nxt + 1); }
      const bool wait_flags = mode == 1 || (mode == 2 && u_cur >= A.n_pack_units);
      for (; it < end_it; ++it) {
        fetch(it + 1 < end_it ? it + 1 : n_it, nidx, nh);             // next item of this unit, else the first of the next unit
        if (wait_flags) {
          const int m = s_seg[1][item_first(h.seg)].msg;
          if (!((acquired >> m) & 1u)) {       // first item of message m in this CTA: acquire its flag
            if (threadIdx.x == 0) {
              const unsigned long long* f = A.umsgs[m].my_flag;
              unsigned long long t0 = 0, t1 = 0;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
              int ok = 1;
              while (ld_acquire_sys(f) < epoch) {
                __nanosleep(40);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > A.timeout_ns) { atomicExch(A.error, RPB200_ETIMEDOUT); ok = 0; break; }
              }
              s_ok = ok;
            }
            __syncthreads();
            failed = failed || (s_ok == 0);
            __syncthreads();
            acquired |= 1u << m;
          }
        }
        if (!failed) move(h, idx);
        if (XCHG && !wait_flags && threadIdx.x == 0) s_credit[s_seg[0][item_first(h.seg)].msg] += (unsigned int)item_count(h.seg);
#pragma unroll
        for (int q = 0; q < EPT; ++q) idx[q] = nidx[q];
        h = nh;
      }
      __syncthreads();                                       // s_tk[k & 1] is visible; everybody is done with unit u_cur
      if (XCHG && mode == 2 && !wait_flags && threadIdx.x == 0) flush_credits();
      u_cur = u_nxt; it = n_it; end_it = n_end;
      u_nxt = u_lo + (int)s_tk[k & 1];
    }
  };

  unsigned int acquired = 0u;            // bit m: this CTA has already seen message m's flag (thread-uniform)
  bool failed = false;
  // ---- phase 1: every unit (HALO_PACKING_FUSED) / the pack units (exchange)
  if (XCHG && A.progressive) {
    run_phase(A.ticket + 0, 0, A.n_units, 2, acquired, failed);
  } else {
  run_phase(A.ticket + 0, 0, XCHG ? A.n_pack_units : A.n_units, 0, acquired, failed);
  if (!XCHG) {
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(A.ticket + 2, 1u);
      if (prev == gridDim.x - 1) { A.ticket[0] = 0u; A.ticket[2] = 0u; }      // every CTA has drawn its last ticket: re-arm
    }
    return;
  }

  // all remote stores of this CTA -> barrier -> ONE system fence -> credit every message it touched
  __syncthreads();
  if (threadIdx.x == 0) flush_credits();

  // ---- phase 2: wait + unpack.  The grid is fully co-resident and phase 1 never waits, so every rank's flags are
  // eventually released.  A message that does not arrive within the time-out is NOT unpacked and the epoch is NOT committed.
  run_phase(A.ticket + 1, A.n_pack_units, A.n_units, 1, acquired, failed);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                     // this CTA's time-out report (if any) is visible before its retirement is counted
    const unsigned int prev = atomicAdd(A.unpack_done, 1u);
    if (prev == gridDim.x - 1) {         // every CTA has read the epoch, drawn its last tickets and finished
      *A.unpack_done = 0u;
      A.ticket[0] = 0u; A.ticket[1] = 0u;
      __threadfence();
      int err = 0;
      asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(err) : "l"(A.error) : "memory");
      if (err == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(A.d_epoch), "l"(epoch) : "memory");
    }
  }
}

// HALO_SENDRECV's receive side: wait for the 26 messages of this rep, then commit the epoch
__global__ void halo_wait_kernel(const halo_msg* __restrict__ umsgs, unsigned long long* __restrict__ d_epoch, int* __restrict__ error,
                                 unsigned long long timeout_ns)
{
  unsigned long long epoch = 0;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(epoch) : "l"(d_epoch) : "memory");
  epoch += 1;
  int ok = 1;
  if (threadIdx.x < NNB) {
    const unsigned long long* f = umsgs[threadIdx.x].my_flag;
    unsigned long long t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(f) < epoch) {
      __nanosleep(40);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) { atomicExch(error, RPB200_ETIMEDOUT); ok = 0; break; }
    }
  }
  ok = __all_sync(0xffffffffu, ok);      // a rep with a missing message is not committed
  if (threadIdx.x == 0 && ok) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(d_epoch), "l"(epoch) : "memory");
}

struct worklist_dev {
  rpb200_halo_seg* d_segs = nullptr;
  int* d_chunk_seg = nullptr;
  long long* d_first = nullptr;
  rpb200_halo_seg* h_stage = nullptr;   // pinned staging for update()
  int nsegs = 0;
  int64_t total_chunks = 0;
  std::vector<int64_t> lens;
  std::vector<int> flags;
  std::vector<long long> first;     // host copy: first chunk of every tuple
  std::vector<rpb200_halo_seg> h_segs;   // host copy of the tuples as built (geometry: len, msg, flags, var)
  // item list of the one-launch pack+unpack kernel, cached on the PACK list for the unpack list it was merged with
  void* d_items = nullptr; const halo_item* d_item_list = nullptr; const int* d_unit_first = nullptr; unsigned int* d_unit_ticket = nullptr;
  int n_units = 0;
  const void* merged_with = nullptr; int merged_order = -1;
};

int worklist_free(worklist_dev& w)
{
  cudaFree(w.d_segs); cudaFree(w.d_chunk_seg); cudaFree(w.d_first); cudaFree(w.d_items);
  if (w.h_stage) cudaFreeHost(w.h_stage);
  w = worklist_dev();
  return 0;
}

// flags bit 0: the segment's cells are strided in the variable (first two list entries not adjacent).
// Setup-time only (one 8-byte D2H copy per tuple); a performance hint, never a correctness input.
int classify(rpb200_halo_seg& s)
{
  if (s.flags & SEG_IDENTITY) { s.flags = SEG_IDENTITY; return 0; }      // set by the library only
  s.flags = 0;
  if (s.len >= 2) {
    int two[2];
    RPB_CHECK(cudaMemcpy(two, s.list, sizeof(two), cudaMemcpyDeviceToHost));
    if (two[1] - two[0] != 1) s.flags |= SEG_STRIDED;
  }
  return 0;
}

int worklist_build(worklist_dev& w, const rpb200_halo_seg* h_segs, int nsegs)
{
  if (nsegs < 0 || (nsegs > 0 && !h_segs)) return RPB200_EINVAL;
  w.nsegs = nsegs;
  std::vector<rpb200_halo_seg> segs(h_segs, h_segs + nsegs);
  std::vector<long long> first(nsegs > 0 ? nsegs : 1, 0);
  std::vector<int> map;
  w.lens.resize(nsegs);
  int64_t chunks = 0;
  for (int s = 0; s < nsegs; ++s) {
    if (segs[s].len < 0 || (segs[s].len > 0 && (!segs[s].buffer || !segs[s].var || (!segs[s].list && !(segs[s].flags & SEG_IDENTITY))))) return RPB200_EINVAL;
    { const int rc = classify(segs[s]); if (rc != 0) return rc; }
    w.flags.push_back(segs[s].flags);
    w.lens[s] = segs[s].len;
    first[s] = chunks;
    const int64_t nc = (segs[s].len + HALO_CHUNK - 1) / HALO_CHUNK;
    for (int64_t k = 0; k < nc; ++k) map.push_back(s);
    chunks += nc;
  }
  if (chunks > 0x7fffffffll) return RPB200_EINVAL;
  w.total_chunks = chunks;
  w.first = first;
  w.h_segs = segs;
  const size_t nb = sizeof(rpb200_halo_seg) * (size_t)(nsegs > 0 ? nsegs : 1);
  RPB_CHECK(cudaMalloc(&w.d_segs, nb));
  RPB_CHECK(cudaMallocHost(&w.h_stage, nb));
  RPB_CHECK(cudaMalloc(&w.d_first, sizeof(long long) * first.size()));
  RPB_CHECK(cudaMalloc(&w.d_chunk_seg, sizeof(int) * (map.size() ? map.size() : 1)));
  if (nsegs > 0) {
    RPB_CHECK(cudaMemcpy(w.d_segs, segs.data(), sizeof(rpb200_halo_seg) * nsegs, cudaMemcpyHostToDevice));
    RPB_CHECK(cudaMemcpy(w.d_first, first.data(), sizeof(long long) * nsegs, cudaMemcpyHostToDevice));
  }
  if (!map.empty()) RPB_CHECK(cudaMemcpy(w.d_chunk_seg, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice));
  return 0;
}

// how long an unpack CTA waits for a message before it reports RPB200_ETIMEDOUT (wall clock, %globaltimer);
// RPB200_HALO_TIMEOUT_MS overrides the 2 s default
unsigned long long halo_timeout_ns()
{
  static unsigned long long ns = 0;
  if (ns == 0) {
    const char* e = getenv("RPB200_HALO_TIMEOUT_MS");
    const long ms = e ? atol(e) : 2000;
    ns = (unsigned long long)(ms > 0 ? ms : 2000) * 1000000ull;
  }
  return ns;
}

struct exchange_args {
  const worklist_dev* other_gen = nullptr;   // generation-1 tuples (same chunk map)
  const halo_msg* msgs = nullptr;
  unsigned int* msg_done = nullptr;
  unsigned long long* d_epoch = nullptr;
  unsigned int* unpack_done = nullptr;
  int* error = nullptr;
};

// two-launch exchange, ctas_per_sm left automatic: launches over more chunks than this run 2 CTAs per SM instead of 4
constexpr int64_t HALO_AUTO_CPS_CHUNKS = 5000;

// seg_lo < 0: the whole work list; else only the tuples [seg_lo, seg_hi)
template <bool PACK, int MODE>
int worklist_launch(const rpb200_ctx* ctx, int kid, const worklist_dev& w, const exchange_args& x, cudaStream_t st,
                    int seg_lo = -1, int seg_hi = -1, int commit = 1)
{
  if (w.total_chunks == 0) return 0;
  int64_t chunk_lo = 0, chunk_hi = w.total_chunks;
  if (seg_lo >= 0) {
    if (seg_hi > w.nsegs || seg_lo >= seg_hi) return RPB200_EINVAL;
    chunk_lo = w.first[seg_lo];
    chunk_hi = seg_hi < w.nsegs ? w.first[seg_hi] : w.total_chunks;
    if (chunk_hi == chunk_lo && !(MODE == 2 && commit)) return 0;        // empty tuples: nothing to launch
  }
  // ctas_per_sm 0 = automatic: 4 CTAs per SM, except for exchange launches over more than HALO_AUTO_CPS_CHUNKS chunks, which
  // take 2 -- measured with both settings on the same boxes (profiles/r02_f/, profiles/r02_mgpu/): 512^3 cells per rank
  // (2313 chunks) 100-106 us at 4 per SM against 110-118 us at 2; 1024^3 (9234 chunks) 504 against 437 us on one rank and
  // 547 against 549 us on 2 x 2 x 2 ranks.
  int cps = ctx->tune[kid].ctas_per_sm;
  if (cps <= 0) cps = (kid == RPB_K_HALO_EXCHANGE_FUSED && chunk_hi - chunk_lo > HALO_AUTO_CPS_CHUNKS) ? 2 : 4;
  int64_t grid = (int64_t)ctx->sm_count * cps;
  if (grid > chunk_hi - chunk_lo) grid = chunk_hi - chunk_lo;
  if (grid < 1) grid = 1;
  // tuning field `unroll` of the two halo kernels: 4 = L2 eviction-priority hints on, else off;
  // tuning field `block_size`: 128 = chunks dealt round-robin (c = b, b + grid, ...) instead of in
  // contiguous ranges, so the slow strided faces are spread over every CTA
  const bool hint = ctx->tune[kid].unroll == 4, strided = ctx->tune[kid].block_size == 128 || ctx->tune[kid].block_size == 160;
  // block_size 192: contiguous ranges, packs walk backwards; 160: round-robin, packs walk backwards
  const int reverse = (PACK && (ctx->tune[kid].block_size == 192 || ctx->tune[kid].block_size == 160)) ? 1 : 0;
#define RPB_HALO_LAUNCH(H, S)                                                                                  \
  halo_kernel<PACK, MODE, H, S><<<(int)grid, HALO_BLOCK, 0, st>>>(                                             \
      w.d_segs, x.other_gen ? x.other_gen->d_segs : w.d_segs, w.d_chunk_seg, w.d_first, (int)chunk_hi,         \
      x.msgs, x.msg_done, x.d_epoch, x.unpack_done, x.error, (int)chunk_lo, commit, reverse, halo_timeout_ns())
  if (hint && strided) RPB_HALO_LAUNCH(true, true);
  else if (hint) RPB_HALO_LAUNCH(true, false);
  else if (strided) RPB_HALO_LAUNCH(false, true);
  else RPB_HALO_LAUNCH(false, false);
#undef RPB_HALO_LAUNCH
  RPB_LAUNCH_CHECK();
  return 0;
}

// comm/HALO_base.cpp:82-116
const int k_offsets[NNB][3] = {
  {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1},
  {-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0},
  {-1, 0, -1}, {-1, 0, 1}, {1, 0, -1}, {1, 0, 1},
  {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1},
  {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1},
  {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}
};

struct box { int64_t lo[3], hi[3]; int64_t len() const { return (hi[0]-lo[0])*(hi[1]-lo[1])*(hi[2]-lo[2]); } };

// comm/HALO_base.cpp:118-166: the send box is the outermost `hw` owned layers, the recv box the ghost layers
box make_box(bool recv, const int off[3], int64_t hw, const int64_t dims[3])
{
  box b;
  for (int a = 0; a < 3; ++a) {
    if (off[a] < 0)      { b.lo[a] = recv ? 0 : hw;                 b.hi[a] = b.lo[a] + hw; }
    else if (off[a] > 0) { b.lo[a] = recv ? hw + dims[a] : dims[a]; b.hi[a] = b.lo[a] + hw; }
    else                 { b.lo[a] = hw;                            b.hi[a] = hw + dims[a]; }
  }
  return b;
}

// the index lists are generated on the device (comm/HALO_base.cpp:228-254, 262-288: k outer, i inner)
__global__ void halo_fill_list_kernel(int* __restrict__ list, int64_t len, int64_t lo0, int64_t lo1, int64_t lo2,
                                      int64_t e0, int64_t e1, int64_t sj, int64_t sk)
{
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < len; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = n % e0, t = n / e0, j = t % e1, k = t / e1;
    list[n] = (int)((lo0 + i) + (lo1 + j) * sj + (lo2 + k) * sk);
  }
}

}  // namespace

struct rpb200_halo_worklist { worklist_dev w; };

struct rpb200_halo_plan {
  rpb200_ctx* ctx = nullptr;
  int64_t dims[3] = {0, 0, 0}, hw = 0, var_size = 0;
  int nvars = 0, my_rank = 0, rank_dims[3] = {1, 1, 1}, nranks = 1;
  int ranks[NNB], send_tags[NNB], recv_tags[NNB], opposite[NNB];
  int64_t pack_len[NNB], unpack_len[NNB];
  int* d_pack_list[NNB];
  int* d_unpack_list[NNB];
  int* d_lists = nullptr;                  // one allocation, 256-byte aligned sub-lists
  // HALO_PACKING_FUSED binding
  worklist_dev pack_wl, unpack_wl;
  bool bound = false;
  // HALO_EXCHANGE_FUSED
  std::vector<double*> vars;
  unsigned char* d_window = nullptr; size_t window_bytes = 0;
  size_t recv_off[2][NNB];                 // byte offsets of receive slot l, generation g, in ANY rank's window
  std::vector<void*> peer_windows; std::vector<bool> peer_opened;
  worklist_dev xpack_wl[2], xunpack_wl[2];
  worklist_dev xsend_wl[2];                // HALO_SENDRECV: identity puts of the caller's send buffers
  bool send_bound = false;
  halo_msg* d_pack_msgs = nullptr; halo_msg* d_unpack_msgs = nullptr; halo_msg* d_send_msgs = nullptr;
  unsigned int* d_msg_done = nullptr;
  int* d_error = nullptr;
  unsigned long long* d_epoch = nullptr;   // committed exchange epoch (device): reps done so far
  unsigned int* d_unpack_done = nullptr;
  bool connected = false;
  // item lists of the one-launch kernels (halo_items_kernel): HALO_PACKING_FUSED merged, exchange pack-then-unpack
  void* d_xchg_block = nullptr; const halo_item* d_items_xchg = nullptr; const int* d_unit_first_xchg = nullptr;
  unsigned int* d_ticket_xchg = nullptr;
  // the progressive (one-ticket) exchange order
  void* d_prog_block = nullptr; const halo_item* d_items_prog = nullptr; const int* d_unit_first_prog = nullptr;
  unsigned int* d_ticket_prog = nullptr; int n_units_prog = 0, n_pack_units_prog = 0;
  int n_units_xchg = 0, n_pack_units_xchg = 0;
};

extern "C" int rpb200_halo_chunk(void) { return HALO_CHUNK; }

extern "C" int rpb200_halo_worklist_create(rpb200_ctx* ctx, const rpb200_halo_seg* h_segs, int nsegs,
                                           rpb200_halo_worklist** out)
{
  if (!ctx || !out) return RPB200_EINVAL;
  *out = nullptr;
  rpb200_halo_worklist* wl = new (std::nothrow) rpb200_halo_worklist();
  if (!wl) return (int)cudaErrorMemoryAllocation;
  std::vector<rpb200_halo_seg> clean(h_segs ? h_segs : nullptr, h_segs ? h_segs + (nsegs > 0 ? nsegs : 0) : nullptr);
  for (rpb200_halo_seg& c : clean) c.flags = 0;                 // "filled by the library; pass 0"
  const int rc = worklist_build(wl->w, clean.data(), nsegs);
  if (rc != 0) { worklist_free(wl->w); delete wl; return rc; }
  *out = wl;
  return 0;
}

extern "C" int rpb200_halo_worklist_update(rpb200_halo_worklist* wl, const rpb200_halo_seg* h_segs, int nsegs,
                                           rpb200_stream_t s)
{
  if (!wl || nsegs != wl->w.nsegs || (nsegs > 0 && !h_segs)) return RPB200_EINVAL;
  for (int i = 0; i < nsegs; ++i) {
    if (h_segs[i].len != wl->w.lens[i]) return RPB200_EINVAL;    // the chunk map depends on the lengths
    wl->w.h_stage[i] = h_segs[i];
    wl->w.h_stage[i].flags = wl->w.flags[i];      // same lengths => same geometry class as at create()
  }
  if (nsegs > 0)
    RPB_CHECK(cudaMemcpyAsync(wl->w.d_segs, wl->w.h_stage, sizeof(rpb200_halo_seg) * nsegs, cudaMemcpyHostToDevice, rpb_stream(s)));
  return 0;
}

extern "C" void rpb200_halo_worklist_destroy(rpb200_halo_worklist* wl)
{
  if (!wl) return;
  worklist_free(wl->w);
  delete wl;
}

extern "C" int rpb200_halo_pack(rpb200_ctx* ctx, const rpb200_halo_worklist* wl, rpb200_stream_t s)
{
  if (!ctx || !wl) return RPB200_EINVAL;
  return worklist_launch<true, 0>(ctx, RPB_K_HALO_PACKING_FUSED, wl->w, exchange_args(), rpb_stream(s));
}

extern "C" int rpb200_halo_unpack(rpb200_ctx* ctx, const rpb200_halo_worklist* wl, rpb200_stream_t s)
{
  if (!ctx || !wl) return RPB200_EINVAL;
  return worklist_launch<false, 0>(ctx, RPB_K_HALO_PACKING_FUSED, wl->w, exchange_args(), rpb_stream(s));
}

// comm/HALO_base.cpp:31-35: the double is truncated when stored into the Index_type dims
extern "C" void rpb200_halo_grid_dims(int64_t target, int64_t dims[3])
{
  const double c = cbrt((double)target) + cbrt(3.0) - 1;
  dims[0] = dims[1] = dims[2] = (int64_t)c;
}

extern "C" void rpb200_halo_plan_destroy(rpb200_halo_plan* p)
{
  if (!p) return;
  cudaFree(p->d_lists);
  worklist_free(p->pack_wl); worklist_free(p->unpack_wl);
  for (int g = 0; g < 2; ++g) { worklist_free(p->xpack_wl[g]); worklist_free(p->xunpack_wl[g]); worklist_free(p->xsend_wl[g]); }
  for (size_t r = 0; r < p->peer_windows.size(); ++r)
    if (p->peer_opened[r] && p->peer_windows[r]) cudaIpcCloseMemHandle(p->peer_windows[r]);
  cudaFree(p->d_window);
  cudaFree(p->d_pack_msgs); cudaFree(p->d_unpack_msgs); cudaFree(p->d_send_msgs); cudaFree(p->d_msg_done); cudaFree(p->d_error);
  cudaFree(p->d_epoch); cudaFree(p->d_unpack_done);
  cudaFree(p->d_xchg_block); cudaFree(p->d_prog_block);
  delete p;
}

extern "C" int rpb200_halo_plan_create(rpb200_ctx* ctx, const int64_t grid_dims[3], int64_t halo_width,
                                       int num_vars, int my_rank, const int rank_dims[3],
                                       rpb200_halo_plan** out)
{
  if (!ctx || !grid_dims || !rank_dims || !out || halo_width < 1 || num_vars < 1) return RPB200_EINVAL;
  *out = nullptr;
  const int P = rank_dims[0] * rank_dims[1] * rank_dims[2];
  if (rank_dims[0] < 1 || rank_dims[1] < 1 || rank_dims[2] < 1 || my_rank < 0 || my_rank >= P) return RPB200_EINVAL;
  for (int a = 0; a < 3; ++a) if (grid_dims[a] < halo_width) return RPB200_EINVAL;
  rpb200_halo_plan* p = new (std::nothrow) rpb200_halo_plan();
  if (!p) return (int)cudaErrorMemoryAllocation;
  p->ctx = ctx; p->hw = halo_width; p->nvars = num_vars; p->my_rank = my_rank; p->nranks = P;
  for (int a = 0; a < 3; ++a) { p->dims[a] = grid_dims[a]; p->rank_dims[a] = rank_dims[a]; }
  const int64_t ext[3] = {grid_dims[0] + 2 * halo_width, grid_dims[1] + 2 * halo_width, grid_dims[2] + 2 * halo_width};
  p->var_size = ext[0] * ext[1] * ext[2];
  if (p->var_size > 0x7fffffffll) { delete p; return RPB200_EINVAL; }     // Int_type index lists

  // rank coordinates, x fastest (HALO_base.cpp:183-186); periodic wrap (:214-221); tags (:192-195, 227, 260)
  int me[3];
  me[2] = my_rank / (rank_dims[0] * rank_dims[1]);
  me[1] = (my_rank - me[2] * rank_dims[0] * rank_dims[1]) / rank_dims[0];
  me[0] = my_rank - me[2] * rank_dims[0] * rank_dims[1] - me[1] * rank_dims[0];
  auto slot = [](const int o[3]) { return (o[0] + 1) + 3 * (o[1] + 1) + 9 * (o[2] + 1); };
  int slot_to_l[27];
  for (int l = 0; l < NNB; ++l) slot_to_l[slot(k_offsets[l])] = l;
  size_t total_ints = 0;
  size_t list_off[2][NNB];
  for (int l = 0; l < NNB; ++l) {
    int nb[3], opp[3];
    for (int a = 0; a < 3; ++a) {
      nb[a] = me[a] + k_offsets[l][a];
      if (nb[a] >= rank_dims[a]) nb[a] = 0; else if (nb[a] < 0) nb[a] = rank_dims[a] - 1;
      opp[a] = -k_offsets[l][a];
    }
    p->ranks[l] = nb[0] + rank_dims[0] * (nb[1] + rank_dims[1] * nb[2]);
    p->send_tags[l] = l;
    p->opposite[l] = p->recv_tags[l] = slot_to_l[slot(opp)];
    p->pack_len[l] = make_box(false, k_offsets[l], halo_width, grid_dims).len();
    p->unpack_len[l] = make_box(true, k_offsets[l], halo_width, grid_dims).len();
    list_off[0][l] = total_ints; total_ints += ((size_t)p->pack_len[l] + 63) / 64 * 64;
    list_off[1][l] = total_ints; total_ints += ((size_t)p->unpack_len[l] + 63) / 64 * 64;
  }
  cudaError_t e = cudaMalloc(&p->d_lists, sizeof(int) * total_ints);
  if (e != cudaSuccess) { delete p; return (int)e; }
  for (int l = 0; l < NNB; ++l) {
    p->d_pack_list[l] = p->d_lists + list_off[0][l];
    p->d_unpack_list[l] = p->d_lists + list_off[1][l];
    for (int r = 0; r < 2; ++r) {
      const box b = make_box(r == 1, k_offsets[l], halo_width, grid_dims);
      const int64_t len = b.len();
      int* dst = r ? p->d_unpack_list[l] : p->d_pack_list[l];
      int64_t blocks = (len + 255) / 256; if (blocks > 4096) blocks = 4096;
      halo_fill_list_kernel<<<(int)blocks, 256>>>(dst, len, b.lo[0], b.lo[1], b.lo[2], b.hi[0] - b.lo[0],
                                                  b.hi[1] - b.lo[1], ext[0], ext[0] * ext[1]);
    }
  }
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { rpb200_halo_plan_destroy(p); return (int)e; }
  *out = p;
  return 0;
}

extern "C" int64_t rpb200_halo_plan_var_size(const rpb200_halo_plan* p) { return p ? p->var_size : 0; }

extern "C" int rpb200_halo_plan_neighbor(const rpb200_halo_plan* p, int l, int* rank, int* send_tag, int* recv_tag,
                                         int64_t* pack_len, int64_t* unpack_len, const int** d_pack_list,
                                         const int** d_unpack_list)
{
  if (!p || l < 0 || l >= NNB) return RPB200_EINVAL;
  if (rank) *rank = p->ranks[l];
  if (send_tag) *send_tag = p->send_tags[l];
  if (recv_tag) *recv_tag = p->recv_tags[l];
  if (pack_len) *pack_len = p->pack_len[l];
  if (unpack_len) *unpack_len = p->unpack_len[l];
  if (d_pack_list) *d_pack_list = p->d_pack_list[l];
  if (d_unpack_list) *d_unpack_list = p->d_unpack_list[l];
  return 0;
}

// ---- item lists of the one-launch kernels ------------------------------------------------------------------------
// Built from the geometry of two work lists alone.  The X TUPLES are the strided tuples (cells not adjacent in the variable)
// of the greatest strided length -- for HALO_base's lists the -x and +x faces of every variable.  Chunk c of every X tuple
// goes into one GROUP of adjacent items, the tuples of one variable next to each other (see halo_items_kernel); every other
// tuple contributes groups {chunk c of the tuples of one message}.
static int64_t plan_chunks(int64_t len) { return (len + HALO_CHUNK - 1) / HALO_CHUNK; }

static int64_t x_tuple_len(const worklist_dev& a, const worklist_dev& b)
{
  int64_t m = 0;
  for (const worklist_dev* w : {&a, &b})
    for (const rpb200_halo_seg& s : w->h_segs) if ((s.flags & SEG_STRIDED) && s.len > m) m = s.len;
  return m >= HALO_CHUNK ? m : 0;                       // short strided tuples (edges) are not worth a group of their own
}
static bool is_x_tuple(const rpb200_halo_seg& s, int64_t xlen) { return xlen > 0 && (s.flags & SEG_STRIDED) && s.len == xlen; }

// the X tuples of both lists, ordered (variable, side, message): {pack(-x), pack(+x), unpack(-x), unpack(+x)} per variable
struct x_ref { const double* var; int side; int msg; int seg; };
static std::vector<x_ref> x_tuples(const worklist_dev& pw, const worklist_dev& uw, int64_t xlen, bool want_pack, bool want_unpack)
{
  std::vector<x_ref> x;
  if (want_pack) for (int i = 0; i < pw.nsegs; ++i) if (is_x_tuple(pw.h_segs[i], xlen)) x.push_back(x_ref{pw.h_segs[i].var, 0, pw.h_segs[i].msg, item_make(i, 1, 0)});
  if (want_unpack) for (int i = 0; i < uw.nsegs; ++i) if (is_x_tuple(uw.h_segs[i], xlen)) x.push_back(x_ref{uw.h_segs[i].var, 1, uw.h_segs[i].msg, item_make(i, 1, ITEM_UNPACK)});
  // variables in order of first appearance (not of address), then side, then message
  std::vector<const double*> order;
  for (const x_ref& r : x) { bool seen = false; for (const double* v : order) seen = seen || v == r.var; if (!seen) order.push_back(r.var); }
  auto rank = [&](const double* v) { size_t k = 0; while (k < order.size() && order[k] != v) ++k; return k; };
  std::stable_sort(x.begin(), x.end(), [&](const x_ref& a, const x_ref& b) {
    const size_t ra = rank(a.var), rb = rank(b.var);
    if (ra != rb) return ra < rb;
    if (a.side != b.side) return a.side < b.side;
    return a.msg < b.msg;
  });
  return x;
}

// A unit list under construction: items + the first item of every unit
struct unit_list {
  std::vector<halo_item> items;
  std::vector<int> first;                              // first[u]; the end sentinel is appended by finish()
  void add_unit(const std::vector<halo_item>& g) { if (g.empty()) return; first.push_back((int)items.size()); items.insert(items.end(), g.begin(), g.end()); }
  void add_singles(const std::vector<halo_item>& g) { for (const halo_item& h : g) { first.push_back((int)items.size()); items.push_back(h); } }
  bool finished = false;
  int units() const { return (int)first.size() - (finished ? 1 : 0); }
  void finish() { first.push_back((int)items.size()); finished = true; }
};

// streaming groups of one list: for every message (run of consecutive tuples with the same msg), chunk-major.  Consecutive
// tuples of a message that share one index list and length (the variables of a neighbour: plan_segments, and the reference's
// own tuples, HALO_PACKING_FUSED-Seq.cpp:43-61) become ONE item: the kernel loads the indices of a chunk once for all of them.
static void stream_groups(const worklist_dev& w, int64_t xlen, int bit, std::vector<std::vector<halo_item>>& sg)
{
  int i = 0;
  while (i < w.nsegs) {
    int j = i;
    int64_t maxc = 0;
    while (j < w.nsegs && w.h_segs[j].msg == w.h_segs[i].msg) { if (!is_x_tuple(w.h_segs[j], xlen)) maxc = std::max(maxc, plan_chunks(w.h_segs[j].len)); ++j; }
    // runs [t, t + n) of tuples sharing list and length
    std::vector<std::pair<int, int>> runs;
    for (int t = i; t < j;) {
      if (is_x_tuple(w.h_segs[t], xlen)) { ++t; continue; }
      int n = 1;
      while (t + n < j && n < 255 && !is_x_tuple(w.h_segs[t + n], xlen) && w.h_segs[t + n].list == w.h_segs[t].list &&
             w.h_segs[t + n].len == w.h_segs[t].len && w.h_segs[t + n].flags == w.h_segs[t].flags) ++n;
      runs.push_back(std::make_pair(t, n));
      t += n;
    }
    for (int64_t c = 0; c < maxc; ++c) {
      std::vector<halo_item> g;
      for (const auto& r : runs)
        if (c < plan_chunks(w.h_segs[r.first].len)) g.push_back(halo_item{item_make(r.first, r.second, bit), (int)c});
      if (!g.empty()) sg.push_back(g);
    }
    i = j;
  }
}

// the X tuples of one variable: consecutive entries of x_tuples() with the same var
static std::vector<std::vector<x_ref>> by_variable(const std::vector<x_ref>& x)
{
  std::vector<std::vector<x_ref>> out;
  for (const x_ref& r : x) {
    if (out.empty() || out.back().front().var != r.var) out.emplace_back();
    out.back().push_back(r);
  }
  return out;
}

// HALO_PACKING_FUSED in one launch.  order 1: X units {pack(-x), pack(+x), unpack(-x), unpack(+x)} of (chunk, variable) spread
// evenly among the streaming units (Bresenham: burst-bound and bandwidth-bound traffic in flight together); 3: X units first;
// 5: two phases in one launch -- every pack unit (X pairs last), then every unpack unit (X pairs first, descending): the
// order of the exchange, where the phases cannot interleave.
static void build_items_merged(const worklist_dev& pw, const worklist_dev& uw, int order, unit_list& L)
{
  const int64_t xlen = x_tuple_len(pw, uw);
  std::vector<std::vector<halo_item>> sp, su;
  stream_groups(pw, xlen, 0, sp);
  stream_groups(uw, xlen, ITEM_UNPACK, su);
  if (order == 5) {
    const auto xp = by_variable(x_tuples(pw, uw, xlen, true, false)), xu = by_variable(x_tuples(pw, uw, xlen, false, true));
    for (const auto& g : sp) L.add_singles(g);
    for (int64_t c = 0; c < plan_chunks(xlen); ++c)
      for (const auto& var : xp) { std::vector<halo_item> g; for (const x_ref& r : var) g.push_back(halo_item{r.seg, (int)c}); L.add_unit(g); }
    for (int64_t c = plan_chunks(xlen) - 1; c >= 0; --c)
      for (size_t k = xu.size(); k-- > 0;) { std::vector<halo_item> g; for (const x_ref& r : xu[k]) g.push_back(halo_item{r.seg, (int)c}); L.add_unit(g); }
    for (const auto& g : su) L.add_singles(g);
    L.finish();
    return;
  }
  const auto xv = by_variable(x_tuples(pw, uw, xlen, true, true));
  // streaming groups alternate pack / unpack
  std::vector<std::vector<halo_item>> sg;
  for (size_t k = 0; k < sp.size() || k < su.size(); ++k) {
    if (k < sp.size()) sg.push_back(sp[k]);
    if (k < su.size()) sg.push_back(su[k]);
  }
  const size_t nx = xv.empty() ? 0 : (size_t)plan_chunks(xlen), ns = sg.size();
  size_t ix = 0, is = 0;
  while (ix < nx || is < ns) {
    bool take_x;
    if (order == 3) take_x = ix < nx;
    else if (ix >= nx) take_x = false;
    else if (is >= ns) take_x = true;
    else take_x = (ix + 1) * ns * 17 <= (is + 1) * nx * 20;      // X units spread evenly over the first 85 % of the list:
                                                                  // the launch ends on light units (tickets balance the rest)
    if (take_x) {
      for (const auto& var : xv) { std::vector<halo_item> g; for (const x_ref& r : var) g.push_back(halo_item{r.seg, (int)ix}); L.add_unit(g); }
      ++ix;
    } else {
      L.add_singles(sg[is++]);
    }
  }
  L.finish();
}

// HALO_EXCHANGE_FUSED: all pack units, then all unpack units.  The X tuples are packed LAST (ascending chunks) and unpacked
// FIRST (descending chunks): the ghost cells share their L2 lines with the owned cells read a moment earlier (LIFO reuse
// across the signal); the -x / +x chunks of one variable form one unit (one CTA: shared sectors, shared DRAM bursts).
// progressive (the one-ticket form): X pair units FIRST on both sides, ascending -- the strided faces are the slowest messages to
// pack and to unpack, so they start first, and their flags are out after a third of the packing.
static void build_items_xchg(const worklist_dev& pw, const worklist_dev& uw, unit_list& L, int* n_pack_units, bool progressive = false)
{
  const int64_t xlen = x_tuple_len(pw, uw);
  std::vector<std::vector<halo_item>> sp, su;
  stream_groups(pw, xlen, 0, sp);
  stream_groups(uw, xlen, ITEM_UNPACK, su);
  const auto xp = by_variable(x_tuples(pw, uw, xlen, true, false)), xu = by_variable(x_tuples(pw, uw, xlen, false, true));
  if (progressive) {
    for (int64_t c = 0; c < plan_chunks(xlen); ++c)
      for (const auto& var : xp) { std::vector<halo_item> g; for (const x_ref& r : var) g.push_back(halo_item{r.seg, (int)c}); L.add_unit(g); }
    for (const auto& g : sp) L.add_singles(g);
    *n_pack_units = L.units();
    for (int64_t c = 0; c < plan_chunks(xlen); ++c)
      for (const auto& var : xu) { std::vector<halo_item> g; for (const x_ref& r : var) g.push_back(halo_item{r.seg, (int)c}); L.add_unit(g); }
    for (const auto& g : su) L.add_singles(g);
    L.finish();
    return;
  }
  for (const auto& g : sp) L.add_singles(g);
  for (int64_t c = 0; c < plan_chunks(xlen); ++c)
    for (const auto& var : xp) { std::vector<halo_item> g; for (const x_ref& r : var) g.push_back(halo_item{r.seg, (int)c}); L.add_unit(g); }
  *n_pack_units = L.units();
  for (int64_t c = plan_chunks(xlen) - 1; c >= 0; --c)
    for (size_t k = xu.size(); k-- > 0;) { std::vector<halo_item> g; for (const x_ref& r : xu[k]) g.push_back(halo_item{r.seg, (int)c}); L.add_unit(g); }
  for (const auto& g : su) L.add_singles(g);
  L.finish();
}

// items and unit_first in ONE device allocation: [items | unit_first]
// items, unit_first and the unit tickets in ONE device allocation: [items | unit_first | 4 ticket words (zero: re-armed by the kernel)]
static int upload_units(const unit_list& L, void** d_block, const halo_item** d_items, const int** d_unit_first, unsigned int** d_ticket)
{
  cudaFree(*d_block); *d_block = nullptr; *d_items = nullptr; *d_unit_first = nullptr; *d_ticket = nullptr;
  if (L.items.empty()) return 0;
  const size_t ib = sizeof(halo_item) * L.items.size(), ub = (sizeof(int) * L.first.size() + 15) / 16 * 16;
  RPB_CHECK(cudaMalloc(d_block, ib + ub + 16));
  RPB_CHECK(cudaMemset(*d_block, 0, ib + ub + 16));
  RPB_CHECK(cudaMemcpy(*d_block, L.items.data(), ib, cudaMemcpyHostToDevice));
  RPB_CHECK(cudaMemcpy((char*)*d_block + ib, L.first.data(), sizeof(int) * L.first.size(), cudaMemcpyHostToDevice));
  *d_items = (const halo_item*)*d_block;
  *d_unit_first = (const int*)((char*)*d_block + ib);
  *d_ticket = (unsigned int*)((char*)*d_block + ib + ub);
  return 0;
}

template <bool XCHG>
static int launch_items(const rpb200_ctx* ctx, int kid, const halo_items_args& A, cudaStream_t st)
{
  if (A.n_units == 0) return 0;
  int resident = 0;             // XCHG: the grid must be fully co-resident (phase 2 spins on flags other ranks' phase 1 releases)
  RPB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, halo_items_kernel<XCHG>, HALO_BLOCK, 0));
  if (resident < 1) return (int)cudaErrorLaunchOutOfResources;
  int cps = ctx->tune[kid].ctas_per_sm > 0 ? ctx->tune[kid].ctas_per_sm : 4;
  if (cps > resident) cps = resident;
  int64_t grid = (int64_t)ctx->sm_count * cps;
  const int64_t most = XCHG ? (A.n_pack_units > A.n_units - A.n_pack_units ? A.n_pack_units : A.n_units - A.n_pack_units) : A.n_units;
  if (grid > most) grid = most;
  if (grid < 1) grid = 1;
  halo_items_kernel<XCHG><<<(int)grid, HALO_BLOCK, 0, st>>>(A);
  RPB_LAUNCH_CHECK();
  return 0;
}

// the merged item list of a (pack, unpack) pair, cached on the pack list; rebuilt when the pair or the order tuning changes
static int merged_order_of(const rpb200_ctx* ctx)
{
  const int u = ctx->tune[RPB_K_HALO_PACKING_FUSED].unroll;       // tuning `unroll`: 3 = X units first, 5 = two phases, else X units mixed in
  return (u == 3 || u == 5) ? u : 1;
}
static int ensure_merged(const rpb200_ctx* ctx, worklist_dev& pw, worklist_dev& uw)
{
  const int order = merged_order_of(ctx);
  if (pw.merged_with == (const void*)&uw && pw.merged_order == order) return 0;
  unit_list L;
  build_items_merged(pw, uw, order, L);
  const int rc = upload_units(L, &pw.d_items, &pw.d_item_list, &pw.d_unit_first, &pw.d_unit_ticket);
  if (rc != 0) { pw.merged_with = nullptr; return rc; }
  pw.n_units = L.units(); pw.merged_with = &uw; pw.merged_order = order;
  return 0;
}

// The one-launch pack + unpack over two work lists; the item list is built on the first call for a (pack, unpack) pair
// (rpb200_halo_plan_bind builds it eagerly).
static int worklist_pack_unpack(rpb200_ctx* ctx, worklist_dev& pw, worklist_dev& uw, cudaStream_t st)
{
  const int unroll = ctx->tune[RPB_K_HALO_PACKING_FUSED].unroll;
  const bool two_launches = unroll == 2 || unroll == 4 || pw.nsegs > ITEMS_MAX_SEGS || uw.nsegs > ITEMS_MAX_SEGS;
  if (two_launches) {                              // tuning `unroll` 2 / 4: the two-launch form (without / with L2 hints)
    const int rc = worklist_launch<true, 0>(ctx, RPB_K_HALO_PACKING_FUSED, pw, exchange_args(), st);
    return rc != 0 ? rc : worklist_launch<false, 0>(ctx, RPB_K_HALO_PACKING_FUSED, uw, exchange_args(), st);
  }
  { const int rc = ensure_merged(ctx, pw, uw); if (rc != 0) return rc; }
  halo_items_args A;
  memset(&A, 0, sizeof(A));
  A.psegs[0] = A.psegs[1] = pw.d_segs;
  A.usegs[0] = A.usegs[1] = uw.d_segs;
  A.items = pw.d_item_list; A.unit_first = pw.d_unit_first; A.ticket = pw.d_unit_ticket; A.n_units = pw.n_units; A.n_pack_units = 0;
  A.npsegs = pw.nsegs; A.nusegs = uw.nsegs;
  return launch_items<false>(ctx, RPB_K_HALO_PACKING_FUSED, A, st);
}

// Host-only: the unit list the one-launch kernels would walk for tuples of the given geometry (no device memory is touched, so
// the CPU test-suite checks the builders: every (tuple, chunk) exactly once, units well-formed, pack units before unpack units
// for the exchange).  order: 1 / 3 / 5 = the HALO_PACKING_FUSED orders, 0 = the exchange order, 6 = the progressive exchange order.
extern "C" int rpb200_debug_halo_units(const int64_t* pack_len, const int* pack_strided, const int* pack_msg, const int* pack_var, int npack,
                                       const int64_t* unpack_len, const int* unpack_strided, const int* unpack_msg, const int* unpack_var,
                                       int nunpack, int order, int* items_out, int max_items, int* unit_first_out, int max_units,
                                       int* n_items, int* n_units, int* n_pack_units)
{
  if (npack < 0 || nunpack < 0 || !n_items || !n_units) return RPB200_EINVAL;
  worklist_dev pw, uw;
  auto fill = [](worklist_dev& w, const int64_t* len, const int* strided, const int* msg, const int* var, int n) {
    w.nsegs = n;
    for (int i = 0; i < n; ++i) {
      rpb200_halo_seg s;
      memset(&s, 0, sizeof(s));
      s.len = len[i]; s.msg = msg[i]; s.flags = strided[i] ? SEG_STRIDED : 0;
      s.var = reinterpret_cast<double*>((uintptr_t)(var[i] + 1) * 4096);      // identity of the variable only
      w.h_segs.push_back(s);
    }
  };
  fill(pw, pack_len, pack_strided, pack_msg, pack_var, npack);
  fill(uw, unpack_len, unpack_strided, unpack_msg, unpack_var, nunpack);
  unit_list L;
  int npu = 0;
  if (order == 0 || order == 6) build_items_xchg(pw, uw, L, &npu, order == 6);
  else build_items_merged(pw, uw, order, L);
  *n_items = (int)L.items.size(); *n_units = L.units();
  if (n_pack_units) *n_pack_units = npu;
  if ((int)L.items.size() > max_items || (int)L.first.size() > max_units + 1) return RPB200_EINVAL;
  for (size_t i = 0; i < L.items.size(); ++i) { items_out[2 * i] = L.items[i].seg; items_out[2 * i + 1] = L.items[i].chunk; }
  for (size_t i = 0; i < L.first.size(); ++i) unit_first_out[i] = L.first[i];
  return 0;
}

extern "C" int rpb200_halo_pack_unpack(rpb200_ctx* ctx, rpb200_halo_worklist* pack, rpb200_halo_worklist* unpack, rpb200_stream_t s)
{
  if (!ctx || !pack || !unpack || pack == unpack) return RPB200_EINVAL;
  RPB_CHECK_DEVICE(ctx);
  return worklist_pack_unpack(ctx, pack->w, unpack->w, rpb_stream(s));
}

// neighbour-major, variable-minor segments (HALO_PACKING_FUSED-Seq.cpp:43-61, 71-97)
static int plan_segments(const rpb200_halo_plan* p, bool pack, double* const* vars, double* const* buffers,
                         std::vector<rpb200_halo_seg>& segs)
{
  segs.clear();
  for (int l = 0; l < NNB; ++l) {
    const int64_t len = pack ? p->pack_len[l] : p->unpack_len[l];
    if (!buffers[l]) return RPB200_EINVAL;
    for (int v = 0; v < p->nvars; ++v) {
      if (!vars[v]) return RPB200_EINVAL;
      rpb200_halo_seg s;
      s.buffer = buffers[l] + (int64_t)v * len;
      s.list = pack ? p->d_pack_list[l] : p->d_unpack_list[l];
      s.var = vars[v];
      s.len = len;
      s.msg = l;
      s.flags = 0;
      segs.push_back(s);
    }
  }
  return 0;
}

extern "C" int rpb200_halo_plan_bind(rpb200_halo_plan* p, double* const* vars, double* const* pack_buffers,
                                     double* const* unpack_buffers)
{
  if (!p || !vars || !pack_buffers || !unpack_buffers) return RPB200_EINVAL;
  worklist_free(p->pack_wl); worklist_free(p->unpack_wl);
  p->bound = false;
  std::vector<rpb200_halo_seg> segs;
  int rc = plan_segments(p, true, vars, pack_buffers, segs);
  if (rc == 0) rc = worklist_build(p->pack_wl, segs.data(), (int)segs.size());
  if (rc == 0) rc = plan_segments(p, false, vars, unpack_buffers, segs);
  if (rc == 0) rc = worklist_build(p->unpack_wl, segs.data(), (int)segs.size());
  if (rc == 0 && p->pack_wl.nsegs <= ITEMS_MAX_SEGS && p->unpack_wl.nsegs <= ITEMS_MAX_SEGS)
    rc = ensure_merged(p->ctx, p->pack_wl, p->unpack_wl);      // eager: rpb200_halo_plan_pack_unpack never allocates
  if (rc != 0) { worklist_free(p->pack_wl); worklist_free(p->unpack_wl); return rc; }
  p->bound = true;
  return 0;
}

// HALO_PACKING_FUSED, one rep in ONE launch: the pack of HALO_PACKING_FUSED-Seq.cpp:43-61 and the unpack of :71-97 touch
// disjoint cells and buffers, so their items may interleave (halo_items_kernel)
extern "C" int rpb200_halo_plan_pack_unpack(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->bound) return RPB200_EINVAL;
  RPB_CHECK_DEVICE(p->ctx);
  return worklist_pack_unpack(p->ctx, p->pack_wl, p->unpack_wl, rpb_stream(s));
}

extern "C" int rpb200_halo_plan_pack(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->bound) return RPB200_EINVAL;
  return worklist_launch<true, 0>(p->ctx, RPB_K_HALO_PACKING_FUSED, p->pack_wl, exchange_args(), rpb_stream(s));
}

extern "C" int rpb200_halo_plan_unpack(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->bound) return RPB200_EINVAL;
  return worklist_launch<false, 0>(p->ctx, RPB_K_HALO_PACKING_FUSED, p->unpack_wl, exchange_args(), rpb_stream(s));
}

// ---- exchange ---------------------------------------------------------------------------------
// window layout (identical on every rank because every rank has the same grid):
//   [0, 256)                  26 arrival flags (uint64), one per receive slot
//   generation 0, generation 1: receive slot l at recv_off[g][l], nvars * unpack_len[l] doubles
extern "C" int rpb200_halo_exchange_window(rpb200_halo_plan* p, double* const* vars, void** d_window,
                                           size_t* bytes, unsigned char ipc_handle[64])
{
  if (!p) return RPB200_EINVAL;          // vars == NULL: a transport-only window (HALO_SENDRECV), no pack/unpack work lists
  if (!p->d_window) {
    size_t off = 256;
    for (int g = 0; g < 2; ++g)
      for (int l = 0; l < NNB; ++l) {
        p->recv_off[g][l] = off;
        off += ((size_t)p->nvars * (size_t)p->unpack_len[l] * sizeof(double) + 255) / 256 * 256;
      }
    p->window_bytes = off;
    RPB_CHECK(cudaMalloc(&p->d_window, off));
    RPB_CHECK(cudaMemset(p->d_window, 0, off));
    RPB_CHECK(cudaMalloc(&p->d_pack_msgs, sizeof(halo_msg) * NNB));
    RPB_CHECK(cudaMalloc(&p->d_unpack_msgs, sizeof(halo_msg) * NNB));
    RPB_CHECK(cudaMalloc(&p->d_msg_done, sizeof(unsigned int) * NNB));
    RPB_CHECK(cudaMemset(p->d_msg_done, 0, sizeof(unsigned int) * NNB));
    RPB_CHECK(cudaMalloc(&p->d_error, sizeof(int)));
    RPB_CHECK(cudaMemset(p->d_error, 0, sizeof(int)));
    RPB_CHECK(cudaMalloc(&p->d_epoch, sizeof(unsigned long long)));
    RPB_CHECK(cudaMemset(p->d_epoch, 0, sizeof(unsigned long long)));
    RPB_CHECK(cudaMalloc(&p->d_unpack_done, sizeof(unsigned int)));
    RPB_CHECK(cudaMemset(p->d_unpack_done, 0, sizeof(unsigned int)));
    RPB_CHECK(cudaDeviceSynchronize());
  }
  if (vars) p->vars.assign(vars, vars + p->nvars); else p->vars.clear();
  if (d_window) *d_window = p->d_window;
  if (bytes) *bytes = p->window_bytes;
  if (ipc_handle) {
    cudaIpcMemHandle_t h;
    RPB_CHECK(cudaIpcGetMemHandle(&h, p->d_window));
    memcpy(ipc_handle, &h, sizeof(h));
  }
  return 0;
}

static int exchange_finish_connect(rpb200_halo_plan* p)
{
  // pack work lists: message l of generation g goes into window[ranks[l]] slot opposite[l]
  std::vector<rpb200_halo_seg> segs;
  std::vector<halo_msg> pm(NNB), um(NNB);
  for (int g = 0; g < 2; ++g) {
    double* dst[NNB]; double* src[NNB];
    for (int l = 0; l < NNB; ++l) {
      const int o = p->opposite[l];
      if (p->pack_len[l] != p->unpack_len[o]) return RPB200_EINVAL;
      dst[l] = (double*)((unsigned char*)p->peer_windows[p->ranks[l]] + p->recv_off[g][o]);
      src[l] = (double*)(p->d_window + p->recv_off[g][l]);
    }
    worklist_free(p->xpack_wl[g]); worklist_free(p->xunpack_wl[g]);
    if (p->vars.empty()) continue;               // transport-only window
    int rc = plan_segments(p, true, p->vars.data(), dst, segs);
    if (rc == 0) rc = worklist_build(p->xpack_wl[g], segs.data(), (int)segs.size());
    if (rc == 0) rc = plan_segments(p, false, p->vars.data(), src, segs);
    if (rc == 0) rc = worklist_build(p->xunpack_wl[g], segs.data(), (int)segs.size());
    if (rc != 0) return rc;
  }
  for (int l = 0; l < NNB; ++l) {
    const int o = p->opposite[l];
    pm[l].remote_flag = (unsigned long long*)p->peer_windows[p->ranks[l]] + o;
    pm[l].my_flag = nullptr;
    pm[l].chunks = (unsigned int)(p->nvars * ((p->pack_len[l] + HALO_CHUNK - 1) / HALO_CHUNK));
    pm[l].pad = 0;
    um[l].remote_flag = nullptr;
    um[l].my_flag = (const unsigned long long*)p->d_window + l;
    um[l].chunks = 0; um[l].pad = 0;
  }
  RPB_CHECK(cudaMemcpy(p->d_pack_msgs, pm.data(), sizeof(halo_msg) * NNB, cudaMemcpyHostToDevice));
  RPB_CHECK(cudaMemcpy(p->d_unpack_msgs, um.data(), sizeof(halo_msg) * NNB, cudaMemcpyHostToDevice));
  p->n_units_xchg = 0; p->n_units_prog = 0;
  if (!p->vars.empty() && NNB * p->nvars <= ITEMS_MAX_SEGS) {
    unit_list L;
    int n_pack = 0;
    build_items_xchg(p->xpack_wl[0], p->xunpack_wl[0], L, &n_pack);
    const int rc = upload_units(L, &p->d_xchg_block, &p->d_items_xchg, &p->d_unit_first_xchg, &p->d_ticket_xchg);
    if (rc != 0) return rc;
    p->n_units_xchg = L.units(); p->n_pack_units_xchg = n_pack;
    unit_list P;
    build_items_xchg(p->xpack_wl[0], p->xunpack_wl[0], P, &n_pack, true);
    const int rc2 = upload_units(P, &p->d_prog_block, &p->d_items_prog, &p->d_unit_first_prog, &p->d_ticket_prog);
    if (rc2 != 0) return rc2;
    p->n_units_prog = P.units(); p->n_pack_units_prog = n_pack;
  }
  p->connected = true;
  return 0;
}

extern "C" int rpb200_halo_exchange_connect(rpb200_halo_plan* p, int nranks, const unsigned char* handles)
{
  if (!p || !p->d_window || nranks != p->nranks || !handles) return RPB200_EINVAL;
  p->peer_windows.assign(nranks, nullptr);
  p->peer_opened.assign(nranks, false);
  bool needed[4096] = {false};
  if (nranks > 4096) return RPB200_EINVAL;
  for (int l = 0; l < NNB; ++l) needed[p->ranks[l]] = true;
  for (int r = 0; r < nranks; ++r) {
    if (r == p->my_rank) { p->peer_windows[r] = p->d_window; continue; }
    if (!needed[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, sizeof(h));
    void* ptr = nullptr;
    RPB_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer_windows[r] = ptr; p->peer_opened[r] = true;
  }
  return exchange_finish_connect(p);
}

extern "C" int rpb200_halo_exchange_connect_ptrs(rpb200_halo_plan* p, int nranks, void* const* d_windows)
{
  if (!p || !p->d_window || nranks != p->nranks || !d_windows) return RPB200_EINVAL;
  p->peer_windows.assign(d_windows, d_windows + nranks);
  p->peer_opened.assign(nranks, false);
  p->peer_windows[p->my_rank] = p->d_window;
  for (int l = 0; l < NNB; ++l) if (!p->peer_windows[p->ranks[l]]) return RPB200_EINVAL;
  return exchange_finish_connect(p);
}

static exchange_args plan_xargs(rpb200_halo_plan* p, bool pack)
{
  exchange_args x;
  x.other_gen = pack ? &p->xpack_wl[1] : &p->xunpack_wl[1];
  x.msgs = pack ? p->d_pack_msgs : p->d_unpack_msgs;
  x.msg_done = p->d_msg_done;
  x.d_epoch = p->d_epoch;
  x.unpack_done = p->d_unpack_done;
  x.error = p->d_error;
  return x;
}

extern "C" int rpb200_halo_exchange_pack(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->connected || p->vars.empty()) return RPB200_EINVAL;
  return worklist_launch<true, 1>(p->ctx, RPB_K_HALO_EXCHANGE_FUSED, p->xpack_wl[0], plan_xargs(p, true), rpb_stream(s));
}

extern "C" int rpb200_halo_exchange_unpack(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->connected || p->vars.empty()) return RPB200_EINVAL;
  return worklist_launch<false, 2>(p->ctx, RPB_K_HALO_EXCHANGE_FUSED, p->xunpack_wl[0], plan_xargs(p, false), rpb_stream(s));
}

// The unfused HALO_EXCHANGE (comm/HALO_EXCHANGE-Cuda.cpp:26-123): one launch per (neighbour, variable) tuple.
// The message flag is released by whichever launch completes the message's last chunk (the credit counters
// persist across launches); `commit` marks the last unpack launch of the rep, which advances the epoch.
extern "C" int rpb200_halo_exchange_pack_seg(rpb200_halo_plan* p, int l, int v, rpb200_stream_t s)
{
  if (!p || !p->connected || p->vars.empty() || l < 0 || l >= NNB || v < 0 || v >= p->nvars) return RPB200_EINVAL;
  const int seg = l * p->nvars + v;
  return worklist_launch<true, 1>(p->ctx, RPB_K_HALO_EXCHANGE_FUSED, p->xpack_wl[0], plan_xargs(p, true), rpb_stream(s), seg, seg + 1, 0);
}

extern "C" int rpb200_halo_exchange_unpack_seg(rpb200_halo_plan* p, int l, int v, int commit, rpb200_stream_t s)
{
  if (!p || !p->connected || p->vars.empty() || l < 0 || l >= NNB || v < 0 || v >= p->nvars) return RPB200_EINVAL;
  const int seg = l * p->nvars + v;
  return worklist_launch<false, 2>(p->ctx, RPB_K_HALO_EXCHANGE_FUSED, p->xunpack_wl[0], plan_xargs(p, false), rpb_stream(s), seg, seg + 1,
                                   commit ? 1 : 0);
}

extern "C" int rpb200_halo_exchange(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->connected || p->vars.empty()) return RPB200_EINVAL;
  const rpb_tuning& t = p->ctx->tune[RPB_K_HALO_EXCHANGE_FUSED];
  if (t.unroll == 2 || t.unroll == 4) {   // tuning field `unroll`: 2 / 4 = the two-launch form (pack + signal, then wait + unpack) without / with L2 hints
    const int rc = rpb200_halo_exchange_pack(p, s);
    return rc != 0 ? rc : rpb200_halo_exchange_unpack(p, s);
  }
  if (!p->d_items_xchg || p->n_units_xchg == 0) {       // more tuples than the item kernel keeps in shared memory: two launches
    const int rc = rpb200_halo_exchange_pack(p, s);
    return rc != 0 ? rc : rpb200_halo_exchange_unpack(p, s);
  }
  halo_items_args A;
  memset(&A, 0, sizeof(A));
  A.psegs[0] = p->xpack_wl[0].d_segs; A.psegs[1] = p->xpack_wl[1].d_segs;
  A.usegs[0] = p->xunpack_wl[0].d_segs; A.usegs[1] = p->xunpack_wl[1].d_segs;
  if (t.unroll == 3 && p->n_units_prog > 0) {       // tuning `unroll` 3: the progressive one-ticket form
    A.items = p->d_items_prog; A.unit_first = p->d_unit_first_prog; A.ticket = p->d_ticket_prog;
    A.n_units = p->n_units_prog; A.n_pack_units = p->n_pack_units_prog; A.progressive = 1;
  } else {
    A.items = p->d_items_xchg; A.unit_first = p->d_unit_first_xchg; A.ticket = p->d_ticket_xchg;
    A.n_units = p->n_units_xchg; A.n_pack_units = p->n_pack_units_xchg;
  }
  A.npsegs = p->xpack_wl[0].nsegs; A.nusegs = p->xunpack_wl[0].nsegs;
  A.pmsgs = p->d_pack_msgs; A.umsgs = p->d_unpack_msgs; A.msg_done = p->d_msg_done; A.d_epoch = p->d_epoch;
  A.unpack_done = p->d_unpack_done; A.error = p->d_error; A.timeout_ns = halo_timeout_ns();
  return launch_items<true>(p->ctx, RPB_K_HALO_EXCHANGE_FUSED, A, rpb_stream(s));
}

// ---- HALO_SENDRECV (comm/HALO_SENDRECV-Seq.cpp:34-52): transport only -----------------------------------
extern "C" int rpb200_halo_sendrecv_bind(rpb200_halo_plan* p, double* const* send_buffers)
{
  if (!p || !p->connected || !send_buffers) return RPB200_EINVAL;
  p->send_bound = false;
  for (int g = 0; g < 2; ++g) {
    worklist_free(p->xsend_wl[g]);
    std::vector<rpb200_halo_seg> segs(NNB);
    for (int l = 0; l < NNB; ++l) {
      if (!send_buffers[l]) return RPB200_EINVAL;
      rpb200_halo_seg& s = segs[l];
      s.buffer = (double*)((unsigned char*)p->peer_windows[p->ranks[l]] + p->recv_off[g][p->opposite[l]]);
      s.list = nullptr;
      s.var = send_buffers[l];
      s.len = (int64_t)p->nvars * p->pack_len[l];
      s.msg = l;
      s.flags = SEG_IDENTITY;
    }
    const int rc = worklist_build(p->xsend_wl[g], segs.data(), NNB);
    if (rc != 0) return rc;
  }
  // the credit counters of a message count chunks of the whole message: one tuple per message here
  std::vector<halo_msg> pm(NNB);
  RPB_CHECK(cudaMemcpy(pm.data(), p->d_pack_msgs, sizeof(halo_msg) * NNB, cudaMemcpyDeviceToHost));
  if (!p->d_send_msgs) RPB_CHECK(cudaMalloc(&p->d_send_msgs, sizeof(halo_msg) * NNB));
  for (int l = 0; l < NNB; ++l)
    pm[l].chunks = (unsigned int)(((int64_t)p->nvars * p->pack_len[l] + HALO_CHUNK - 1) / HALO_CHUNK);
  RPB_CHECK(cudaMemcpy(p->d_send_msgs, pm.data(), sizeof(halo_msg) * NNB, cudaMemcpyHostToDevice));
  p->send_bound = true;
  return 0;
}

extern "C" int rpb200_halo_sendrecv_put(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->connected || !p->send_bound) return RPB200_EINVAL;
  exchange_args x = plan_xargs(p, true);
  x.other_gen = &p->xsend_wl[1];
  x.msgs = p->d_send_msgs;
  return worklist_launch<true, 1>(p->ctx, RPB_K_HALO_EXCHANGE_FUSED, p->xsend_wl[0], x, rpb_stream(s));
}

extern "C" int rpb200_halo_sendrecv_wait(rpb200_halo_plan* p, rpb200_stream_t s)
{
  if (!p || !p->connected || !p->send_bound) return RPB200_EINVAL;
  halo_wait_kernel<<<1, 32, 0, rpb_stream(s)>>>(p->d_unpack_msgs, p->d_epoch, p->d_error, halo_timeout_ns());
  RPB_LAUNCH_CHECK();
  return 0;
}

extern "C" int rpb200_halo_sendrecv(rpb200_halo_plan* p, rpb200_stream_t s)
{
  const int rc = rpb200_halo_sendrecv_put(p, s);
  return rc != 0 ? rc : rpb200_halo_sendrecv_wait(p, s);
}

extern "C" int rpb200_halo_recv_buffer(rpb200_halo_plan* p, int l, const double** d_ptr, int64_t* len)
{
  if (!p || !p->d_window || l < 0 || l >= NNB || !d_ptr) return RPB200_EINVAL;
  unsigned long long epoch = 0;                       // committed reps: the last one wrote generation epoch & 1
  RPB_CHECK(cudaMemcpy(&epoch, p->d_epoch, sizeof(epoch), cudaMemcpyDeviceToHost));
  *d_ptr = (const double*)(p->d_window + p->recv_off[epoch & 1ull][l]);
  if (len) *len = (int64_t)p->nvars * p->unpack_len[l];
  return 0;
}

extern "C" int rpb200_halo_exchange_status(rpb200_halo_plan* p)
{
  if (!p || !p->d_error) return RPB200_EINVAL;
  int e = 0;
  RPB_CHECK(cudaMemcpy(&e, p->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  return e;
}
