// pa.cu -- Apps_MASS3DPA, Apps_DIFFUSION3DPA, Apps_CONVECTION3DPA (partial-assembly FEM operators).
//
// Replaces apps/{MASS3DPA,DIFFUSION3DPA,CONVECTION3DPA}-Cuda.cpp: one 25- or 64-thread CTA per
// element, every 1-D contraction a separate shared-memory round trip (8-9 __syncthreads per
// element), basis matrices re-read from shared/global memory for every multiply.
//
// Design here (all three kernels): a CTA works on a BATCH of E elements in three stages, with the
// 1-D contractions grouped so that two of them happen entirely in registers:
//   stage A  task = (element, dz):  the dz-slab of X (D1D x D1D values) is contracted along x and
//            then along y in registers -> a Q1D x Q1D slab per field, written to shared memory;
//   stage B  task = (element, qy,qx pencil): contract along z, apply the quadrature-point operator
//            D (loaded straight from global memory: a pencil index is the fastest-varying index of
//            D, so a warp reads whole contiguous lines), contract back along z;
//   stage C  task = (element, dz):  contract back along y then x in registers, accumulate into Y.
// Only 2 shared-memory exchanges and 3 barriers per batch of E elements; the slab rows in shared
// memory have an odd stride, so every access pattern is bank-conflict free.  The basis matrices
// live in __constant__ memory (refreshed per call by a stream-ordered device-to-device copy), so
// each FMA takes its basis operand from the constant bank instead of a shared-memory load.
// The kernels are HBM-bound by design: algorithmic traffic only (X, D read once, Y read+written).
// The __constant__ tables are module-level state, one set per DEVICE: calls on one stream are ordered by the stream, and a
// call that arrives on another stream than the device's previous PA call is ordered behind it with an event (pa_order
// below), so PA calls on different streams of one device are safe -- they serialise instead of racing on the tables.
//
// Memory pipeline: all three kernels are PERSISTENT and stage their element tensors with bulk-async
// copies (cp.async.bulk + mbarrier, the TMA engine: no registers, no LSU queue) into a shared-memory
// ring, one or two batches ahead of the math.  DIFFUSION / CONVECTION ring D, X and Y; MASS rings D
// (its X/Y slabs are thread-contiguous 128-byte rows, which only bank-conflict in shared memory, so
// they are prefetched into registers one batch ahead instead).  Measured on B200 at NE = 4 M: MASS
// 4.43 -> 5.69 TB/s, DIFFUSION 4.86 -> 6.50 TB/s, CONVECTION 4.81 -> 5.61 TB/s.
//
// Floating point: contractions are regrouped (z first on the way back) and use FMA.  With the
// suite's integer-valued data every partial sum is exact, so results are bit-identical to Base_Seq;
// for general data the difference is rounding-level (see tests/test_apps_gpu.py).
#include "common.cuh"

namespace {

// effective dense basis tables, refreshed per call
__constant__ double c_mass_B[20];    // [q][d]  = B[q + 5 d]     (MASS3DPA.hpp: Bsmem[q][d])
__constant__ double c_mass_Bt[20];   // [d][q]  = Bt[q + 4 d]    (MASS3DPA.hpp: Btsmem[d][q])
__constant__ double c_conv[36];      // B[q][d] (12) | G[q][d] (12) | Bt[d][q] (12)
__constant__ double c_diff[48];      // B[q][d] | G*sign [q][d] | Bt[d][q] | Gt*sign [d][q]

// ------------------------------------------------------------------------------------------------
// MASS3DPA: D1D = 4, Q1D = 5
// ------------------------------------------------------------------------------------------------
// Persistent kernel.  D (1000 of the 2536 bytes per element) is staged by bulk-async copies
// (cp.async.bulk + mbarrier: the TMA path, no registers, no LSU queue) into shared memory; the X slab of the NEXT
// batch is prefetched into registers while the current batch is in stages B and C, and the Y slab is requested
// before stage B.  The kernel is latency-bound on resident warps, so the default keeps ONE D stage per (one-warp)
// CTA -- requested at the top of its own batch, behind stage A -- and spends the shared memory on 12 CTAs per SM
// instead of a second stage (6261 vs 5760 GB/s).
// YRING: the Y slab of the next batch is staged too -- every stage-A/C thread bulk-copies its own 128-byte (e, dz) row
// into a 144-byte-pitched shared-memory row (conflict-free 128-bit reads) -- instead of being loaded through the LSU.
// DS = stages of the D ring: 2 = one batch ahead; 1 = requested at the top of the batch it belongs to (less shared memory
// per CTA, hence more CTAs per SM to hide the exposed latency).
// LM (line-major global accesses; needs E*4 == BLOCK): thread t's (e, dz) slab of X / Y is the 128-byte line t of the batch.
// Loading it with four 256-bit loads per thread makes every warp instruction touch 32 different lines, one 32-byte sector
// each.  With LM, instruction j moves the 32 consecutive 32-byte pieces j*1024 + 32t of the batch instead (8 whole lines), and
// a 144-byte-pitched shared-memory tile (aliasing T while T is dead) turns pieces into slabs (X, before stage A) and slabs
// back into pieces (the result, after stage C; the old Y stays in piece form and is added there).  Same arithmetic, same order.
template <int E, int BLOCK, int MINB, bool YRING, int DS = 2, bool LM = false>
__global__ void __launch_bounds__(BLOCK, MINB)
mass3dpa_kernel(const double* __restrict__ D, const double* __restrict__ X, double* __restrict__ Y,
                int64_t NE)
{
  constexpr int ND = 4, NQ = 5, SLAB = NQ * NQ;          // 25 values per (element, dz) slab
  constexpr int BT = (E * SLAB + BLOCK - 1) / BLOCK;     // stage-B tasks per thread
  constexpr unsigned int DBYTES = E * 125 * sizeof(double);
  static_assert(E * ND <= BLOCK, "one stage-A/C task per thread");
  static_assert(DBYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  extern __shared__ __align__(128) unsigned char pa_smem[];
  static_assert(DS == 2 || !YRING, "the staged-Y variant uses the two-stage ring");
  static_assert(!LM || (E * ND == BLOCK && !YRING && E * ND * 18 <= E * ND * SLAB), "line-major mode: one line per thread, tile inside T");
  double* Ds = reinterpret_cast<double*>(pa_smem);                         // [DS][E*125]
  double* T = Ds + DS * E * 125;                                            // [E*4][25], stride 25 (odd)
  constexpr int YPITCH = 18;                                               // doubles: 16 + 2 pad = 144 bytes
  double* Yr = T + E * ND * SLAB;                                          // [2][E*4][YPITCH]   (YRING only)
  unsigned long long* full = reinterpret_cast<unsigned long long*>(Yr + (YRING ? 2 * E * ND * YPITCH : 0));   // [2]

  const int t = threadIdx.x;
  const int64_t nbatch = (NE + E - 1) / E;
  if (t == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // a ragged last batch is copied by the threads themselves (its byte count need not be a multiple of 16)
  auto issue_D = [&](int64_t batch, int stage) {          // thread 0 only
    const int64_t e0 = batch * E;
    if (NE - e0 >= E) {
      mbar_arrive_expect_tx(&full[stage], DBYTES + (YRING ? E * ND * 128u : 0u));
      bulk_g2s(Ds + stage * E * 125, D + e0 * 125, DBYTES, &full[stage]);
    } else {
      mbar_arrive(&full[stage]);
    }
  };
  // every stage-A/C thread: its own Y row of a FULL batch (after thread 0 announced the bytes)
  auto issue_Y = [&](int64_t batch, int stage) {
    const int64_t e0 = batch * E;
    if (NE - e0 >= E && t < E * ND) bulk_g2s(Yr + (stage * E * ND + t) * YPITCH, Y + e0 * 64 + t * 16, 128u, &full[stage]);
  };
  auto load_X = [&](int64_t batch, dbl4 (&xv)[ND]) {
    const int64_t e0 = batch * E;
    const int cnt = LM ? E : (int)((NE - e0) < E ? (NE - e0) : E);
    if constexpr (LM) {                                     // pieces: instruction j = the batch's bytes [1024j, 1024j + 1024)
#pragma unroll
      for (int j = 0; j < ND; ++j)
        if (j * (BLOCK / 4) + (t >> 2) < cnt * ND) xv[j] = ldg256_stream(X + e0 * 64 + j * (BLOCK * 4) + t * 4);
    } else if (t < cnt * ND) {
      const double* xp = X + (e0 + (t >> 2)) * 64 + (t & 3) * 16;
#pragma unroll
      for (int dy = 0; dy < ND; ++dy) xv[dy] = ldg256_stream(xp + 4 * dy);
    }
  };

  int64_t batch = blockIdx.x;
  dbl4 xv[ND], xn[ND];
  if (batch < nbatch) {
    if (t == 0) issue_D(batch, 0);
    if (YRING) { __syncthreads(); issue_Y(batch, 0); }
    load_X(batch, xv);
  }
  for (int it = 0; batch < nbatch; batch += gridDim.x, ++it) {
    const int stage = DS == 2 ? (it & 1) : 0;
    const int64_t e0 = batch * E;
    const int cnt = LM ? E : (int)((NE - e0) < E ? (NE - e0) : E);     // LM is launched for whole batches only (NE % E == 0)
    const bool has_slab = t < cnt * ND;
    const int64_t next = batch + gridDim.x;
    // the other stage was last read in stage B of iteration it-1, which ended with a barrier
    if (DS == 2) { if (t == 0 && next < nbatch) issue_D(next, stage ^ 1); }
    else if (t == 0 && it > 0) issue_D(batch, 0);             // single stage: this batch's D, now that stage B of it-1 is over
    if (YRING && next < nbatch) { __syncthreads(); issue_Y(next, stage ^ 1); }

    // ---- stage A: (e, dz) -> contract x, then y
    if constexpr (LM) {                                      // pieces -> slabs through the tile (T is dead here)
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        if (j * (BLOCK / 4) + (t >> 2) >= cnt * ND) continue;  // lines beyond a ragged last batch were not loaded
        double* d = T + (j * (BLOCK / 4) + (t >> 2)) * YPITCH + (t & 3) * 4;
        *reinterpret_cast<double2*>(d) = make_double2(xv[j].x, xv[j].y);
        *reinterpret_cast<double2*>(d + 2) = make_double2(xv[j].z, xv[j].w);
      }
      __syncthreads();
      if (has_slab)
#pragma unroll
      for (int dy = 0; dy < ND; ++dy) {
        const double2 lo = *reinterpret_cast<const double2*>(T + t * YPITCH + 4 * dy), hi = *reinterpret_cast<const double2*>(T + t * YPITCH + 4 * dy + 2);
        xv[dy].x = lo.x; xv[dy].y = lo.y; xv[dy].z = hi.x; xv[dy].w = hi.y;
      }
      __syncthreads();                                       // the tile is T: every slab is out before stage A writes T
    }
    if (has_slab) {
      double x[ND][ND];
#pragma unroll
      for (int dy = 0; dy < ND; ++dy) { x[dy][0] = xv[dy].x; x[dy][1] = xv[dy].y; x[dy][2] = xv[dy].z; x[dy][3] = xv[dy].w; }
      double a[ND][NQ];
#pragma unroll
      for (int dy = 0; dy < ND; ++dy)
#pragma unroll
        for (int qx = 0; qx < NQ; ++qx) {
          double s = 0.0;
#pragma unroll
          for (int dx = 0; dx < ND; ++dx) s = fma(x[dy][dx], c_mass_B[qx * ND + dx], s);
          a[dy][qx] = s;
        }
      double* tp = T + t * SLAB;
#pragma unroll
      for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
        for (int qx = 0; qx < NQ; ++qx) {
          double s = 0.0;
#pragma unroll
          for (int dy = 0; dy < ND; ++dy) s = fma(a[dy][qx], c_mass_B[qy * ND + dy], s);
          tp[qy * NQ + qx] = s;
        }
    }
    // requests that land during stages B and C: this batch's Y slab, the next batch's X slab
    double* yp = Y + (e0 + (t >> 2)) * 64 + (t & 3) * 16;
    dbl4 yo[ND];
    if constexpr (LM) {                                      // the old Y in piece form
#pragma unroll
      for (int j = 0; j < ND; ++j)
        if (j * (BLOCK / 4) + (t >> 2) < cnt * ND) yo[j] = ldg256(Y + e0 * 64 + j * (BLOCK * 4) + t * 4);
    } else if (has_slab && !(YRING && cnt == E)) {
#pragma unroll
      for (int dy = 0; dy < ND; ++dy) yo[dy] = ldg256(yp + 4 * dy);
    }
    if (next < nbatch) load_X(next, xn);
    __syncthreads();

    // ---- stage B: (e, pencil) -> contract z, scale by D (from the ring), contract z back
    mbar_wait(&full[stage], DS == 2 ? ((it >> 1) & 1) : (it & 1));
    const double* dsm = Ds + stage * E * 125;
#pragma unroll
    for (int bt = 0; bt < BT; ++bt) {
      const int p = t + bt * BLOCK;
      if (p < cnt * SLAB) {
        const int e = p / SLAB, pen = p - e * SLAB;
        double dq[NQ];
        if (cnt == E) {
#pragma unroll
          for (int qz = 0; qz < NQ; ++qz) dq[qz] = dsm[e * 125 + qz * SLAB + pen];
        } else {
#pragma unroll
          for (int qz = 0; qz < NQ; ++qz) dq[qz] = __ldg(D + (e0 + e) * 125 + qz * SLAB + pen);
        }
        double* tp = T + e * ND * SLAB + pen;
        double u[ND];
#pragma unroll
        for (int dz = 0; dz < ND; ++dz) u[dz] = tp[dz * SLAB];
        double q[NQ];
#pragma unroll
        for (int qz = 0; qz < NQ; ++qz) {
          double s = 0.0;
#pragma unroll
          for (int dz = 0; dz < ND; ++dz) s = fma(u[dz], c_mass_B[qz * ND + dz], s);
          q[qz] = s * dq[qz];
        }
#pragma unroll
        for (int dz = 0; dz < ND; ++dz) {
          double s = 0.0;
#pragma unroll
          for (int qz = 0; qz < NQ; ++qz) s = fma(q[qz], c_mass_Bt[dz * NQ + qz], s);
          tp[dz * SLAB] = s;
        }
      }
    }
    __syncthreads();

    // ---- stage C: (e, dz) -> contract y, then x, accumulate into Y
    if (YRING && cnt == E && has_slab) {        // the staged Y row (the mbarrier wait of stage B covered it)
      const double* yr = Yr + (stage * E * ND + t) * YPITCH;
#pragma unroll
      for (int dy = 0; dy < ND; ++dy) {
        const double2 lo = *reinterpret_cast<const double2*>(yr + 4 * dy), hi = *reinterpret_cast<const double2*>(yr + 4 * dy + 2);
        yo[dy].x = lo.x; yo[dy].y = lo.y; yo[dy].z = hi.x; yo[dy].w = hi.y;
      }
    }
    double a[ND][NQ];
    if (has_slab) {
      const double* tp = T + t * SLAB;
#pragma unroll
      for (int dy = 0; dy < ND; ++dy)
#pragma unroll
        for (int qx = 0; qx < NQ; ++qx) a[dy][qx] = 0.0;
#pragma unroll
      for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
        for (int qx = 0; qx < NQ; ++qx) {
          const double c = tp[qy * NQ + qx];
#pragma unroll
          for (int dy = 0; dy < ND; ++dy) a[dy][qx] = fma(c, c_mass_Bt[dy * NQ + qy], a[dy][qx]);
        }
    }
    if constexpr (LM) __syncthreads();                       // every thread has read T: it becomes the tile again
    if (has_slab) {
#pragma unroll
      for (int dy = 0; dy < ND; ++dy) {
        double r[ND];
#pragma unroll
        for (int dx = 0; dx < ND; ++dx) {
          double s = 0.0;
#pragma unroll
          for (int qx = 0; qx < NQ; ++qx) s = fma(a[dy][qx], c_mass_Bt[dx * NQ + qx], s);
          r[dx] = s;
        }
        if constexpr (LM) {                                  // result slab -> tile row t
          *reinterpret_cast<double2*>(T + t * YPITCH + 4 * dy) = make_double2(r[0], r[1]);
          *reinterpret_cast<double2*>(T + t * YPITCH + 4 * dy + 2) = make_double2(r[2], r[3]);
        } else {
          dbl4 o;
          o.x = yo[dy].x + r[0]; o.y = yo[dy].y + r[1]; o.z = yo[dy].z + r[2]; o.w = yo[dy].w + r[3];
          stg256(yp + 4 * dy, o);
        }
      }
    }
    if constexpr (LM) {                                      // tile -> pieces, added to the old Y pieces, stored line-major
      __syncthreads();
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        if (j * (BLOCK / 4) + (t >> 2) >= cnt * ND) continue;
        const double* d = T + (j * (BLOCK / 4) + (t >> 2)) * YPITCH + (t & 3) * 4;
        const double2 lo = *reinterpret_cast<const double2*>(d), hi = *reinterpret_cast<const double2*>(d + 2);
        dbl4 o;
        o.x = yo[j].x + lo.x; o.y = yo[j].y + lo.y; o.z = yo[j].z + hi.x; o.w = yo[j].w + hi.y;
        stg256(Y + e0 * 64 + j * (BLOCK * 4) + t * 4, o);
      }
    }
#pragma unroll
    for (int dy = 0; dy < ND; ++dy) xv[dy] = xn[dy];
    __syncthreads();      // T is rewritten by the next stage A
  }
}

// ------------------------------------------------------------------------------------------------
// shared pieces of the D1D = 3, Q1D = 4 kernels
// ------------------------------------------------------------------------------------------------
constexpr int PD = 3, PQ = 4, PQ2 = PQ * PQ;      // 16 pencils per element
constexpr int ROW = 3 * PQ2 + 1;                  // 49: three 4x4 slabs per (e,dz) task + 1 pad (odd)

// cooperative, coalesced copy of a batch of X (27 doubles per element) into shared memory
template <int BLOCK>
__device__ __forceinline__ void load_x27(double* xs, const double* __restrict__ X, int64_t e0, int cnt)
{
  const double* src = X + e0 * 27;
  for (int i = threadIdx.x; i < cnt * 27; i += BLOCK) xs[i] = __ldg(src + i);
}
template <int BLOCK>
__device__ __forceinline__ void add_y27(const double* ys, double* __restrict__ Y, int64_t e0, int cnt)
{
  double* dst = Y + e0 * 27;
  for (int i = threadIdx.x; i < cnt * 27; i += BLOCK) dst[i] += ys[i];
}


// ------------------------------------------------------------------------------------------------
// Shared-memory ring of the D1D = 3 kernels: stage s holds one batch of {D, X, Y}, filled by three
// bulk-async copies (cp.async.bulk + mbarrier) issued S-1 batches ahead of the math by thread 0.
// ------------------------------------------------------------------------------------------------
template <int E, int DSZ, int S>
struct pa_ring {
  double* Dr;   // [S][E*DSZ]
  double* Xr;   // [S][E*27]
  double* Yr;   // [S][E*27]
  double* T;    // [E*PD*ROW]
  double* R;    // [E*27]   stage-C results before the coalesced write-out
  unsigned long long* full;   // [S]
  static constexpr size_t bytes = sizeof(double) * (size_t)(S * E * DSZ + 2 * S * E * 27 + E * PD * ROW + E * 27) + sizeof(unsigned long long) * S;

  __device__ __forceinline__ void carve(unsigned char* smem)
  {
    Dr = reinterpret_cast<double*>(smem);
    Xr = Dr + S * E * DSZ;
    Yr = Xr + S * E * 27;
    T = Yr + S * E * 27;
    R = T + E * PD * ROW;
    full = reinterpret_cast<unsigned long long*>(R + E * 27);
    if (threadIdx.x == 0) {
      for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
  }
  // thread 0: request batch `batch` into stage `stage` (a ragged last batch is fetched by the threads themselves)
  __device__ __forceinline__ void issue(const double* __restrict__ D, const double* __restrict__ X, const double* __restrict__ Y,
                                        int64_t NE, int64_t batch, int stage)
  {
    const int64_t e0 = batch * E;
    if (NE - e0 >= E) {
      constexpr unsigned int DB = E * DSZ * sizeof(double), XB = E * 27 * sizeof(double);
      static_assert(DB % 16 == 0 && XB % 16 == 0, "bulk copies move multiples of 16 bytes");
      mbar_arrive_expect_tx(&full[stage], DB + 2 * XB);
      bulk_g2s(Dr + stage * E * DSZ, D + e0 * DSZ, DB, &full[stage]);
      bulk_g2s(Xr + stage * E * 27, X + e0 * 27, XB, &full[stage]);
      bulk_g2s(Yr + stage * E * 27, Y + e0 * 27, XB, &full[stage]);
    } else {
      mbar_arrive(&full[stage]);
    }
  }
  template <int BLOCK>
  __device__ __forceinline__ void fetch_ragged(const double* __restrict__ D, const double* __restrict__ X, const double* __restrict__ Y,
                                               int64_t e0, int cnt, int stage)
  {
    for (int i = threadIdx.x; i < cnt * DSZ; i += BLOCK) Dr[stage * E * DSZ + i] = D[e0 * DSZ + i];
    for (int i = threadIdx.x; i < cnt * 27; i += BLOCK) { Xr[stage * E * 27 + i] = X[e0 * 27 + i]; Yr[stage * E * 27 + i] = Y[e0 * 27 + i]; }
    __syncthreads();
  }
};

// ------------------------------------------------------------------------------------------------
// CONVECTION3DPA
// ------------------------------------------------------------------------------------------------
template <int E, int BLOCK, int S, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
convection3dpa_kernel(const double* __restrict__ D, const double* __restrict__ X, double* __restrict__ Y,
                      int64_t NE)
{
  extern __shared__ __align__(128) unsigned char pa_smem[];
  pa_ring<E, 192, S> ring;
  ring.carve(pa_smem);
  double* T = ring.T;

  const double* cB = c_conv;        // [q][d]
  const double* cG = c_conv + 12;   // [q][d]
  const double* cBt = c_conv + 24;  // [d][q]

  const int64_t nbatch = (NE + E - 1) / E;
  if (threadIdx.x == 0)
    for (int s = 0; s < S - 1; ++s) {
      const int64_t b = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (b < nbatch) ring.issue(D, X, Y, NE, b, s);
    }
  int it = 0;
  for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x, ++it) {
    const int stage = it % S;
    const int64_t e0 = batch * E;
    const int cnt = (int)((NE - e0) < E ? (NE - e0) : E);
    // stage (it-1) mod S was released by the barrier that ended the previous iteration
    const int64_t ahead = batch + (int64_t)(S - 1) * gridDim.x;
    if (threadIdx.x == 0 && ahead < nbatch) ring.issue(D, X, Y, NE, ahead, (it + S - 1) % S);
    mbar_wait(&ring.full[stage], (it / S) & 1);
    if (cnt < E) ring.template fetch_ragged<BLOCK>(D, X, Y, e0, cnt, stage);
    const double* XS = ring.Xr + stage * E * 27;
    const double* Dsm = ring.Dr + stage * E * 192;
    const double* YS = ring.Yr + stage * E * 27;
    double* RS = ring.R;

    // ---- stage A: (e,dz): Bu,Gu (contract x) then BBu, GBu, BGu (contract y)
    for (int t = threadIdx.x; t < cnt * PD; t += BLOCK) {
      const double* xp = XS + t * 9;
      double bu[PD][PQ], gu[PD][PQ];
#pragma unroll
      for (int dy = 0; dy < PD; ++dy) {
        const double x0 = xp[dy * 3], x1 = xp[dy * 3 + 1], x2 = xp[dy * 3 + 2];
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) {
          bu[dy][qx] = fma(cB[qx * PD + 2], x2, fma(cB[qx * PD + 1], x1, cB[qx * PD] * x0));
          gu[dy][qx] = fma(cG[qx * PD + 2], x2, fma(cG[qx * PD + 1], x1, cG[qx * PD] * x0));
        }
      }
      double* tp = T + t * ROW;
#pragma unroll
      for (int qy = 0; qy < PQ; ++qy)
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) {
          double bbu = 0.0, gbu = 0.0, bgu = 0.0;
#pragma unroll
          for (int dy = 0; dy < PD; ++dy) {
            bbu = fma(cB[qy * PD + dy], bu[dy][qx], bbu);
            gbu = fma(cG[qy * PD + dy], bu[dy][qx], gbu);
            bgu = fma(cB[qy * PD + dy], gu[dy][qx], bgu);
          }
          tp[qy * PQ + qx] = bbu;
          tp[PQ2 + qy * PQ + qx] = gbu;
          tp[2 * PQ2 + qy * PQ + qx] = bgu;
        }
    }
    __syncthreads();

    // ---- stage B: (e,pencil): contract z, apply D, contract z back (into slab 0)
    for (int p = threadIdx.x; p < cnt * PQ2; p += BLOCK) {
      const int e = p >> 4, pen = p & 15;
      const double* dp = Dsm + e * 192 + pen;
      double o[3][PQ];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int qz = 0; qz < PQ; ++qz) o[c][qz] = dp[c * 64 + qz * PQ2];
      double* tp = T + e * PD * ROW + pen;
      double bbu[PD], gbu[PD], bgu[PD];
#pragma unroll
      for (int dz = 0; dz < PD; ++dz) {
        bbu[dz] = tp[dz * ROW]; gbu[dz] = tp[dz * ROW + PQ2]; bgu[dz] = tp[dz * ROW + 2 * PQ2];
      }
      double dgu[PQ];
#pragma unroll
      for (int qz = 0; qz < PQ; ++qz) {
        double gz = 0.0, gy = 0.0, gx = 0.0;
#pragma unroll
        for (int dz = 0; dz < PD; ++dz) {
          gz = fma(cG[qz * PD + dz], bbu[dz], gz);     // GBBu
          gy = fma(cB[qz * PD + dz], gbu[dz], gy);     // BGBu
          gx = fma(cB[qz * PD + dz], bgu[dz], gx);     // BBGu
        }
        dgu[qz] = fma(o[2][qz], gz, fma(o[1][qz], gy, o[0][qz] * gx));
      }
#pragma unroll
      for (int dz = 0; dz < PD; ++dz) {
        double s = 0.0;
#pragma unroll
        for (int qz = 0; qz < PQ; ++qz) s = fma(cBt[dz * PQ + qz], dgu[qz], s);
        tp[dz * ROW] = s;
      }
    }
    __syncthreads();

    // ---- stage C: (e,dz): contract y then x; results go to XS (reused as Y staging)
    for (int t = threadIdx.x; t < cnt * PD; t += BLOCK) {
      const double* tp = T + t * ROW;
      double a[PD][PQ];
#pragma unroll
      for (int dy = 0; dy < PD; ++dy)
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) a[dy][qx] = 0.0;
#pragma unroll
      for (int qy = 0; qy < PQ; ++qy)
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) {
          const double v = tp[qy * PQ + qx];
#pragma unroll
          for (int dy = 0; dy < PD; ++dy) a[dy][qx] = fma(cBt[dy * PQ + qy], v, a[dy][qx]);
        }
      double* yp = RS + t * 9;
#pragma unroll
      for (int dy = 0; dy < PD; ++dy)
#pragma unroll
        for (int dx = 0; dx < PD; ++dx) {
          double s = 0.0;
#pragma unroll
          for (int qx = 0; qx < PQ; ++qx) s = fma(cBt[dx * PQ + qx], a[dy][qx], s);
          yp[dy * 3 + dx] = s;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 27; i += BLOCK) Y[e0 * 27 + i] = YS[i] + RS[i];
    __syncthreads();      // the ring stage, T and RS are reused by the next iterations
  }
}

// ------------------------------------------------------------------------------------------------
// DIFFUSION3DPA
// ------------------------------------------------------------------------------------------------
// Builds the four dense tables from the reference's aliased half-stored basis array, reproducing
// the Base_Seq fill order (DIFFUSION3DPA-Seq.cpp:45-49, 75-79; DIFFUSION3DPA.hpp:245-268, 301-305).
__global__ void diffusion_tables_kernel(const double* __restrict__ Basis, const double* __restrict__ dBasis,
                                        double* __restrict__ out /*48*/)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto qi = [](int q, int d) { return (q <= d) ? q : PQ - 1 - q; };
  auto dj = [](int q, int d) { return (q <= d) ? d : PD - 1 - d; };
  auto qk = [](int q, int d) { return (q <= d) ? PQ - 1 - q : q; };
  auto dl = [](int q, int d) { return (q <= d) ? PD - 1 - d : d; };
  auto sg = [](int q, int d) { return (q <= d) ? -1.0 : 1.0; };
  double t[12];
  for (int i = 0; i < 12; ++i) t[i] = 0.0;
  for (int dy = 0; dy < PD; ++dy)          // fill #1: B[i][j] then G[k][l], viewed [4][3]
    for (int qx = 0; qx < PQ; ++qx) {
      t[qi(qx, dy) * PD + dj(qx, dy)] = Basis[qx + PQ * dy];
      t[qk(qx, dy) * PD + dl(qx, dy)] = dBasis[qx + PQ * dy] * sg(qx, dy);
    }
  for (int q = 0; q < PQ; ++q)
    for (int d = 0; d < PD; ++d) {
      out[q * PD + d] = t[qi(q, d) * PD + dj(q, d)];                     // B(q,d) as steps 3-5 read it
      out[12 + q * PD + d] = t[qk(q, d) * PD + dl(q, d)] * sg(q, d);     // G(q,d) * sign
    }
  for (int d = 0; d < PD; ++d)             // fill #2 on top of fill #1, viewed [3][4]
    for (int q = 0; q < PQ; ++q) {
      t[dj(q, d) * PQ + qi(q, d)] = Basis[q + PQ * d];
      t[dl(q, d) * PQ + qk(q, d)] = dBasis[q + PQ * d] * sg(q, d);
    }
  for (int d = 0; d < PD; ++d)
    for (int q = 0; q < PQ; ++q) {
      out[24 + d * PQ + q] = t[dj(q, d) * PQ + qi(q, d)];                // Bt(d,q) as steps 7-9 read it
      out[36 + d * PQ + q] = t[dl(q, d) * PQ + qk(q, d)] * sg(q, d);     // Gt(d,q) * sign
    }
}

template <int E, int BLOCK, int S, int MINB, bool SYMM>
__global__ void __launch_bounds__(BLOCK, MINB)
diffusion3dpa_kernel(const double* __restrict__ D, const double* __restrict__ X, double* __restrict__ Y,
                     int64_t NE)
{
  extern __shared__ __align__(128) unsigned char pa_smem[];
  pa_ring<E, 384, S> ring;
  ring.carve(pa_smem);
  double* T = ring.T;

  const double* cB = c_diff;         // [q][d]
  const double* cG = c_diff + 12;    // [q][d], sign folded in
  const double* cBt = c_diff + 24;   // [d][q]
  const double* cGt = c_diff + 36;   // [d][q], sign folded in

  const int64_t nbatch = (NE + E - 1) / E;
  if (threadIdx.x == 0)
    for (int s = 0; s < S - 1; ++s) {
      const int64_t b = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (b < nbatch) ring.issue(D, X, Y, NE, b, s);
    }
  int it = 0;
  for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x, ++it) {
    const int stage = it % S;
    const int64_t e0 = batch * E;
    const int cnt = (int)((NE - e0) < E ? (NE - e0) : E);
    // stage (it-1) mod S was released by the barrier that ended the previous iteration
    const int64_t ahead = batch + (int64_t)(S - 1) * gridDim.x;
    if (threadIdx.x == 0 && ahead < nbatch) ring.issue(D, X, Y, NE, ahead, (it + S - 1) % S);
    mbar_wait(&ring.full[stage], (it / S) & 1);
    if (cnt < E) ring.template fetch_ragged<BLOCK>(D, X, Y, e0, cnt, stage);
    const double* XS = ring.Xr + stage * E * 27;
    const double* Dsm = ring.Dr + stage * E * 384;
    const double* YS = ring.Yr + stage * E * 27;
    double* RS = ring.R;

    // ---- stage A (steps 3,4): (e,dz) -> DQQ0 = G_x B_y, DQQ1 = B_x G_y, DQQ2 = B_x B_y
    for (int t = threadIdx.x; t < cnt * PD; t += BLOCK) {
      const double* xp = XS + t * 9;
      double b0[PD][PQ], g0[PD][PQ];     // DDQ0 (B in x), DDQ1 (G in x)
#pragma unroll
      for (int dy = 0; dy < PD; ++dy) {
        const double x0 = xp[dy * 3], x1 = xp[dy * 3 + 1], x2 = xp[dy * 3 + 2];
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) {
          b0[dy][qx] = fma(x2, cB[qx * PD + 2], fma(x1, cB[qx * PD + 1], x0 * cB[qx * PD]));
          g0[dy][qx] = fma(x2, cG[qx * PD + 2], fma(x1, cG[qx * PD + 1], x0 * cG[qx * PD]));
        }
      }
      double* tp = T + t * ROW;
#pragma unroll
      for (int qy = 0; qy < PQ; ++qy)
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) {
          double u = 0.0, v = 0.0, w = 0.0;
#pragma unroll
          for (int dy = 0; dy < PD; ++dy) {
            u = fma(g0[dy][qx], cB[qy * PD + dy], u);
            v = fma(b0[dy][qx], cG[qy * PD + dy], v);
            w = fma(b0[dy][qx], cB[qy * PD + dy], w);
          }
          tp[qy * PQ + qx] = u;
          tp[PQ2 + qy * PQ + qx] = v;
          tp[2 * PQ2 + qy * PQ + qx] = w;
        }
    }
    __syncthreads();

    // ---- stage B (steps 5 and 9's z contraction): (e,pencil)
    for (int p = threadIdx.x; p < cnt * PQ2; p += BLOCK) {
      const int e = p >> 4, pen = p & 15;
      const double* dp = Dsm + e * 384 + pen;             // stride SYM = 6 slabs per element
      double* tp = T + e * PD * ROW + pen;
      double q0[PD], q1[PD], q2[PD];
#pragma unroll
      for (int dz = 0; dz < PD; ++dz) {
        q0[dz] = tp[dz * ROW]; q1[dz] = tp[dz * ROW + PQ2]; q2[dz] = tp[dz * ROW + 2 * PQ2];
      }
      double r0[PQ], r1[PQ], r2[PQ];
#pragma unroll
      for (int qz = 0; qz < PQ; ++qz) {
        double gX = 0.0, gY = 0.0, gZ = 0.0;
#pragma unroll
        for (int dz = 0; dz < PD; ++dz) {
          gX = fma(q0[dz], cB[qz * PD + dz], gX);
          gY = fma(q1[dz], cB[qz * PD + dz], gY);
          gZ = fma(q2[dz], cG[qz * PD + dz], gZ);
        }
        const double* dq = dp + qz * PQ2;
        const double O11 = dq[0], O12 = dq[64], O13 = dq[128];
        double O21, O22, O23, O31, O32, O33;
        if (SYMM) {
          O21 = O12; O22 = dq[192]; O23 = dq[256];
          O31 = O13; O32 = O23;     O33 = dq[320];
        } else {   // DIFFUSION3DPA.hpp:389-397 reads slabs 3..8 (beyond the SYM=6 stride, as the reference does)
          // slabs 6..8 lie beyond this element's SYM = 6 slabs, in the next element's D, exactly as the
          // reference indexes them; they are read from global memory (the ring holds E elements only)
          const double* gq = D + (e0 + e) * 384 + pen + qz * PQ2;
          O21 = dq[192]; O22 = dq[256]; O23 = dq[320];
          O31 = __ldg(gq + 384); O32 = __ldg(gq + 448); O33 = __ldg(gq + 512);
        }
        r0[qz] = fma(O13, gZ, fma(O12, gY, O11 * gX));
        r1[qz] = fma(O23, gZ, fma(O22, gY, O21 * gX));
        r2[qz] = fma(O33, gZ, fma(O32, gY, O31 * gX));
      }
#pragma unroll
      for (int dz = 0; dz < PD; ++dz) {
        double u = 0.0, v = 0.0, w = 0.0;
#pragma unroll
        for (int qz = 0; qz < PQ; ++qz) {
          u = fma(r0[qz], cBt[dz * PQ + qz], u);
          v = fma(r1[qz], cBt[dz * PQ + qz], v);
          w = fma(r2[qz], cGt[dz * PQ + qz], w);
        }
        tp[dz * ROW] = u; tp[dz * ROW + PQ2] = v; tp[dz * ROW + 2 * PQ2] = w;
      }
    }
    __syncthreads();

    // ---- stage C (steps 8,7): (e,dz): field0: Bt_y Gt_x, field1: Gt_y Bt_x, field2: Bt_y Bt_x
    for (int t = threadIdx.x; t < cnt * PD; t += BLOCK) {
      const double* tp = T + t * ROW;
      double a0[PD][PQ], a1[PD][PQ], a2[PD][PQ];
#pragma unroll
      for (int dy = 0; dy < PD; ++dy)
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) { a0[dy][qx] = 0.0; a1[dy][qx] = 0.0; a2[dy][qx] = 0.0; }
#pragma unroll
      for (int qy = 0; qy < PQ; ++qy)
#pragma unroll
        for (int qx = 0; qx < PQ; ++qx) {
          const double f0 = tp[qy * PQ + qx], f1 = tp[PQ2 + qy * PQ + qx], f2 = tp[2 * PQ2 + qy * PQ + qx];
#pragma unroll
          for (int dy = 0; dy < PD; ++dy) {
            a0[dy][qx] = fma(f0, cBt[dy * PQ + qy], a0[dy][qx]);
            a1[dy][qx] = fma(f1, cGt[dy * PQ + qy], a1[dy][qx]);
            a2[dy][qx] = fma(f2, cBt[dy * PQ + qy], a2[dy][qx]);
          }
        }
      double* yp = RS + t * 9;
#pragma unroll
      for (int dy = 0; dy < PD; ++dy)
#pragma unroll
        for (int dx = 0; dx < PD; ++dx) {
          double u = 0.0, v = 0.0, w = 0.0;
#pragma unroll
          for (int qx = 0; qx < PQ; ++qx) {
            u = fma(a0[dy][qx], cGt[dx * PQ + qx], u);
            v = fma(a1[dy][qx], cBt[dx * PQ + qx], v);
            w = fma(a2[dy][qx], cBt[dx * PQ + qx], w);
          }
          yp[dy * 3 + dx] = (u + v + w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 27; i += BLOCK) Y[e0 * 27 + i] = YS[i] + RS[i];
    __syncthreads();      // the ring stage, T and RS are reused by the next iterations
  }
}

// tables for MASS / CONVECTION: plain re-indexing of the caller's device arrays
__global__ void mass_tables_kernel(const double* __restrict__ B, const double* __restrict__ Bt, double* out)
{
  const int i = threadIdx.x;
  if (i < 20) {
    { const int q = i / 4, d = i % 4; out[i] = B[q + 5 * d]; }            // Bsmem[q][d]
    { const int d = i / 5, q = i % 5; out[20 + i] = Bt[q + 4 * d]; }      // Btsmem[d][q]
  }
}
__global__ void conv_tables_kernel(const double* __restrict__ B, const double* __restrict__ Bt,
                                   const double* __restrict__ G, double* out)
{
  const int i = threadIdx.x;
  if (i < 12) {
    { const int q = i / 3, d = i % 3; out[i] = B[q + 4 * d]; out[12 + i] = G[q + 4 * d]; }   // cpa_B / cpa_G
    { const int d = i / 4, q = i % 4; out[24 + i] = Bt[d + 3 * q]; }                         // cpa_Bt(d,q)
  }
}

constexpr int PA_E = 32, PA_BLOCK = 128;

inline int pa_grid(const rpb200_ctx* ctx, int kid, int64_t NE)
{
  const int64_t nbatch = (NE + PA_E - 1) / PA_E;
  const int cps = ctx->tune[kid].ctas_per_sm;
  if (cps <= 0) return (int)(nbatch > 0x7fffffff ? 0x7fffffff : nbatch);
  const int64_t g = (int64_t)ctx->sm_count * cps;
  return (int)(g < nbatch ? g : nbatch);
}

// persistent launch of a ring kernel: MINB CTAs per SM, dynamic shared memory = the ring
template <int E, int BLOCK, int S, int MINB, int DSZ, typename K>
cudaError_t launch_ring_kernel(K kernel, const rpb200_ctx* ctx, const double* D, const double* X, double* Y, int64_t NE,
                               cudaStream_t st)
{
  constexpr size_t smem = pa_ring<E, DSZ, S>::bytes;
  static_assert(smem * MINB <= 227 * 1024, "ring does not fit");
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t nbatch = (NE + E - 1) / E;
  int64_t grid = (int64_t)ctx->sm_count * MINB;
  if (grid > nbatch) grid = nbatch;
  kernel<<<(int)grid, BLOCK, smem, st>>>(D, X, Y, NE);
  return cudaGetLastError();
}

template <int E, int BLOCK, int MINB, bool YRING = false, int DS = 2, bool LM = false>
cudaError_t launch_mass(const rpb200_ctx* ctx, const double* D, const double* X, double* Y, int64_t NE, cudaStream_t st)
{
  constexpr size_t smem = sizeof(double) * (DS * E * 125 + E * 4 * 25 + (YRING ? 2 * E * 4 * 18 : 0)) + 2 * sizeof(unsigned long long);
  static_assert(smem * MINB <= 227 * 1024, "ring does not fit");
  cudaError_t e = cudaFuncSetAttribute(mass3dpa_kernel<E, BLOCK, MINB, YRING, DS, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t nbatch = (NE + E - 1) / E;
  int64_t grid = (int64_t)ctx->sm_count * MINB;
  if (grid > nbatch) grid = nbatch;
  mass3dpa_kernel<E, BLOCK, MINB, YRING, DS, LM><<<(int)grid, BLOCK, smem, st>>>(D, X, Y, NE);
  return cudaGetLastError();
}

}  // namespace

// ---- one set of __constant__ tables per device: order a call behind the previous PA call on another stream ------------
namespace {
struct pa_device_state { cudaStream_t last; cudaEvent_t done; int have_event; int recorded; };
pa_device_state g_pa[64];
volatile int g_pa_lock = 0;
struct pa_lock { pa_lock() { while (__sync_lock_test_and_set(&g_pa_lock, 1)) { } } ~pa_lock() { __sync_lock_release(&g_pa_lock); } };

bool pa_capturing(cudaStream_t st)
{
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  return st != nullptr && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone;
}
// before the tables are refreshed on `st`
int pa_order_begin(const rpb200_ctx* ctx, cudaStream_t st)
{
  if (ctx->device < 0 || ctx->device >= 64 || pa_capturing(st)) return 0;   // captured calls: ordered by the graph's author
  pa_lock guard;
  pa_device_state& P = g_pa[ctx->device];
  if (P.recorded && P.last != st) RPB_CHECK(cudaStreamWaitEvent(st, P.done, 0));
  return 0;
}
// after the kernel that reads the tables has been enqueued on `st`
int pa_order_end(const rpb200_ctx* ctx, cudaStream_t st)
{
  if (ctx->device < 0 || ctx->device >= 64 || pa_capturing(st)) return 0;
  pa_lock guard;
  pa_device_state& P = g_pa[ctx->device];
  if (!P.have_event) { RPB_CHECK(cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming)); P.have_event = 1; }
  RPB_CHECK(cudaEventRecord(P.done, st));
  P.last = st;
  P.recorded = 1;
  return 0;
}
}  // namespace

extern "C" int rpb200_mass3dpa(rpb200_ctx* ctx, const double* B, const double* Bt, const double* D,
                               const double* X, double* Y, int64_t NE, rpb200_stream_t s)
{
  if (!ctx || NE < 0 || (NE > 0 && (!B || !Bt || !D || !X || !Y))) return RPB200_EINVAL;
  if (NE == 0) return 0;
  if (!rpb_aligned(X, 32) || !rpb_aligned(Y, 32)) return RPB200_EINVAL;   // 512-byte elements
  if (!rpb_aligned(D, 16)) return RPB200_EINVAL;
  cudaStream_t st = rpb_stream(s);
  RPB_SCRATCH(sc, ctx, st);
  mass_tables_kernel<<<1, 32, 0, st>>>(B, Bt, sc->d_basis_tables);
  RPB_LAUNCH_CHECK();
  { const int rc_ = pa_order_begin(ctx, st); if (rc_ != 0) return rc_; }
  RPB_CHECK(cudaMemcpyToSymbolAsync(c_mass_B, sc->d_basis_tables, 20 * sizeof(double), 0, cudaMemcpyDeviceToDevice, st));
  RPB_CHECK(cudaMemcpyToSymbolAsync(c_mass_Bt, sc->d_basis_tables + 20, 20 * sizeof(double), 0, cudaMemcpyDeviceToDevice, st));
  // elements per CTA / threads / CTAs per SM / D stages: 8/32/11/1 with line-major accesses is the best of the sweeps in
  // profiles/r01_pa_variants.md
  switch (ctx->tune[RPB_K_MASS3DPA].unroll) {
    case 10: RPB_CHECK((launch_mass<16, 64, 4>(ctx, D, X, Y, NE, st))); break;
    case 12: RPB_CHECK((launch_mass<8, 64, 6>(ctx, D, X, Y, NE, st))); break;
    case 13: RPB_CHECK((launch_mass<16, 128, 3>(ctx, D, X, Y, NE, st))); break;
    case 14: RPB_CHECK((launch_mass<8, 32, 9>(ctx, D, X, Y, NE, st))); break;
    case 16: RPB_CHECK((launch_mass<8, 32, 7, true>(ctx, D, X, Y, NE, st))); break;
    case 19: RPB_CHECK((launch_mass<4, 32, 16>(ctx, D, X, Y, NE, st))); break;
    case 20: RPB_CHECK((launch_mass<4, 32, 12>(ctx, D, X, Y, NE, st))); break;
    case 21: RPB_CHECK((launch_mass<8, 64, 7>(ctx, D, X, Y, NE, st))); break;
    case 22: RPB_CHECK((launch_mass<4, 32, 14>(ctx, D, X, Y, NE, st))); break;
    case 23: RPB_CHECK((launch_mass<8, 32, 10, false, 1>(ctx, D, X, Y, NE, st))); break;
    case 24: RPB_CHECK((launch_mass<8, 32, 12, false, 1>(ctx, D, X, Y, NE, st))); break;
    case 25: RPB_CHECK((launch_mass<8, 32, 14, false, 1>(ctx, D, X, Y, NE, st))); break;
    case 26: RPB_CHECK((launch_mass<8, 32, 8, false, 1>(ctx, D, X, Y, NE, st))); break;
    case 27: RPB_CHECK((launch_mass<8, 32, 11, false, 1>(ctx, D, X, Y, NE, st))); break;
    case 28: RPB_CHECK((launch_mass<8, 32, 13, false, 1>(ctx, D, X, Y, NE, st))); break;
    case 29: RPB_CHECK((launch_mass<16, 64, 6, false, 1>(ctx, D, X, Y, NE, st))); break;
    case 17: RPB_CHECK((launch_mass<8, 32, 6, true>(ctx, D, X, Y, NE, st))); break;
    case 18: RPB_CHECK((launch_mass<16, 64, 3, true>(ctx, D, X, Y, NE, st))); break;
    case 15: RPB_CHECK((launch_mass<16, 64, 5>(ctx, D, X, Y, NE, st))); break;
    case 30: RPB_CHECK((launch_mass<8, 32, 8>(ctx, D, X, Y, NE, st))); break;
    case 31:                                                  // line-major X / Y accesses (whole batches only)
      if (NE % 8 == 0) RPB_CHECK((launch_mass<8, 32, 12, false, 1, true>(ctx, D, X, Y, NE, st)));
      else RPB_CHECK((launch_mass<8, 32, 12, false, 1>(ctx, D, X, Y, NE, st)));
      break;
    case 32:
      if (NE % 8 == 0) RPB_CHECK((launch_mass<8, 32, 11, false, 1, true>(ctx, D, X, Y, NE, st)));
      else RPB_CHECK((launch_mass<8, 32, 12, false, 1>(ctx, D, X, Y, NE, st)));
      break;
    case 33:
      if (NE % 8 == 0) RPB_CHECK((launch_mass<8, 32, 10, false, 1, true>(ctx, D, X, Y, NE, st)));
      else RPB_CHECK((launch_mass<8, 32, 12, false, 1>(ctx, D, X, Y, NE, st)));
      break;
    case 34:
      if (NE % 8 == 0) RPB_CHECK((launch_mass<8, 32, 9, false, 1, true>(ctx, D, X, Y, NE, st)));
      else RPB_CHECK((launch_mass<8, 32, 12, false, 1>(ctx, D, X, Y, NE, st)));
      break;
    default:                                                  // = 32: 6905 GB/s at NE = 4 M; 31 (12 CTAs): 6800; slab-per-thread accesses (24): 6250;
                                                              // with two D stages and 8 CTAs (30): 5760 (profiles/r01_pa_variants.md)
      if (NE % 8 == 0) RPB_CHECK((launch_mass<8, 32, 11, false, 1, true>(ctx, D, X, Y, NE, st)));
      else RPB_CHECK((launch_mass<8, 32, 12, false, 1>(ctx, D, X, Y, NE, st)));
      break;
  }
  RPB_LAUNCH_CHECK();
  return pa_order_end(ctx, st);
}

extern "C" int rpb200_convection3dpa(rpb200_ctx* ctx, const double* Basis, const double* tBasis,
                                     const double* dBasis, const double* D, const double* X, double* Y,
                                     int64_t NE, rpb200_stream_t s)
{
  if (!ctx || NE < 0 || (NE > 0 && (!Basis || !tBasis || !dBasis || !D || !X || !Y))) return RPB200_EINVAL;
  if (NE == 0) return 0;
  if (!rpb_aligned(D, 16) || !rpb_aligned(X, 16) || !rpb_aligned(Y, 16)) return RPB200_EINVAL;
  cudaStream_t st = rpb_stream(s);
  RPB_SCRATCH(sc, ctx, st);
  conv_tables_kernel<<<1, 32, 0, st>>>(Basis, tBasis, dBasis, sc->d_basis_tables);
  RPB_LAUNCH_CHECK();
  { const int rc_ = pa_order_begin(ctx, st); if (rc_ != 0) return rc_; }
  RPB_CHECK(cudaMemcpyToSymbolAsync(c_conv, sc->d_basis_tables, 36 * sizeof(double), 0, cudaMemcpyDeviceToDevice, st));
  // tuning field `unroll` selects the launch shape {elements per CTA, threads, ring stages, CTAs per SM}
  // (sweep: profiles/r01_pa_variants.md)
#define RPB_CONV(E, B, S, M) RPB_CHECK((launch_ring_kernel<E, B, S, M, 192>(convection3dpa_kernel<E, B, S, M>, ctx, D, X, Y, NE, st)))
  switch (ctx->tune[RPB_K_CONVECTION3DPA].unroll) {
    case 10: RPB_CONV(16, 256, 2, 2); break;
    case 11: RPB_CONV(8, 128, 2, 4); break;
    case 12: RPB_CONV(8, 128, 2, 3); break;
    case 13: RPB_CONV(8, 128, 3, 3); break;
    case 14: RPB_CONV(16, 128, 2, 2); break;
    case 15: RPB_CONV(8, 64, 2, 5); break;
    case 16: RPB_CONV(4, 64, 2, 8); break;
    case 17: RPB_CONV(16, 256, 3, 1); break;
    default: RPB_CONV(8, 128, 2, 5); break;      // 6539 GB/s at NE = 4 M (8/128/2/4: 6028, 16/256/2/2: 5612)
  }
#undef RPB_CONV
  RPB_LAUNCH_CHECK();
  return pa_order_end(ctx, st);
}

extern "C" int rpb200_diffusion3dpa(rpb200_ctx* ctx, const double* Basis, const double* dBasis,
                                    const double* D, const double* X, double* Y, int64_t NE,
                                    int symmetric, rpb200_stream_t s)
{
  if (!ctx || NE < 0 || (NE > 0 && (!Basis || !dBasis || !D || !X || !Y))) return RPB200_EINVAL;
  if (NE == 0) return 0;
  if (!rpb_aligned(D, 16) || !rpb_aligned(X, 16) || !rpb_aligned(Y, 16)) return RPB200_EINVAL;
  cudaStream_t st = rpb_stream(s);
  RPB_SCRATCH(sc, ctx, st);
  diffusion_tables_kernel<<<1, 32, 0, st>>>(Basis, dBasis, sc->d_basis_tables);
  RPB_LAUNCH_CHECK();
  { const int rc_ = pa_order_begin(ctx, st); if (rc_ != 0) return rc_; }
  RPB_CHECK(cudaMemcpyToSymbolAsync(c_diff, sc->d_basis_tables, 48 * sizeof(double), 0, cudaMemcpyDeviceToDevice, st));
  if (symmetric)
    RPB_CHECK((launch_ring_kernel<8, 128, 2, 3, 384>(diffusion3dpa_kernel<8, 128, 2, 3, true>, ctx, D, X, Y, NE, st)));
  else
    RPB_CHECK((launch_ring_kernel<8, 128, 2, 3, 384>(diffusion3dpa_kernel<8, 128, 2, 3, false>, ctx, D, X, Y, NE, st)));
  RPB_LAUNCH_CHECK();
  return pa_order_end(ctx, st);
}
