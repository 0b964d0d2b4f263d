/*
 * rpb200.h -- C ABI of the Base_B200 kernel variant for the RAJA Performance Suite.
 *
 * This is the drop-in boundary: a reference-side `run B200 variant` method (one per
 * kernel class, the analogue of `runCudaVariant`, KernelBase.hpp:457-487) calls
 * exactly one of these entry points between startTimer() and stopTimer().
 * INTEGRATION.md shows the reference-side stub for every kernel.
 *
 * Conventions
 *   - plain pointers + int64_t sizes; every data pointer is DEVICE memory owned by
 *     the caller (the kernel object's setUp/tearDown, KernelBase.cpp:359-377);
 *   - `stream` is a cudaStream_t passed as an opaque pointer (NULL = legacy default
 *     stream); every call is asynchronous on that stream, nothing synchronises;
 *   - return 0 on success, otherwise a cudaError_t value (or RPB200_EINVAL);
 *     no exceptions, no CPU fallback: without a usable sm_100 device
 *     rpb200_create() fails and nothing else may be called;
 *   - all file:line citations are relative to /root/reference/src.
 */
#ifndef RPB200_H
#define RPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPB200_EINVAL (-22)

typedef struct rpb200_ctx rpb200_ctx;   /* per-device context (scratch, SM count, tunings) */
typedef void* rpb200_stream_t;          /* cudaStream_t */

/* ---- context ---------------------------------------------------------------------
 * Replaces the per-kernel camp::resources::Cuda handle + reducer scratch the
 * reference's GPU variants build (KernelBase.hpp:242-248, GPUUtils.hpp:250-330). */
int         rpb200_create(int device, rpb200_ctx** out);
void        rpb200_destroy(rpb200_ctx* ctx);
const char* rpb200_error_string(int err);
int         rpb200_sm_count(const rpb200_ctx* ctx);
const char* rpb200_version(void);

/* Launch tuning, the analogue of the reference's block-size tunings
 * (GPUUtils.hpp:345-373).  kernel = full kernel name ("Stream_TRIAD");
 * <=0 keeps the built-in default for that field.                                    */
int rpb200_set_tuning(rpb200_ctx* ctx, const char* kernel, int block_size,
                      int ctas_per_sm, int unroll);

/* ---- Stream group ----------------------------------------------------------------
 * stream/COPY-Cuda.cpp:26, MUL-Cuda.cpp:26, ADD-Cuda.cpp:27, TRIAD-Cuda.cpp:26-58.
 * TRIAD/MUL use separate multiply and add roundings (no FMA), so every output is
 * bit-identical to Base_Seq.                                                        */
int rpb200_stream_copy (rpb200_ctx*, double* c, const double* a, int64_t n, rpb200_stream_t);
int rpb200_stream_mul  (rpb200_ctx*, double* b, const double* c, double alpha, int64_t n, rpb200_stream_t);
int rpb200_stream_add  (rpb200_ctx*, double* c, const double* a, const double* b, int64_t n, rpb200_stream_t);
int rpb200_stream_triad(rpb200_ctx*, double* a, const double* b, const double* c, double alpha,
                        int64_t n, rpb200_stream_t);
/* stream/DOT-Cuda.cpp:28-99.  *d_out = (accumulate ? *d_out : 0) + init + sum a[i]*b[i].
 * d_out is a device double; `accumulate` lets the caller keep DOT's running m_dot
 * (DOT-Seq.cpp:45) on the device instead of syncing every rep.  Deterministic:
 * fixed partition, fixed combine order.                                             */
int rpb200_stream_dot  (rpb200_ctx*, const double* a, const double* b, int64_t n, double init,
                        double* d_out, int accumulate, rpb200_stream_t);

/* ---- Algorithm group -------------------------------------------------------------
 * algorithm/REDUCE_SUM-Cuda.cpp:31-171: *d_out = init + sum x[i].                   */
int rpb200_reduce_sum(rpb200_ctx*, const double* x, int64_t n, double init, double* d_out,
                      rpb200_stream_t);
/* algorithm/SCAN-Cuda.cpp:34-188 + common/CudaGridScan.hpp: exclusive prefix sum,
 * single pass (decoupled look-back), no per-call memset.                            */
int rpb200_scan_exclusive(rpb200_ctx*, const double* x, double* y, int64_t n, rpb200_stream_t);
/* algorithm/SORT-Cuda.cpp:35-43 (RAJA::sort -> cub::DeviceRadixSort::SortKeys):
 * ascending in-place sort of n doubles (IEEE total order on non-NaN values, -0 < +0).
 * scratch: rpb200_sort_scratch_bytes(n, pairs) bytes of device memory.              */
size_t rpb200_sort_scratch_bytes(int64_t n, int pairs);
int rpb200_sort_keys_f64(rpb200_ctx*, double* keys, int64_t n, void* scratch, size_t scratch_bytes,
                         rpb200_stream_t);
/* algorithm/SORTPAIRS-Cuda.cpp:35-43: stable sort of (key, value) by key.           */
int rpb200_sort_pairs_f64(rpb200_ctx*, double* keys, double* vals, int64_t n, void* scratch,
                          size_t scratch_bytes, rpb200_stream_t);

/* ---- Apps group ------------------------------------------------------------------
 * apps/MASS3DPA-Cuda.cpp:25-112 (D1D=4,Q1D=5): Y += Mass(B,Bt,D) X, NE elements.    */
int rpb200_mass3dpa(rpb200_ctx*, const double* B, const double* Bt, const double* D,
                    const double* X, double* Y, int64_t NE, rpb200_stream_t);
/* apps/DIFFUSION3DPA-Cuda.cpp:25-130 (D1D=3,Q1D=4,SYM=6).  The basis tables follow the
 * Base_Seq fill order of the reference's aliased shared array (DIFFUSION3DPA.hpp:301-305). */
int rpb200_diffusion3dpa(rpb200_ctx*, const double* Basis, const double* dBasis, const double* D,
                         const double* X, double* Y, int64_t NE, int symmetric, rpb200_stream_t);
/* apps/CONVECTION3DPA-Cuda.cpp:24-127 (D1D=3,Q1D=4,VDIM=3).                          */
int rpb200_convection3dpa(rpb200_ctx*, const double* Basis, const double* tBasis,
                          const double* dBasis, const double* D, const double* X, double* Y,
                          int64_t NE, rpb200_stream_t);
/* apps/LTIMES-Cuda.cpp:44-102: phi[z][g][m] += sum_d ell[m][d] * psi[z][g][d].       */
int rpb200_ltimes(rpb200_ctx*, double* phi, const double* ell, const double* psi,
                  int64_t num_d, int64_t num_g, int64_t num_m, int64_t num_z, rpb200_stream_t);

/* ---- Comm group ------------------------------------------------------------------
 * comm/HALO_PACKING_FUSED-Cuda.cpp:52-197, HALO_EXCHANGE_FUSED-Cuda.cpp:52-206.
 * One descriptor per (neighbour, variable) segment -- the reference's
 * (buffer, list, var, len) tuples (HALO_PACKING_FUSED.hpp:62-71), held in DEVICE
 * memory.  pack:   buffer[i] = var[list[i]];  unpack: var[list[i]] = buffer[i].     */
typedef struct rpb200_halo_seg {
  double*    buffer;   /* contiguous message segment (may be a peer-GPU pointer)      */
  const int* list;     /* index list (Int_type, HALO_base.cpp:197-254)                */
  double*    var;      /* the grid variable                                           */
  int64_t    len;      /* elements in this segment                                    */
  int64_t    work_begin; /* exclusive prefix of ceil(len/chunk) over segments          */
} rpb200_halo_seg;

/* elements one work chunk covers; the host fills work_begin in these units            */
int     rpb200_halo_chunk(void);
int rpb200_halo_pack  (rpb200_ctx*, const rpb200_halo_seg* d_segs, int nsegs, int64_t total_chunks,
                       rpb200_stream_t);
int rpb200_halo_unpack(rpb200_ctx*, const rpb200_halo_seg* d_segs, int nsegs, int64_t total_chunks,
                       rpb200_stream_t);

/* Fused exchange over NVLink peer memory (replaces MPI_Irecv/Isend/Waitall,
 * HALO_EXCHANGE_FUSED-Cuda.cpp:109-196).  The pack kernel stores straight into the
 * receiving rank's unpack buffers through peer pointers (segment.buffer), then
 * publishes `epoch` to one flag per destination rank; the unpack kernel waits until
 * every source rank's flag in ITS OWN flag array reached `epoch`.
 *   d_peer_flags[r] : device pointer (peer-mapped) to rank r's flag array, entry [my_rank]
 *   d_my_flags      : this rank's flag array, one uint64 per rank
 *   d_src_ranks     : the nsrc distinct ranks this rank receives from                */
int rpb200_halo_pack_signal(rpb200_ctx*, const rpb200_halo_seg* d_segs, int nsegs,
                            int64_t total_chunks, uint64_t* const* d_peer_flags, int npeers,
                            uint64_t epoch, rpb200_stream_t);
int rpb200_halo_wait_unpack(rpb200_ctx*, const rpb200_halo_seg* d_segs, int nsegs,
                            int64_t total_chunks, const uint64_t* d_my_flags,
                            const int* d_src_ranks, int nsrc, uint64_t epoch, rpb200_stream_t);

/* CUDA IPC plumbing so one-process-per-GPU ranks can map each other's buffers
 * (what `--cuda-mpi-data-space CudaDevice` + CUDA-aware MPI would do underneath,
 * RunParams.hpp:358).  handle = 64 bytes (cudaIpcMemHandle_t).                      */
#define RPB200_IPC_HANDLE_BYTES 64
int rpb200_ipc_export(void* d_ptr, unsigned char handle[RPB200_IPC_HANDLE_BYTES]);
int rpb200_ipc_open(const unsigned char handle[RPB200_IPC_HANDLE_BYTES], void** d_ptr_out);
int rpb200_ipc_close(void* d_ptr);

/* ---- device memory + timing helpers for C/C++ hosts (the suite harness) ---------
 * Replace allocAndInitData, copyData and deallocData for DataSpace::CudaDevice
 * (DataUtils.hpp:196-409, CudaDataUtils.hpp:163-291).                               */
int rpb200_malloc(void** d_ptr, size_t bytes);
int rpb200_free(void* d_ptr);
int rpb200_malloc_host(void** h_ptr, size_t bytes);     /* pinned */
int rpb200_free_host(void* h_ptr);
int rpb200_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, rpb200_stream_t);
int rpb200_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, rpb200_stream_t);
int rpb200_memset(void* d_ptr, int value, size_t bytes, rpb200_stream_t);
int rpb200_stream_synchronize(rpb200_stream_t);
int rpb200_device_synchronize(void);
/* cudaEvent pair timing (north_star: "KernelBase timing via cudaEvents")            */
typedef struct rpb200_timer rpb200_timer;
int rpb200_timer_create(rpb200_timer** out);
int rpb200_timer_start(rpb200_timer*, rpb200_stream_t);
int rpb200_timer_stop(rpb200_timer*, rpb200_stream_t);
int rpb200_timer_elapsed_ms(rpb200_timer*, float* ms);   /* synchronises on the stop event */
void rpb200_timer_destroy(rpb200_timer*);

#ifdef __cplusplus
}
#endif
#endif /* RPB200_H */
