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
 *   - STREAMS: one context serves any number of streams CONCURRENTLY.  Everything a kernel writes
 *     besides the caller's arrays (reduction partials and tickets, scan / index-list look-back
 *     descriptors, tickets and epochs, PA basis tables) exists once per (context, stream) -- the
 *     reference likewise gives every reducer its own scratch, GPUUtils.hpp:250-330 -- and is
 *     attached to a stream at its first call (no allocation for the first 8 streams; see
 *     rpb200_stream_attach).  The three Apps_*3DPA kernels keep their basis tables in
 *     __constant__ memory, one set per DEVICE: calls on different streams are safe but
 *     serialise (an event orders each call behind the device's previous PA call).
 *     Caller-provided scratch (SORT) and halo plans are the caller's to keep to one stream at
 *     a time.  Look-back epochs live in device memory, so every entry point may be captured
 *     into a CUDA graph and the graph replayed any number of times;
 *   - DEVICE: a context belongs to the device it was created for; the calling thread's
 *     current device must be that device at every call (checked: RPB200_EDEVICE).
 *     rpb200_create() leaves the caller's current device unchanged;
 *   - return 0 on success, otherwise a cudaError_t value (or one of the RPB200_E* codes);
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
#define RPB200_ETIMEDOUT (-110)
#define RPB200_EDEVICE (-19)            /* the current CUDA device is not the context's            */
#define RPB200_ENOSLOT (-24)            /* more than 64 streams attached to one context            */

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
/* Optional.  attach: give `stream` its scratch set NOW (allocating one if the 8 pre-allocated sets are taken, and sizing
 * its look-back state to the largest rpb200_scan_reserve / rpb200_indexlist_reserve so far; synchronises) -- required only
 * before CAPTURING calls on a stream that is the 9th or later stream of the context, since a capture may not allocate.
 * detach: the stream is about to be destroyed; its set is kept for the next stream that attaches (at most 64 attached
 * streams per context: RPB200_ENOSLOT).                                                                               */
int rpb200_stream_attach(rpb200_ctx* ctx, rpb200_stream_t stream);
int rpb200_stream_detach(rpb200_ctx* ctx, rpb200_stream_t stream);

/* Launch tuning, the analogue of the reference's block-size tunings
 * (GPUUtils.hpp:345-373).  kernel = full kernel name ("Stream_TRIAD");
 * <=0 keeps the current value of that field (ctas_per_sm: < 0).  Meaning of the fields:
 *   Stream_*, Stream_DOT, Algorithm_REDUCE_SUM, Algorithm_SCAN (small n), Basic_INDEXLIST (small n):
 *       threads per CTA, persistent CTAs per SM (0 = one tile per CTA), independent vectors per thread;
 *   Comm_HALO_PACKING_FUSED / Comm_HALO_EXCHANGE_FUSED, two-launch forms: block_size 256 = contiguous chunk ranges per CTA,
 *       192 = the same with pack launches walking the list backwards (default), 128 = round-robin; ctas_per_sm; unroll 4 = L2
 *       eviction-priority hints on.  One-launch forms (rpb200_halo_plan_pack_unpack, rpb200_halo_exchange): ctas_per_sm;
 *       exchange unroll 1 = ONE launch per rep over the unit list (every pack unit, signal, wait + every unpack unit),
 *       3 = ONE launch with pack and unpack units on one ticket and messages signalled unit by unit (the unpack of early
 *       messages overlaps the packing of late ones), 2 / 4 = pack launch + unpack launch (default 2); exchange ctas_per_sm
 *       0 (default) = automatic: 4, or 2 for the two launches of a rep of more than 5000 chunks (1024^3 cells per rank);
 *       the one-launch forms need every CTA of a rank resident while its peers pack: one rank per GPU;
 *   Algorithm_SORT / Algorithm_SORTPAIRS: unroll 8 = digit histograms in shared bins instead of the lane-private 16-bit
 *       counters (default since round 2: profiles/r02_a_optin.log); 7 = no pass tests its tiles for uniformity;
 *       ctas_per_sm 1 = the look-back reads one tile descriptor at a time instead of four;
 *   Apps_MASS3DPA / Apps_CONVECTION3DPA: unroll selects a launch shape (csrc/pa.cu; 1 = default);
 *   Apps_LTIMES: ctas_per_sm; unroll 5..8 = psi staged through a bulk-async ring, 10 = row-chunk A fragments, else line-major (default);
 *   Polybench_GEMM: block_size 64 / 96 / 128 / 160 = CTA tiling (else automatic), unroll 8 = 32-deep stages.
 * Every setting computes the same result; the defaults are the measured best (profiles/).            */
int rpb200_set_tuning(rpb200_ctx* ctx, const char* kernel, int block_size,
                      int ctas_per_sm, int unroll);
/* The current launch shape of `kernel` (what an ncu capture must record to be comparable with a later run).            */
int rpb200_get_tuning(const rpb200_ctx* ctx, const char* kernel, int* block_size, int* ctas_per_sm, int* unroll);
/* Back to the built-in (measured best) launch shape of `kernel`; NULL = of every kernel.  The suite harness brackets
 * each non-default tuning of a kernel with set / reset (KernelBase::execute).                                          */
int rpb200_reset_tuning(rpb200_ctx* ctx, const char* kernel);

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
/* Optional: grow the look-back state of every scratch set of the context for n elements now (synchronises), so that
 * later rpb200_scan_exclusive calls never allocate -- required before capturing them into a CUDA graph.  The look-back
 * epoch is read from device memory by the kernel, so a captured scan replays correctly any number of times.  */
int rpb200_scan_reserve(rpb200_ctx*, int64_t n);
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

/* ---- widened rows -----------------------------------------------------
 * basic/INDEXLIST-Cuda.cpp:33-265 and basic/INDEXLIST_3LOOP-Cuda.cpp:60-128: both define
 *   list[0..len) = ascending { i in [0,n) : x[i] < 0.0 },  *d_len = len   (INDEXLIST.hpp:17-25,
 * INDEXLIST_3LOOP.hpp:17-24); one single-pass compaction serves both, without the 3LOOP
 * `counts` temporary.  list: Int_type (int), d_len: device Index_type (int64).  Entries of
 * list at and beyond len are left untouched (the reference checksums them too).       */
int rpb200_indexlist(rpb200_ctx*, const double* x, int* list, int64_t n, int64_t* d_len,
                     rpb200_stream_t);
/* Optional: grow the look-back state for n elements now (synchronises); required before
 * capturing rpb200_indexlist into a CUDA graph.                                        */
int rpb200_indexlist_reserve(rpb200_ctx*, int64_t n);
/* algorithm/MEMSET-Cuda.cpp:27-76: x[i] = val, the write-only calibration stream; Algorithm_MEMCPY
 * (MEMCPY-Cuda.cpp:27-76, y[i] = x[i]) is rpb200_stream_copy.                              */
int rpb200_memset_f64(rpb200_ctx*, double* x, double val, int64_t n, rpb200_stream_t);
/* polybench/POLYBENCH_GEMM-Cuda.cpp:44-85: C[i][j] = sum_k alpha * A[i][k] * B[k][j], row-major
 * A (ni x nk), B (nk x nj), C (ni x nj).  `beta` is dead in the reference body
 * (POLYBENCH_GEMM.hpp:32-39: "C *= beta" is overwritten by "C = dot") and is ignored here too. */
int rpb200_polybench_gemm(rpb200_ctx*, const double* A, const double* B, double* C,
                          int64_t ni, int64_t nj, int64_t nk, double alpha, double beta,
                          rpb200_stream_t);

/* ---- Comm group ------------------------------------------------------------------
 * comm/HALO_PACKING_FUSED-Cuda.cpp:52-197, HALO_EXCHANGE_FUSED-Cuda.cpp:52-206.
 *
 * (1) Generic fused pack / unpack over the reference's (buffer, list, var, len) tuples
 *     (HALO_PACKING_FUSED-Cuda.cpp:24-40 keeps them in pinned host arrays the kernel reads
 *     over PCIe and refills every rep).  Here a WORK LIST holds them in device memory next
 *     to a chunk -> tuple map, so one launch covers every tuple with no search and no idle
 *     CTA:  pack:  buffer[i] = var[list[i]]      unpack:  var[list[i]] = buffer[i].       */
typedef struct rpb200_halo_seg {
  double*    buffer;   /* contiguous message segment (device; may be a peer-GPU pointer)  */
  const int* list;     /* index list (Int_type, HALO_base.cpp:197-254), device            */
  double*    var;      /* the grid variable, device                                       */
  int64_t    len;      /* elements in this segment                                        */
  int32_t    msg;      /* message (neighbour) ordinal 0..25 this segment belongs to       */
  int32_t    flags;    /* filled by the library (alignment classes); pass 0               */
} rpb200_halo_seg;

typedef struct rpb200_halo_worklist rpb200_halo_worklist;
int     rpb200_halo_chunk(void);          /* elements one CTA moves                       */
/* h_segs: HOST array of nsegs tuples (copied).  update() re-uploads tuples of the same
 * lengths (stream-ordered; what the reference does every rep).                            */
int rpb200_halo_worklist_create(rpb200_ctx*, const rpb200_halo_seg* h_segs, int nsegs,
                                rpb200_halo_worklist** out);
int rpb200_halo_worklist_update(rpb200_halo_worklist*, const rpb200_halo_seg* h_segs, int nsegs,
                                rpb200_stream_t);
void rpb200_halo_worklist_destroy(rpb200_halo_worklist*);
int rpb200_halo_pack  (rpb200_ctx*, const rpb200_halo_worklist*, rpb200_stream_t);
int rpb200_halo_unpack(rpb200_ctx*, const rpb200_halo_worklist*, rpb200_stream_t);
/* pack(`pack`) and unpack(`unpack`) of one rep in ONE launch, for callers whose two lists touch DISJOINT memory -- true of
 * HALO_PACKING_FUSED (HALO_PACKING_FUSED-Seq.cpp:43-61 reads owned cells into the pack buffers, :71-97 writes ghost cells
 * from the unpack buffers), where the result then equals pack followed by unpack.  The first call for a (pack, unpack) pair
 * builds and uploads the merged item list (synchronises: make it before capturing into a graph); see
 * rpb200_halo_plan_pack_unpack below for what the item order buys.                                                      */
int rpb200_halo_pack_unpack(rpb200_ctx*, rpb200_halo_worklist* pack, rpb200_halo_worklist* unpack, rpb200_stream_t);
/* Test hook, host-only (no device is touched): the unit list the one-launch kernels walk for tuples of the given geometry
 * (lengths, strided flags, message ordinals, variable ids).  order 1 / 3 / 5 = the HALO_PACKING_FUSED orders (x units mixed
 * in / first / two phases), 0 = the exchange order, 6 = the progressive exchange order.  items_out: (tuple index | 1 << 30 for the unpack side, chunk) pairs;
 * unit_first_out: n_units + 1 entries.  Lets the CPU test-suite check that every (tuple, chunk) is moved exactly once.   */
int rpb200_debug_halo_units(const int64_t* pack_len, const int* pack_strided, const int* pack_msg, const int* pack_var, int npack,
                            const int64_t* unpack_len, const int* unpack_strided, const int* unpack_msg, const int* unpack_var,
                            int nunpack, int order, int* items_out, int max_items, int* unit_first_out, int max_units,
                            int* n_items, int* n_units, int* n_pack_units);

/* (2) HALO_base: the 26-neighbour periodic decomposition and its index lists
 *     (comm/HALO_base.cpp:31-35 grid dims, :82-116 offsets, :118-166 extents, :169-291
 *     create_lists).  A plan owns the 52 device index lists of one rank.                  */
#define RPB200_HALO_NEIGHBORS 26
typedef struct rpb200_halo_plan rpb200_halo_plan;
void rpb200_halo_grid_dims(int64_t target_problem_size, int64_t dims[3]);
int  rpb200_halo_plan_create(rpb200_ctx*, const int64_t grid_dims[3], int64_t halo_width,
                             int num_vars, int my_rank, const int rank_dims[3],
                             rpb200_halo_plan** out);
void rpb200_halo_plan_destroy(rpb200_halo_plan*);
int64_t rpb200_halo_plan_var_size(const rpb200_halo_plan*);   /* (nx+2h)(ny+2h)(nz+2h)     */
/* neighbour l: its rank, the send/recv tags, list lengths and DEVICE list pointers        */
int  rpb200_halo_plan_neighbor(const rpb200_halo_plan*, int l, int* rank, int* send_tag,
                               int* recv_tag, int64_t* pack_len, int64_t* unpack_len,
                               const int** d_pack_list, const int** d_unpack_list);
/* HALO_PACKING_FUSED: bind the caller's vars[num_vars], pack_buffers[26], unpack_buffers[26]
 * (device pointers; buffer l holds num_vars segments of pack_len/unpack_len, variable-minor,
 * HALO_PACKING_FUSED-Seq.cpp:43-61, 71-97), then run the two fused launches of one rep.    */
int  rpb200_halo_plan_bind(rpb200_halo_plan*, double* const* vars, double* const* pack_buffers,
                           double* const* unpack_buffers);
int  rpb200_halo_plan_pack(rpb200_halo_plan*, rpb200_stream_t);
int  rpb200_halo_plan_unpack(rpb200_halo_plan*, rpb200_stream_t);
/* One rep of HALO_PACKING_FUSED in ONE launch.  The pack (HALO_PACKING_FUSED-Seq.cpp:43-61: owned cells -> pack buffers) and
 * the unpack (:71-97: unpack buffers -> ghost cells) touch disjoint cells and disjoint buffers, so the result does not
 * depend on how their work interleaves; the launch walks a list of UNITS dealt round-robin to the CTAs, where chunk c of
 * pack(-x), pack(+x), unpack(-x), unpack(+x) of a variable is ONE unit (one CTA, back to back) -- those four touch the same
 * one or two 32-byte sectors per grid row (csrc/halo.cu: halo_items_kernel).  Same result as pack then unpack.
 * Tuning Comm_HALO_PACKING_FUSED `unroll`: 2 = run the two launches instead; 3 = x-face units first instead of mixed in;
 * 5 = two phases in the one launch (every pack unit, then every unpack unit, x faces last / first).                      */
int  rpb200_halo_plan_pack_unpack(rpb200_halo_plan*, rpb200_stream_t);

/* (3) HALO_EXCHANGE_FUSED over NVLink peer memory (replaces MPI_Irecv / MPI_Isend /
 *     MPI_Waitall on host-pinned buffers, HALO_EXCHANGE_FUSED-Cuda.cpp:109-196).
 *     Every rank owns a WINDOW in device memory: 26 arrival flags + two generations of its 26
 *     receive buffers.  The pack kernel of rank r stores message l straight into the window
 *     of rank ranks[l] (receive slot = the opposite neighbour, i.e. the slot whose recv_tag is
 *     l, HALO_base.cpp:260) and, when the last chunk of a message has been written, releases
 *     that slot's flag with the exchange epoch.  The unpack kernel acquires each slot's flag
 *     before scattering it.  Two buffer generations + the symmetric neighbour relation make a
 *     separate "buffer free" acknowledgement unnecessary (Comm).
 *     Set-up: window() -> exchange the 64-byte IPC handles between ranks by any means (MPI,
 *     torch.distributed, a file) -> connect().  connect_ptrs() is the same for ranks that
 *     live in one process (plain device pointers, e.g. several ranks simulated on one GPU). */
int  rpb200_halo_exchange_window(rpb200_halo_plan*, double* const* vars, void** d_window,
                                 size_t* bytes, unsigned char ipc_handle[64]);
int  rpb200_halo_exchange_connect(rpb200_halo_plan*, int nranks, const unsigned char* ipc_handles);
int  rpb200_halo_exchange_connect_ptrs(rpb200_halo_plan*, int nranks, void* const* d_windows);
/* one rep = pack_signal then wait_unpack (or exchange() for both); all stream-ordered      */
int  rpb200_halo_exchange_pack(rpb200_halo_plan*, rpb200_stream_t);
int  rpb200_halo_exchange_unpack(rpb200_halo_plan*, rpb200_stream_t);
/* The unfused HALO_EXCHANGE (comm/HALO_EXCHANGE-Cuda.cpp:26-123, HALO_EXCHANGE-Seq.cpp:34-116): one launch per
 * (neighbour l, variable v) tuple instead of one per rep.  Message l is released to its receiver by whichever pack
 * launch completes its last chunk; `commit` != 0 marks the LAST unpack launch of the rep (it advances the epoch).   */
int  rpb200_halo_exchange_pack_seg(rpb200_halo_plan*, int l, int v, rpb200_stream_t);
int  rpb200_halo_exchange_unpack_seg(rpb200_halo_plan*, int l, int v, int commit, rpb200_stream_t);
/* HALO_SENDRECV (comm/HALO_SENDRECV.cpp:21-123, HALO_SENDRECV-Seq.cpp:34-52): transport only.  Message l is the caller's
 * send_buffers[l] (num_vars * pack_len[l] doubles, device); rpb200_halo_sendrecv puts every message into the receive slot
 * of its destination (the slot whose recv_tag is l) over NVLink, signals it, and waits for this rank's 26 messages:
 * MPI_Irecv x26 / MPI_Isend x26 / MPI_Waitall x2 as one put kernel + one wait kernel, no host synchronisation.
 * bind() after rpb200_halo_exchange_connect*.  rpb200_halo_recv_buffer returns where message l of the last completed rep
 * lies in this rank's window (it synchronises; the location alternates between two generations).                  */
int  rpb200_halo_sendrecv_bind(rpb200_halo_plan*, double* const* d_send_buffers);
int  rpb200_halo_sendrecv(rpb200_halo_plan*, rpb200_stream_t);            /* = put, then wait */
/* the two halves, for hosts that drive several ranks through ONE stream (all puts must be queued before a wait spins) */
int  rpb200_halo_sendrecv_put(rpb200_halo_plan*, rpb200_stream_t);
int  rpb200_halo_sendrecv_wait(rpb200_halo_plan*, rpb200_stream_t);
int  rpb200_halo_recv_buffer(rpb200_halo_plan*, int l, const double** d_ptr, int64_t* len);
int  rpb200_halo_exchange(rpb200_halo_plan*, rpb200_stream_t);
/* 0, or RPB200_ETIMEDOUT if an unpack CTA gave up waiting for a message (synchronises).  The wait is bounded in wall-clock
 * time (%globaltimer; 2 s, RPB200_HALO_TIMEOUT_MS overrides); a message that did not arrive is NOT unpacked and the rep's
 * epoch is NOT committed: the caller must treat a non-zero status as fatal for the plan (the suite stubs abort).        */
int  rpb200_halo_exchange_status(rpb200_halo_plan*);

/* CUDA IPC plumbing so one-process-per-GPU ranks can map each other's buffers
 * (what `--cuda-mpi-data-space CudaDevice` + CUDA-aware MPI would do underneath,
 * RunParams.hpp:358).  handle = 64 bytes (cudaIpcMemHandle_t).                      */
#define RPB200_IPC_HANDLE_BYTES 64
int rpb200_ipc_export(void* d_ptr, unsigned char handle[RPB200_IPC_HANDLE_BYTES]);
int rpb200_ipc_open(const unsigned char handle[RPB200_IPC_HANDLE_BYTES], void** d_ptr_out);
int rpb200_ipc_close(void* d_ptr);
/* one process driving several GPUs: let `device` read/write `peer_device` memory through plain pointers */
int rpb200_enable_peer_access(int device, int peer_device);

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
