/*
 * rpb_oracle.h -- CPU restatement of the RAJAPerf hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity checker for the Base_B200 variant.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load it.  Nothing under rajaperf_b200/ (the product) links, imports or calls it.
 *
 * Every function cites the reference file:line (relative to /root/reference/src)
 * whose behaviour it restates.  Plain C, scalar, single thread unless the name
 * ends in _omp.  Compiled with -ffp-contract=off so no FMA contraction happens,
 * like the reference's x86-64 Base_Seq build.
 *
 * Parity pin: tests/test_oracle_kat.py checks every orc_kat_* driver against the
 * known-answer checksums minted from the reference's own Base_Seq build
 * (tests/golden/ref_checksums.json; the MPI-only exchange
 * kernels: tests/golden/ref_checksums_mpi1.json, from the reference built against
 * oracle/mpi_stub and run on 1-8 ranks).
 */
#ifndef RPB_ORACLE_H
#define RPB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic inputs + checksum: common/DataUtils.cpp ------------------- */
void   orc_reset_init_count(void);                     /* DataUtils.cpp:131-139 */
int    orc_get_init_count(void);
void   orc_init_real(double* p, int64_t n);            /* DataUtils.cpp:504-513 */
void   orc_init_const(double* p, int64_t n, double v); /* DataUtils.cpp:518-525 */
void   orc_init_rand_value(double* p, int64_t n);      /* DataUtils.cpp:560-569 */
void   orc_init_rand_sign(double* p, int64_t n);       /* DataUtils.cpp:542-555 */
void   orc_init_int(int* p, int64_t n);                /* DataUtils.cpp:477-497 */
double orc_init_scalar(void);                          /* DataUtils.cpp:589-595 */
long double orc_checksum(const double* p, int64_t n, double scale); /* :600-621 */

/* ---- Stream group: stream/{COPY,MUL,ADD,TRIAD,DOT}.hpp bodies ------------ */
void   orc_stream_copy(double* c, const double* a, int64_t n);
void   orc_stream_mul(double* b, const double* c, double alpha, int64_t n);
void   orc_stream_add(double* c, const double* a, const double* b, int64_t n);
void   orc_stream_triad(double* a, const double* b, const double* c, double alpha, int64_t n);
double orc_stream_dot(const double* a, const double* b, int64_t n, double init);

/* ---- Algorithm group ------------------------------------------------------ */
double orc_reduce_sum(const double* x, int64_t n, double init);
void   orc_scan_exclusive(const double* x, double* y, int64_t n);
void   orc_sort(double* x, int64_t n);
void   orc_sort_pairs(double* keys, double* vals, int64_t n);

/* ---- Apps group ------------------------------------------------------------ */
void orc_mass3dpa(const double* B, const double* Bt, const double* D,
                  const double* X, double* Y, int64_t NE);
void orc_diffusion3dpa(const double* Basis, const double* dBasis, const double* D,
                       const double* X, double* Y, int64_t NE, int symmetric);
void orc_convection3dpa(const double* Basis, const double* tBasis, const double* dBasis,
                        const double* D, const double* X, double* Y, int64_t NE);
void orc_ltimes(double* phi, const double* ell, const double* psi,
                int64_t num_d, int64_t num_g, int64_t num_m, int64_t num_z);
/* effective DIFFUSION3DPA basis tables after the aliased fills (see .c) */
void orc_diffusion3dpa_tables(const double* Basis, const double* dBasis,
                              double* fill1 /*12*/, double* fill2 /*12*/);

/* ---- Comm group: comm/HALO_base.cpp ---------------------------------------- */
#define ORC_HALO_NEIGHBORS 26
extern const int orc_halo_offsets[ORC_HALO_NEIGHBORS][3];   /* HALO_base.cpp:82-116 */
void    orc_halo_grid_dims(int64_t target_size, int64_t dims[3]);  /* :31-35 */
int64_t orc_halo_extent_len(int is_recv, int l, int64_t halo_width, const int64_t dims[3]);
/* fills list (length orc_halo_extent_len) with flat indices, i fastest */
void    orc_halo_make_list(int is_recv, int l, int64_t halo_width, const int64_t dims[3], int* list);
/* neighbour ranks and tags for a rank in a periodic px*py*pz grid (:183-227,260) */
void    orc_halo_neighbors(int rank, const int pdims[3], int ranks[ORC_HALO_NEIGHBORS],
                           int send_tags[ORC_HALO_NEIGHBORS], int recv_tags[ORC_HALO_NEIGHBORS]);
/* buffer[i] = var[list[i]] / var[list[i]] = buffer[i]   (HALO_base.hpp:25-29) */
void    orc_halo_pack(double* buffer, const int* list, const double* var, int64_t len);
void    orc_halo_unpack(double* var, const int* list, const double* buffer, int64_t len);

/* ---- widened rows ----------------------------------------------- */
long double orc_checksum_int(const int* p, int64_t n, double scale);             /* DataUtils.cpp:623-629 */
int64_t orc_indexlist(const double* x, int* list, int64_t n);                     /* basic/INDEXLIST-Seq.cpp:40-52 */
int64_t orc_indexlist_3loop(const double* x, int* list, int64_t n);               /* basic/INDEXLIST_3LOOP-Seq.cpp:43-65 */
void    orc_polybench_gemm(const double* A, const double* B, double* C, int64_t ni, int64_t nj, int64_t nk,
                           double alpha, double beta);                            /* polybench/POLYBENCH_GEMM-Seq.cpp:37-47 */
void    orc_polybench_gemm_dims(int64_t target, int64_t* ni, int64_t* nj, int64_t* nk);  /* POLYBENCH_GEMM.cpp:24-35 */
long double orc_kat_halo_sendrecv(int64_t target_size, int reps, int halo_width, int num_vars, const int pdims[3]);
long double orc_kat_memcpy(int64_t target_size, int reps);
long double orc_kat_memset(int64_t target_size, int reps);
long double orc_kat_indexlist(int64_t target_size, int reps);
long double orc_kat_indexlist_3loop(int64_t target_size, int reps);
long double orc_kat_polybench_gemm(int64_t target_size, int reps);
/* ---- whole-kernel known-answer drivers: setUp -> reps -> checksum ---------- */
/* target_size = --size value (<=0: default size); reps = --checkrun N.        */
long double orc_kat_stream_copy(int64_t target_size, int reps);
long double orc_kat_stream_mul(int64_t target_size, int reps);
long double orc_kat_stream_add(int64_t target_size, int reps);
long double orc_kat_stream_triad(int64_t target_size, int reps);
long double orc_kat_stream_dot(int64_t target_size, int reps);
long double orc_kat_reduce_sum(int64_t target_size, int reps);
long double orc_kat_scan(int64_t target_size, int reps);
long double orc_kat_sort(int64_t target_size, int reps);
long double orc_kat_sortpairs(int64_t target_size, int reps);
long double orc_kat_mass3dpa(int64_t target_size, int reps);
long double orc_kat_diffusion3dpa(int64_t target_size, int reps);
long double orc_kat_convection3dpa(int64_t target_size, int reps);
long double orc_kat_ltimes(int64_t target_size, int reps, int num_d, int num_g, int num_m);
long double orc_kat_halo_packing_fused(int64_t target_size, int reps, int halo_width, int num_vars);
/* simulates px*py*pz ranks in one process; returns the rank-averaged checksum
 * the reference's report prints (Executor.cpp:1392-1467) and, if per_rank is
 * non-NULL, each rank's checksum. */
long double orc_kat_halo_exchange_fused(int64_t target_size, int reps, int halo_width,
                                        int num_vars, const int pdims[3],
                                        long double* per_rank);

/* pointer-out forms for ctypes (a returned long double is narrowed to double) */
void orc_checksum_int_out(const int* p, int64_t n, double scale, long double* out);
void orc_checksum_out(const double* p, int64_t n, double scale, long double* out);
int  orc_kat(const char* kernel_full_name, int64_t target_size, int reps,
             const int* iparams, long double* out);

/* ---- multi-threaded timing legs for bench.py's cpu_baseline ---------------- */
/* OpenMP restatements of stream/<K>-OMP.cpp (parallel for / reduction).         */
int    orc_omp_threads(void);
void   orc_stream_copy_omp(double* c, const double* a, int64_t n);
void   orc_stream_mul_omp(double* b, const double* c, double alpha, int64_t n);
void   orc_stream_add_omp(double* c, const double* a, const double* b, int64_t n);
void   orc_stream_triad_omp(double* a, const double* b, const double* c, double alpha, int64_t n);
double orc_stream_dot_omp(const double* a, const double* b, int64_t n, double init);
double orc_reduce_sum_omp(const double* x, int64_t n, double init);

#ifdef __cplusplus
}
#endif
#endif
