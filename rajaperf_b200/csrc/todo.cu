// todo.cu -- entry points not implemented yet: they fail loudly (cudaErrorNotSupported), never fall back.
#include "common.cuh"
#define NOT_YET return (int)cudaErrorNotSupported
extern "C" int rpb200_scan_exclusive(rpb200_ctx*, const double*, double*, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" size_t rpb200_sort_scratch_bytes(int64_t, int) { return 0; }
extern "C" int rpb200_sort_keys_f64(rpb200_ctx*, double*, int64_t, void*, size_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_sort_pairs_f64(rpb200_ctx*, double*, double*, int64_t, void*, size_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_mass3dpa(rpb200_ctx*, const double*, const double*, const double*, const double*, double*, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_diffusion3dpa(rpb200_ctx*, const double*, const double*, const double*, const double*, double*, int64_t, int, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_convection3dpa(rpb200_ctx*, const double*, const double*, const double*, const double*, const double*, double*, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_ltimes(rpb200_ctx*, double*, const double*, const double*, int64_t, int64_t, int64_t, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_halo_chunk(void) { return 0; }
extern "C" int rpb200_halo_pack(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_halo_unpack(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_halo_pack_signal(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, uint64_t* const*, int, uint64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_halo_wait_unpack(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, const uint64_t*, const int*, int, uint64_t, rpb200_stream_t) { NOT_YET; }
