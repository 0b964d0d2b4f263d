// todo.cu -- entry points not implemented yet: they fail loudly (cudaErrorNotSupported), never fall back.
#include "common.cuh"
#define NOT_YET return (int)cudaErrorNotSupported
extern "C" int rpb200_halo_chunk(void) { return 0; }
extern "C" int rpb200_halo_pack(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_halo_unpack(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_halo_pack_signal(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, uint64_t* const*, int, uint64_t, rpb200_stream_t) { NOT_YET; }
extern "C" int rpb200_halo_wait_unpack(rpb200_ctx*, const rpb200_halo_seg*, int, int64_t, const uint64_t*, const int*, int, uint64_t, rpb200_stream_t) { NOT_YET; }
