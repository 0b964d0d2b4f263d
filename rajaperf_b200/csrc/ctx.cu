// ctx.cu -- context, tunings, device-memory / timing / IPC helpers of the C ABI.
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>

static const char* const k_kernel_names[RPB_K_COUNT] = {
  "Stream_COPY", "Stream_MUL", "Stream_ADD", "Stream_TRIAD", "Stream_DOT",
  "Algorithm_REDUCE_SUM", "Algorithm_SCAN", "Algorithm_SORT", "Algorithm_SORTPAIRS",
  "Apps_MASS3DPA", "Apps_DIFFUSION3DPA", "Apps_CONVECTION3DPA", "Apps_LTIMES",
  "Comm_HALO_PACKING_FUSED", "Comm_HALO_EXCHANGE_FUSED",
  "Basic_INDEXLIST", "Polybench_GEMM"
};

// Built-in defaults; see profiles/ for the sweeps that picked them.
static void default_tunings(rpb200_ctx* c)
{
  for (int k = 0; k < RPB_K_COUNT; ++k) c->tune[k] = rpb_tuning{256, 8, 4};
  c->tune[RPB_K_COPY]       = rpb_tuning{512, 0, 2};
  c->tune[RPB_K_MUL]        = rpb_tuning{512, 0, 2};
  c->tune[RPB_K_ADD]        = rpb_tuning{512, 0, 2};
  c->tune[RPB_K_TRIAD]      = rpb_tuning{512, 0, 2};
  c->tune[RPB_K_DOT]        = rpb_tuning{256, 8, 2};
  c->tune[RPB_K_REDUCE_SUM] = rpb_tuning{256, 4, 8};
  c->tune[RPB_K_SCAN]       = rpb_tuning{512, 4, 4};
  c->tune[RPB_K_MASS3DPA]       = rpb_tuning{128, 0, 1};
  c->tune[RPB_K_DIFFUSION3DPA]  = rpb_tuning{128, 0, 1};
  c->tune[RPB_K_CONVECTION3DPA] = rpb_tuning{128, 0, 1};
  c->tune[RPB_K_INDEXLIST]      = rpb_tuning{512, 4, 4};
  // halo kernels.  HALO_PACKING_FUSED: ONE launch over the unit list (unroll 1: x-face units mixed in with the streaming
  // units; 3: first; 5: two phases; 2 / 4: the two-launch form without / with L2 hints), 2 CTAs per SM -- fewer resident CTAs
  // are FASTER here: 86 us at 2 per SM against 100 us at 4 per SM at 512^3, 309 against 407 us at 1024^3; the two-launch
  // form takes 96 / 491 us (profiles/r02_f/).  Two-launch forms: block_size 256 = contiguous chunk ranges, 192 = the same
  // with the PACK launches walking the work list backwards (the strided x faces are packed last, so the unpack -- which
  // walks forward and starts with the ghost cells sharing their L2 lines -- finds them resident), 128 = round-robin.
  // HALO_EXCHANGE_FUSED: unroll 2 = pack launch + unpack launch (default: 100-106 us at 512^3 on one rank against 104-122 us
  // for the one-launch form; at 1024^3 both forms take ~435 us at 2 CTAs per SM), 1 = one launch over the unit list;
  // ctas_per_sm 0 = automatic (csrc/halo.cu: 4 per SM, 2 for the two launches of a rep of more than 5000 chunks -- 1024^3).
  c->tune[RPB_K_HALO_PACKING_FUSED]  = rpb_tuning{192, 2, 1};
  c->tune[RPB_K_HALO_EXCHANGE_FUSED] = rpb_tuning{192, 0, 2};
}

extern "C" const char* rpb200_version(void) { return "rajaperf-b200 0.1 (sm_100a)"; }

// ---- per-stream scratch ---------------------------------------------------------------
namespace {
struct slot_lock {
  rpb200_ctx* c;
  explicit slot_lock(rpb200_ctx* ctx) : c(ctx) { while (__sync_lock_test_and_set(&c->lock, 1)) { } }
  ~slot_lock() { __sync_lock_release(&c->lock); }
};

constexpr size_t FIXED_PARTIALS = sizeof(double) * RPB_MAX_PARTIALS * 2;
constexpr size_t FIXED_BYTES = FIXED_PARTIALS + 256 /* tickets */ + 256 /* epochs */ + 512 /* basis tables */;

int slot_alloc(rpb_scratch* sc)
{
  if (sc->d_fixed) return 0;
  RPB_CHECK(cudaMalloc(&sc->d_fixed, FIXED_BYTES));
  cudaError_t e = cudaMemset(sc->d_fixed, 0, FIXED_BYTES);
  if (e != cudaSuccess) { cudaFree(sc->d_fixed); sc->d_fixed = nullptr; return (int)e; }
  char* p = (char*)sc->d_fixed;
  sc->d_partials = (double*)p;
  sc->d_ticket = (unsigned int*)(p + FIXED_PARTIALS);
  sc->d_scan_ticket = sc->d_ticket + 8;
  sc->d_epoch = (unsigned long long*)(p + FIXED_PARTIALS + 256);
  sc->d_basis_tables = (double*)(p + FIXED_PARTIALS + 512);
  return 0;
}
void slot_free(rpb_scratch* sc)
{
  cudaFree(sc->d_fixed);
  cudaFree(sc->d_scan_state);
  cudaFree(sc->d_ilist_state);
  memset(sc, 0, sizeof(*sc));
}
}  // namespace

// grows only; fresh memory is zeroed once: epoch 0 is never used, so fresh descriptors read "not ready"
int rpb_grow_state(void** d_state, size_t* bytes, size_t need, cudaStream_t st)
{
  if (need <= *bytes) return 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (st != nullptr && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
    return (int)cudaErrorStreamCaptureUnsupported;      // rpb200_scan_reserve / rpb200_indexlist_reserve before capturing
  RPB_CHECK(cudaStreamSynchronize(st));
  if (*d_state) RPB_CHECK(cudaFree(*d_state));
  *d_state = nullptr; *bytes = 0;
  const size_t cap_bytes = need + need / 2 + 4096;
  RPB_CHECK(cudaMalloc(d_state, cap_bytes));
  RPB_CHECK(cudaMemset(*d_state, 0, cap_bytes));
  *bytes = cap_bytes;
  return 0;
}

rpb_scratch* rpb_get_scratch(rpb200_ctx* ctx, cudaStream_t st, int* err)
{
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { *err = (int)e; return nullptr; }
  if (dev != ctx->device) { *err = RPB200_EDEVICE; return nullptr; }
  slot_lock guard(ctx);
  rpb_scratch* sc = &ctx->slot[ctx->last_slot];
  if (sc->attached && sc->stream == st) return sc;
  int free_ready = -1, free_any = -1;
  for (int i = 0; i < RPB_MAX_STREAMS; ++i) {
    rpb_scratch* c = &ctx->slot[i];
    if (c->attached) {
      if (c->stream == st) { ctx->last_slot = i; return c; }
    } else if (c->d_fixed) {
      if (free_ready < 0) free_ready = i;
    } else if (free_any < 0) {
      free_any = i;
    }
  }
  const int i = free_ready >= 0 ? free_ready : free_any;
  if (i < 0) { *err = RPB200_ENOSLOT; return nullptr; }
  sc = &ctx->slot[i];
  if (!sc->d_fixed) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (st != nullptr && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
      *err = (int)cudaErrorStreamCaptureUnsupported;
      return nullptr;
    }
    const int rc = slot_alloc(sc);
    if (rc != 0) { *err = rc; return nullptr; }
  }
  sc->stream = st;
  sc->attached = 1;
  ctx->last_slot = i;
  return sc;
}

extern "C" int rpb200_stream_attach(rpb200_ctx* ctx, rpb200_stream_t s)
{
  if (!ctx) return RPB200_EINVAL;
  RPB_SCRATCH(sc, ctx, rpb_stream(s));
  int rc = rpb_grow_state(&sc->d_scan_state, &sc->scan_state_bytes, ctx->scan_reserve_bytes, nullptr);
  if (rc == 0) rc = rpb_grow_state(&sc->d_ilist_state, &sc->ilist_state_bytes, ctx->ilist_reserve_bytes, nullptr);
  return rc;
}

extern "C" int rpb200_stream_detach(rpb200_ctx* ctx, rpb200_stream_t s)
{
  if (!ctx) return RPB200_EINVAL;
  cudaStream_t st = rpb_stream(s);
  slot_lock guard(ctx);
  for (int i = 0; i < RPB_MAX_STREAMS; ++i)
    if (ctx->slot[i].attached && ctx->slot[i].stream == st) {
      ctx->slot[i].attached = 0;      // memory is kept for the next stream that attaches; its state is re-armed
      ctx->slot[i].stream = nullptr;
      return 0;
    }
  return RPB200_EINVAL;
}

extern "C" int rpb200_create(int device, rpb200_ctx** out)
{
  if (!out) return RPB200_EINVAL;
  *out = nullptr;
  int ndev = 0;
  RPB_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return (int)cudaErrorInvalidDevice;
  int prev = -1;
  RPB_CHECK(cudaGetDevice(&prev));
  cudaDeviceProp prop;
  RPB_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    // sm_100a cubins only: there is no other code path to fall back to.
    return (int)cudaErrorNoKernelImageForDevice;
  }
  rpb200_ctx* c = (rpb200_ctx*)calloc(1, sizeof(rpb200_ctx));
  if (!c) return (int)cudaErrorMemoryAllocation;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  default_tunings(c);
  // the caller's current device is left as it was (a host holding contexts for several GPUs selects the device itself
  // before each call; every entry point checks it and returns RPB200_EDEVICE on a mismatch)
  cudaError_t e = cudaSetDevice(device);
  for (int i = 0; e == cudaSuccess && i < RPB_PREALLOC_STREAMS; ++i) e = (cudaError_t)slot_alloc(&c->slot[i]);
  cudaSetDevice(prev);
  if (e != cudaSuccess) { rpb200_destroy(c); return (int)e; }
  *out = c;
  return 0;
}

extern "C" void rpb200_destroy(rpb200_ctx* c)
{
  if (!c) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  for (int i = 0; i < RPB_MAX_STREAMS; ++i) slot_free(&c->slot[i]);
  if (prev >= 0) cudaSetDevice(prev);
  free(c);
}

extern "C" const char* rpb200_error_string(int err)
{
  if (err == 0) return "success";
  if (err == RPB200_EINVAL) return "rpb200: invalid argument";
  if (err == RPB200_ETIMEDOUT) return "rpb200: timed out waiting for a halo message";
  if (err == RPB200_EDEVICE) return "rpb200: the current CUDA device is not the context's device";
  if (err == RPB200_ENOSLOT) return "rpb200: too many streams attached to one context (rpb200_stream_detach)";
  return cudaGetErrorString((cudaError_t)err);
}

extern "C" int rpb200_sm_count(const rpb200_ctx* c) { return c ? c->sm_count : 0; }

extern "C" int rpb200_set_tuning(rpb200_ctx* c, const char* kernel, int block_size,
                                 int ctas_per_sm, int unroll)
{
  if (!c || !kernel) return RPB200_EINVAL;
  for (int k = 0; k < RPB_K_COUNT; ++k) {
    if (strcmp(kernel, k_kernel_names[k]) == 0) {
      if (block_size > 0) {
        if (block_size % 32 != 0 || block_size > 1024) return RPB200_EINVAL;
        c->tune[k].block_size = block_size;
      }
      if (ctas_per_sm >= 0) c->tune[k].ctas_per_sm = ctas_per_sm;
      if (unroll > 0) c->tune[k].unroll = unroll;
      return 0;
    }
  }
  return RPB200_EINVAL;
}

extern "C" int rpb200_get_tuning(const rpb200_ctx* c, const char* kernel, int* block_size, int* ctas_per_sm, int* unroll)
{
  if (!c || !kernel) return RPB200_EINVAL;
  for (int k = 0; k < RPB_K_COUNT; ++k)
    if (strcmp(kernel, k_kernel_names[k]) == 0) {
      if (block_size) *block_size = c->tune[k].block_size;
      if (ctas_per_sm) *ctas_per_sm = c->tune[k].ctas_per_sm;
      if (unroll) *unroll = c->tune[k].unroll;
      return 0;
    }
  return RPB200_EINVAL;
}

extern "C" int rpb200_reset_tuning(rpb200_ctx* c, const char* kernel)
{
  if (!c) return RPB200_EINVAL;
  rpb200_ctx d;                       // only the tuning table of this scratch copy is used
  default_tunings(&d);
  if (!kernel) { memcpy(c->tune, d.tune, sizeof(c->tune)); return 0; }
  for (int k = 0; k < RPB_K_COUNT; ++k)
    if (strcmp(kernel, k_kernel_names[k]) == 0) { c->tune[k] = d.tune[k]; return 0; }
  return RPB200_EINVAL;
}

// ---- memory helpers ----------------------------------------------------------------
extern "C" int rpb200_malloc(void** p, size_t bytes) { RPB_CHECK(cudaMalloc(p, bytes)); return 0; }
extern "C" int rpb200_free(void* p) { RPB_CHECK(cudaFree(p)); return 0; }
extern "C" int rpb200_malloc_host(void** p, size_t bytes) { RPB_CHECK(cudaMallocHost(p, bytes)); return 0; }
extern "C" int rpb200_free_host(void* p) { RPB_CHECK(cudaFreeHost(p)); return 0; }
extern "C" int rpb200_memcpy_h2d(void* d, const void* h, size_t bytes, rpb200_stream_t s)
{ RPB_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, rpb_stream(s))); return 0; }
extern "C" int rpb200_memcpy_d2h(void* h, const void* d, size_t bytes, rpb200_stream_t s)
{ RPB_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, rpb_stream(s))); return 0; }
extern "C" int rpb200_memset(void* d, int v, size_t bytes, rpb200_stream_t s)
{ RPB_CHECK(cudaMemsetAsync(d, v, bytes, rpb_stream(s))); return 0; }
extern "C" int rpb200_stream_synchronize(rpb200_stream_t s)
{ RPB_CHECK(cudaStreamSynchronize(rpb_stream(s))); return 0; }
extern "C" int rpb200_device_synchronize(void) { RPB_CHECK(cudaDeviceSynchronize()); return 0; }

// ---- cudaEvent timer ---------------------------------------------------------------
struct rpb200_timer { cudaEvent_t start, stop; };

extern "C" int rpb200_timer_create(rpb200_timer** out)
{
  if (!out) return RPB200_EINVAL;
  rpb200_timer* t = (rpb200_timer*)calloc(1, sizeof(rpb200_timer));
  if (!t) return (int)cudaErrorMemoryAllocation;
  cudaError_t e;
  if ((e = cudaEventCreate(&t->start)) != cudaSuccess) { free(t); return (int)e; }
  if ((e = cudaEventCreate(&t->stop)) != cudaSuccess) { cudaEventDestroy(t->start); free(t); return (int)e; }
  *out = t;
  return 0;
}
extern "C" int rpb200_timer_start(rpb200_timer* t, rpb200_stream_t s)
{ RPB_CHECK(cudaEventRecord(t->start, rpb_stream(s))); return 0; }
extern "C" int rpb200_timer_stop(rpb200_timer* t, rpb200_stream_t s)
{ RPB_CHECK(cudaEventRecord(t->stop, rpb_stream(s))); return 0; }
extern "C" int rpb200_timer_elapsed_ms(rpb200_timer* t, float* ms)
{
  RPB_CHECK(cudaEventSynchronize(t->stop));
  RPB_CHECK(cudaEventElapsedTime(ms, t->start, t->stop));
  return 0;
}
extern "C" void rpb200_timer_destroy(rpb200_timer* t)
{
  if (!t) return;
  cudaEventDestroy(t->start);
  cudaEventDestroy(t->stop);
  free(t);
}

// ---- CUDA IPC ----------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == RPB200_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int rpb200_ipc_export(void* d_ptr, unsigned char handle[RPB200_IPC_HANDLE_BYTES])
{
  cudaIpcMemHandle_t h;
  RPB_CHECK(cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(handle, &h, sizeof(h));
  return 0;
}
extern "C" int rpb200_ipc_open(const unsigned char handle[RPB200_IPC_HANDLE_BYTES], void** out)
{
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  RPB_CHECK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int rpb200_ipc_close(void* d_ptr) { RPB_CHECK(cudaIpcCloseMemHandle(d_ptr)); return 0; }

extern "C" int rpb200_enable_peer_access(int device, int peer_device)
{
  if (device == peer_device) return 0;
  int can = 0;
  RPB_CHECK(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) return (int)cudaErrorPeerAccessUnsupported;
  int cur = 0;
  RPB_CHECK(cudaGetDevice(&cur));
  RPB_CHECK(cudaSetDevice(device));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  cudaSetDevice(cur);
  return (int)e;
}
