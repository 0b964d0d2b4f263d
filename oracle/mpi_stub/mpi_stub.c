/* mpi_stub.c -- see mpi.h: test infrastructure only.
 *
 * Two transports behind the same few calls:
 *   - one rank (default): everything in process, a message is a memcpy between a send and a receive of the same tag;
 *   - P ranks (RPB_MPI_SIZE = P > 1, RPB_MPI_RANK, RPB_MPI_SHM = a file under /dev/shm created zero-filled by the
 *     launcher, tests/golden/make_golden.py): P processes of the same binary share one arena -- a mutex + condition
 *     variable, a log of buffered messages {src, dst, tag, payload} in send order (so matching is MPI's: by source and tag,
 *     non-overtaking), a barrier, and one 4 KiB slot per rank for the collectives.  Sends are buffered and complete at
 *     once; a receive completes in Wait* when its message is in the log; 60 s without progress aborts (deadlock).
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

typedef struct {
  int kind;        /* 0 free, 1 receive, 2 send */
  int done;        /* receive: a message was delivered; send: always (buffered)                */
  int tag;
  void* buf;       /* receive: destination                                                     */
  size_t bytes;
  void* data;      /* send: private copy of the payload while no receive has been posted yet   */
  int delivered;   /* send: payload already copied into a receive                              */
  int released;    /* send: its request was waited on before a receive took the payload        */
  int peer;        /* source of a receive / destination of a send                              */
} slot_t;

static slot_t* g_slots = 0;
static int g_nslots = 0;

static void die(const char* what)
{
  fprintf(stderr, "mpi_stub: %s\n", what);
  abort();
}

static size_t type_bytes(MPI_Datatype t)
{
  switch (t) {
    case MPI_BYTE: case MPI_CHAR: return 1;
    case MPI_INT: case MPI_FLOAT: return 4;
    case MPI_LONG: case MPI_LONG_LONG: case MPI_UNSIGNED_LONG: case MPI_DOUBLE: return 8;
    case MPI_LONG_DOUBLE: return sizeof(long double);
    default: die("unknown datatype"); return 0;
  }
}

static int new_slot(void)
{
  for (int i = 0; i < g_nslots; ++i) if (g_slots[i].kind == 0) return i;
  const int n = g_nslots ? 2 * g_nslots : 64;
  g_slots = (slot_t*)realloc(g_slots, sizeof(slot_t) * (size_t)n);
  if (!g_slots) die("out of memory");
  memset(g_slots + g_nslots, 0, sizeof(slot_t) * (size_t)(n - g_nslots));
  const int i = g_nslots;
  g_nslots = n;
  return i;
}

/* ---- the P-rank transport ------------------------------------------------------------------------------------ */
#define SHM_MAX_RANKS 64
#define SHM_SLOT_BYTES 4096
#define SHM_MAX_MSGS (1 << 20)

typedef struct {
  int src, dst, tag, taken;
  size_t bytes, offset;          /* payload at arena + offset */
} shm_msg_t;

typedef struct {
  volatile int ready;            /* rank 0 initialised the mutex and the condition variable */
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int bar_count, bar_gen;
  /* the message log and the payload heap are RINGS: first_untaken .. nmsgs are monotonically increasing message numbers
   * (slot = number % SHM_MAX_MSGS); payloads are allocated at heap_top, wrapping to heap_base, and freed in log order as
   * the oldest messages are taken.  A sender that finds no room waits (ranks that run ahead of their peers -- GPU variants
   * on 8 ranks -- once exhausted a bump allocator that was only rewound when NOTHING was in flight: round 2).            */
  size_t nmsgs, first_untaken;
  size_t heap_top, heap_end, heap_base;
  char coll[SHM_MAX_RANKS][SHM_SLOT_BYTES];
  shm_msg_t msgs[SHM_MAX_MSGS];
} shm_t;

static int g_size = 1, g_rank = 0;
static shm_t* g_shm = 0;
static char* g_arena = 0;

static void shm_lock(void) { if (pthread_mutex_lock(&g_shm->mu) != 0) die("mutex"); }
static void shm_unlock(void) { pthread_mutex_unlock(&g_shm->mu); }
static void shm_wait(void)       /* mutex held */
{
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  ts.tv_sec += 60;
  if (pthread_cond_timedwait(&g_shm->cv, &g_shm->mu, &ts) == ETIMEDOUT) die("no progress for 60 s: deadlock or a dead rank");
}

static void shm_barrier(void)
{
  shm_lock();
  const int gen = g_shm->bar_gen;
  if (++g_shm->bar_count == g_size) { g_shm->bar_count = 0; g_shm->bar_gen = gen + 1; pthread_cond_broadcast(&g_shm->cv); }
  else while (g_shm->bar_gen == gen) shm_wait();
  shm_unlock();
}

#define SHM_MSG(n) (&g_shm->msgs[(n) % SHM_MAX_MSGS])

/* mutex held: the oldest untaken message src -> me with this tag (its message number), or -1 */
static long shm_find(int src, int tag)
{
  for (size_t i = g_shm->first_untaken; i < g_shm->nmsgs; ++i) {
    const shm_msg_t* m = SHM_MSG(i);
    if (!m->taken && m->dst == g_rank && m->src == src && m->tag == tag) return (long)i;
  }
  return -1;
}

/* mutex held: copy message i out and retire it; the log and the heap are freed from their old end */
static void shm_take(long i, void* buf, size_t bytes)
{
  shm_msg_t* m = SHM_MSG((size_t)i);
  if (m->bytes > bytes) die("message longer than the posted receive");
  memcpy(buf, g_arena + m->offset, m->bytes);
  m->taken = 1;
  const size_t before = g_shm->first_untaken;
  while (g_shm->first_untaken < g_shm->nmsgs && SHM_MSG(g_shm->first_untaken)->taken) ++g_shm->first_untaken;
  if (g_shm->first_untaken == g_shm->nmsgs) g_shm->heap_top = g_shm->heap_base;       /* nothing in flight */
  if (g_shm->first_untaken != before) pthread_cond_broadcast(&g_shm->cv);             /* room for a waiting sender */
}

/* mutex held: room for `need` payload bytes in the ring (the live region runs from the oldest untaken message's payload to
 * heap_top, possibly wrapped); returns the offset or 0 */
static size_t shm_alloc(size_t need)
{
  if (g_shm->first_untaken == g_shm->nmsgs) {                 /* empty */
    if (g_shm->heap_base + need > g_shm->heap_end) die("one message is larger than the arena (--arena-mb)");
    g_shm->heap_top = g_shm->heap_base + need;
    return g_shm->heap_base;
  }
  const size_t head = SHM_MSG(g_shm->first_untaken)->offset;
  size_t top = g_shm->heap_top;
  if (top >= head) {                                          /* not wrapped: [head, top) is live */
    if (top + need <= g_shm->heap_end) { g_shm->heap_top = top + need; return top; }
    if (g_shm->heap_base + need <= head && head > g_shm->heap_base) { g_shm->heap_top = g_shm->heap_base + need; return g_shm->heap_base; }
    return 0;
  }
  if (top + need <= head) { g_shm->heap_top = top + need; return top; }              /* wrapped: [head, end) and [base, top) are live */
  return 0;
}

int MPI_Init(int* argc, char*** argv)
{
  (void)argc; (void)argv;
  const char* sz = getenv("RPB_MPI_SIZE");
  if (!sz || atoi(sz) <= 1) return MPI_SUCCESS;            /* one rank: the in-process transport */
  const char* rk = getenv("RPB_MPI_RANK");
  const char* path = getenv("RPB_MPI_SHM");
  if (!rk || !path) die("RPB_MPI_SIZE > 1 needs RPB_MPI_RANK and RPB_MPI_SHM");
  g_size = atoi(sz); g_rank = atoi(rk);
  if (g_size > SHM_MAX_RANKS || g_rank < 0 || g_rank >= g_size) die("bad RPB_MPI_SIZE / RPB_MPI_RANK");
  const int fd = open(path, O_RDWR);
  if (fd < 0) die("cannot open RPB_MPI_SHM");
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(shm_t) + (1u << 20)) die("RPB_MPI_SHM is too small");
  void* p = mmap(0, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) die("mmap");
  close(fd);
  g_shm = (shm_t*)p; g_arena = (char*)p;
  if (g_rank == 0) {
    pthread_mutexattr_t ma; pthread_condattr_t ca;
    pthread_mutexattr_init(&ma); pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
    pthread_condattr_init(&ca); pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
    pthread_mutex_init(&g_shm->mu, &ma); pthread_cond_init(&g_shm->cv, &ca);
    g_shm->heap_base = (sizeof(shm_t) + 63) & ~(size_t)63;
    g_shm->heap_top = g_shm->heap_base; g_shm->heap_end = (size_t)st.st_size;
    g_shm->nmsgs = 0; g_shm->first_untaken = 0;
    __atomic_store_n(&g_shm->ready, 1, __ATOMIC_RELEASE);
  } else {
    while (!__atomic_load_n(&g_shm->ready, __ATOMIC_ACQUIRE)) usleep(100);
  }
  shm_barrier();
  return MPI_SUCCESS;
}

int MPI_Finalize(void) { if (g_size > 1) shm_barrier(); return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) { (void)comm; exit(code ? code : 1); }
int MPI_Comm_rank(MPI_Comm comm, int* rank) { (void)comm; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int* size) { (void)comm; *size = g_size; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm comm) { (void)comm; if (g_size > 1) shm_barrier(); return MPI_SUCCESS; }

double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define REDUCE_LOOP(T)                                                                         \
  for (int i = 0; i < count; ++i) {                                                            \
    T acc = ((const T*)g_shm->coll[0])[i];                                                     \
    for (int r = 1; r < g_size; ++r) {                                                         \
      const T v = ((const T*)g_shm->coll[r])[i];                                               \
      acc = op == MPI_SUM ? acc + v : op == MPI_MIN ? (v < acc ? v : acc) : (v > acc ? v : acc); \
    }                                                                                          \
    ((T*)recv)[i] = acc;                                                                       \
  }

int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
  (void)comm;
  const size_t bytes = (size_t)count * type_bytes(type);
  if (g_size == 1) {                                       /* one rank: every reduction is the identity */
    if (send != recv) memcpy(recv, send, bytes);
    return MPI_SUCCESS;
  }
  if (bytes > SHM_SLOT_BYTES) die("Allreduce larger than the collective slot");
  if (op != MPI_SUM && op != MPI_MIN && op != MPI_MAX) die("unsupported reduction");
  memcpy(g_shm->coll[g_rank], send, bytes);
  shm_barrier();
  switch (type) {                                          /* every rank folds the P slots in rank order */
    case MPI_LONG_DOUBLE: REDUCE_LOOP(long double) break;
    case MPI_DOUBLE: REDUCE_LOOP(double) break;
    case MPI_FLOAT: REDUCE_LOOP(float) break;
    case MPI_INT: REDUCE_LOOP(int) break;
    case MPI_LONG: case MPI_LONG_LONG: REDUCE_LOOP(long long) break;
    default: die("unsupported Allreduce datatype");
  }
  shm_barrier();
  return MPI_SUCCESS;
}

int MPI_Gather(const void* send, int scount, MPI_Datatype stype, void* recv, int rcount, MPI_Datatype rtype,
               int root, MPI_Comm comm)
{
  (void)rcount; (void)rtype; (void)comm;
  const size_t bytes = (size_t)scount * type_bytes(stype);
  if (g_size == 1) { memcpy(recv, send, bytes); return MPI_SUCCESS; }
  if (bytes > SHM_SLOT_BYTES) die("Gather larger than the collective slot");
  memcpy(g_shm->coll[g_rank], send, bytes);
  shm_barrier();
  if (g_rank == root) for (int r = 0; r < g_size; ++r) memcpy((char*)recv + (size_t)r * bytes, g_shm->coll[r], bytes);
  shm_barrier();
  return MPI_SUCCESS;
}

int MPI_Allgather(const void* send, int scount, MPI_Datatype stype, void* recv, int rcount, MPI_Datatype rtype, MPI_Comm comm)
{
  (void)rcount; (void)rtype; (void)comm;
  const size_t bytes = (size_t)scount * type_bytes(stype);
  if (g_size == 1) { memcpy(recv, send, bytes); return MPI_SUCCESS; }
  if (bytes > SHM_SLOT_BYTES) die("Allgather larger than the collective slot");
  memcpy(g_shm->coll[g_rank], send, bytes);
  shm_barrier();
  for (int r = 0; r < g_size; ++r) memcpy((char*)recv + (size_t)r * bytes, g_shm->coll[r], bytes);
  shm_barrier();
  return MPI_SUCCESS;
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  (void)comm;
  if (g_size == 1) return MPI_SUCCESS;
  const size_t bytes = (size_t)count * type_bytes(type);
  if (bytes > SHM_SLOT_BYTES) die("Bcast larger than the collective slot");
  if (g_rank == root) memcpy(g_shm->coll[root], buf, bytes);
  shm_barrier();
  if (g_rank != root) memcpy(buf, g_shm->coll[root], bytes);
  shm_barrier();
  return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req)
{
  (void)comm;
  if (source < 0 || source >= g_size) die("Irecv from a rank outside the communicator");
  const size_t bytes = (size_t)count * type_bytes(type);
  const int r = new_slot();
  g_slots[r].kind = 1; g_slots[r].done = 0; g_slots[r].tag = tag; g_slots[r].buf = buf; g_slots[r].bytes = bytes;
  g_slots[r].data = 0; g_slots[r].delivered = 0; g_slots[r].released = 0; g_slots[r].peer = source;
  if (g_size > 1) { *req = r; return MPI_SUCCESS; }        /* P ranks: matched against the shared log in Wait* */
  /* an earlier buffered send with this tag?  (oldest first: MPI's non-overtaking rule) */
  for (int i = 0; i < g_nslots; ++i) {
    slot_t* s = &g_slots[i];
    if (s->kind == 2 && !s->delivered && s->tag == tag) {
      if (s->bytes > bytes) die("message longer than the posted receive");
      memcpy(buf, s->data, s->bytes);
      free(s->data); s->data = 0; s->delivered = 1;
      if (s->released) s->kind = 0;
      g_slots[r].done = 1;
      break;
    }
  }
  *req = r;
  return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req)
{
  (void)comm;
  if (dest < 0 || dest >= g_size) die("Isend to a rank outside the communicator");
  const size_t bytes = (size_t)count * type_bytes(type);
  const int r = new_slot();
  g_slots[r].kind = 2; g_slots[r].done = 1; g_slots[r].tag = tag; g_slots[r].buf = 0; g_slots[r].bytes = bytes;
  g_slots[r].data = 0; g_slots[r].delivered = 0; g_slots[r].released = 0; g_slots[r].peer = dest;
  if (g_size > 1) {                                        /* P ranks: append to the shared log, buffered */
    shm_lock();
    const size_t need = (bytes + 63) & ~(size_t)63;
    size_t off = 0;
    for (;;) {                                             /* a full ring: wait for receivers (60 s without progress: die) */
      if (g_shm->nmsgs - g_shm->first_untaken < SHM_MAX_MSGS && (need == 0 || (off = shm_alloc(need)) != 0)) break;
      shm_wait();
    }
    shm_msg_t* m = SHM_MSG(g_shm->nmsgs);
    g_shm->nmsgs++;
    m->src = g_rank; m->dst = dest; m->tag = tag; m->taken = 0; m->bytes = bytes; m->offset = need ? off : g_shm->heap_base;
    memcpy(g_arena + m->offset, buf, bytes);
    pthread_cond_broadcast(&g_shm->cv);
    shm_unlock();
    g_slots[r].delivered = 1;                              /* the payload left this process: the request is complete */
    *req = r;
    return MPI_SUCCESS;
  }
  for (int i = 0; i < g_nslots; ++i) {                     /* the oldest matching posted receive */
    slot_t* s = &g_slots[i];
    if (s->kind == 1 && !s->done && s->tag == tag) {
      if (bytes > s->bytes) die("message longer than the posted receive");
      memcpy(s->buf, buf, bytes);
      s->done = 1;
      g_slots[r].delivered = 1;
      break;
    }
  }
  if (!g_slots[r].delivered) {                             /* nobody is listening yet: keep a copy */
    g_slots[r].data = malloc(bytes ? bytes : 1);
    if (!g_slots[r].data) die("out of memory");
    memcpy(g_slots[r].data, buf, bytes);
  }
  *req = r;
  return MPI_SUCCESS;
}

static void complete(MPI_Request* req)
{
  const int r = *req;
  if (r == MPI_REQUEST_NULL) return;
  if (r < 0 || r >= g_nslots || g_slots[r].kind == 0) die("bad request");
  if (g_slots[r].kind == 1 && !g_slots[r].done) die("deadlock: waiting for a message nobody sent (one rank, no progress possible)");
  if (g_slots[r].kind == 2 && !g_slots[r].delivered) return;      /* a buffered send stays alive until it is received */
  g_slots[r].kind = 0;
  *req = MPI_REQUEST_NULL;
}

int MPI_Wait(MPI_Request* req, MPI_Status* status)
{
  (void)status;
  if (g_size > 1 && *req != MPI_REQUEST_NULL && g_slots[*req].kind == 1 && !g_slots[*req].done) {
    slot_t* s = &g_slots[*req];
    shm_lock();
    long i;
    while ((i = shm_find(s->peer, s->tag)) < 0) shm_wait();
    shm_take(i, s->buf, s->bytes);
    shm_unlock();
    s->done = 1;
  }
  if (*req != MPI_REQUEST_NULL && g_slots[*req].kind == 2 && !g_slots[*req].delivered) {
    /* completing a buffered send locally is legal; the slot itself lives on until a receive takes the data */
    g_slots[*req].released = 1;
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
  }
  complete(req);
  return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request* reqs, MPI_Status* statuses)
{
  (void)statuses;
  for (int i = 0; i < count; ++i) MPI_Wait(&reqs[i], MPI_STATUS_IGNORE);
  return MPI_SUCCESS;
}

int MPI_Waitany(int count, MPI_Request* reqs, int* index, MPI_Status* status)
{
  (void)status;
  int active = 0;
  for (int i = 0; i < count; ++i) {
    if (reqs[i] == MPI_REQUEST_NULL) continue;
    ++active;
    const slot_t* s = &g_slots[reqs[i]];
    if (s->kind == 2 || s->done) { MPI_Wait(&reqs[i], MPI_STATUS_IGNORE); *index = i; return MPI_SUCCESS; }
  }
  if (active && g_size > 1) {                              /* P ranks: whichever receive finds its message first */
    shm_lock();
    for (;;) {
      for (int i = 0; i < count; ++i) {
        if (reqs[i] == MPI_REQUEST_NULL) continue;
        slot_t* s = &g_slots[reqs[i]];
        const long m = shm_find(s->peer, s->tag);
        if (m >= 0) {
          shm_take(m, s->buf, s->bytes);
          shm_unlock();
          s->done = 1;
          MPI_Wait(&reqs[i], MPI_STATUS_IGNORE);
          *index = i;
          return MPI_SUCCESS;
        }
      }
      shm_wait();
    }
  }
  if (active) die("deadlock in Waitany: no request can complete on one rank");
  *index = MPI_UNDEFINED;
  return MPI_SUCCESS;
}
