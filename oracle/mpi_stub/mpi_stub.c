/* mpi_stub.c -- see mpi.h: one rank, in process, test infrastructure only. */
#include "mpi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  int kind;        /* 0 free, 1 receive, 2 send */
  int done;        /* receive: a message was delivered; send: always (buffered)                */
  int tag;
  void* buf;       /* receive: destination                                                     */
  size_t bytes;
  void* data;      /* send: private copy of the payload while no receive has been posted yet   */
  int delivered;   /* send: payload already copied into a receive                              */
  int released;    /* send: its request was waited on before a receive took the payload        */
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

int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) { (void)comm; exit(code ? code : 1); }
int MPI_Comm_rank(MPI_Comm comm, int* rank) { (void)comm; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int* size) { (void)comm; *size = 1; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm comm) { (void)comm; return MPI_SUCCESS; }

double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
  (void)op; (void)comm;                                   /* one rank: every reduction is the identity */
  if (send != recv) memcpy(recv, send, (size_t)count * type_bytes(type));
  return MPI_SUCCESS;
}

int MPI_Gather(const void* send, int scount, MPI_Datatype stype, void* recv, int rcount, MPI_Datatype rtype,
               int root, MPI_Comm comm)
{
  (void)rcount; (void)rtype; (void)root; (void)comm;
  memcpy(recv, send, (size_t)scount * type_bytes(stype));
  return MPI_SUCCESS;
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  (void)buf; (void)count; (void)type; (void)root; (void)comm;
  return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req)
{
  (void)comm;
  if (source != 0) die("Irecv from a rank other than 0 on a one-rank communicator");
  const size_t bytes = (size_t)count * type_bytes(type);
  const int r = new_slot();
  g_slots[r].kind = 1; g_slots[r].done = 0; g_slots[r].tag = tag; g_slots[r].buf = buf; g_slots[r].bytes = bytes;
  g_slots[r].data = 0; g_slots[r].delivered = 0; g_slots[r].released = 0;
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
  if (dest != 0) die("Isend to a rank other than 0 on a one-rank communicator");
  const size_t bytes = (size_t)count * type_bytes(type);
  const int r = new_slot();
  g_slots[r].kind = 2; g_slots[r].done = 1; g_slots[r].tag = tag; g_slots[r].buf = 0; g_slots[r].bytes = bytes;
  g_slots[r].data = 0; g_slots[r].delivered = 0; g_slots[r].released = 0;
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
  if (active) die("deadlock in Waitany: no request can complete on one rank");
  *index = MPI_UNDEFINED;
  return MPI_SUCCESS;
}
