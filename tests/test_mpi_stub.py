"""The one-rank MPI stand-in (oracle/mpi_stub: test infrastructure that lets the unmodified reference run its MPI-only
Comm kernels here) must match messages the way MPI does on a self-communicator: by tag, oldest first, in any legal
order of Irecv / Isend / Wait*.  The exchange goldens (tests/golden/ref_checksums_mpi1.json) rest on these rules."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "mpi_stub", "libmpistub.so")
MPI_DOUBLE, MPI_BYTE, MPI_LONG_DOUBLE, MPI_SUM = 8, 1, 9, 1
MPI_UNDEFINED, MPI_REQUEST_NULL = -32766, -1


@pytest.fixture(scope="module")
def mpi():
    if not os.path.exists(SO):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    L = ctypes.CDLL(SO)
    L.MPI_Wtime.restype = ctypes.c_double
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_rank_size_and_identity_collectives(mpi):
    r, s = ctypes.c_int(-1), ctypes.c_int(-1)
    mpi.MPI_Comm_rank(0, ctypes.byref(r)); mpi.MPI_Comm_size(0, ctypes.byref(s))
    assert (r.value, s.value) == (0, 1)
    a, b = np.arange(5, dtype=np.float64), np.zeros(5)
    assert mpi.MPI_Allreduce(_p(a), _p(b), 5, MPI_DOUBLE, MPI_SUM, 0) == 0 and np.array_equal(a, b)
    c = np.zeros(40, dtype=np.uint8)
    assert mpi.MPI_Gather(_p(a), 40, MPI_BYTE, _p(c), 40, MPI_BYTE, 0, 0) == 0 and np.array_equal(c.view(np.float64), a)
    assert mpi.MPI_Barrier(0) == 0 and mpi.MPI_Wtime() > 0.0


def test_receives_posted_first_are_matched_by_tag(mpi):
    """The suite's order: Irecv x N, Isend x N, Waitall (HALO_EXCHANGE_FUSED-Seq.cpp:35-116)."""
    n = 26
    recv = [np.full(4, -1.0) for _ in range(n)]
    send = [np.full(4, float(t)) for t in range(n)]
    rreq, sreq = (ctypes.c_int * n)(), (ctypes.c_int * n)()
    for t in range(n):                                   # receive l carries tag 25 - l, like recv_tag = opposite neighbour
        mpi.MPI_Irecv(_p(recv[t]), 4, MPI_DOUBLE, 0, n - 1 - t, 0, ctypes.byref(rreq, 4 * t))
    for t in range(n):
        mpi.MPI_Isend(_p(send[t]), 4, MPI_DOUBLE, 0, t, 0, ctypes.byref(sreq, 4 * t))
    assert mpi.MPI_Waitall(n, rreq, None) == 0 and mpi.MPI_Waitall(n, sreq, None) == 0
    assert all(r == MPI_REQUEST_NULL for r in rreq) and all(r == MPI_REQUEST_NULL for r in sreq)
    for t in range(n):
        assert np.all(recv[t] == float(n - 1 - t))


def test_sends_before_receives_are_buffered_and_do_not_overtake(mpi):
    first, second = np.full(3, 1.0), np.full(3, 2.0)
    s1, s2, r1, r2 = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    mpi.MPI_Isend(_p(first), 3, MPI_DOUBLE, 0, 7, 0, ctypes.byref(s1))
    mpi.MPI_Isend(_p(second), 3, MPI_DOUBLE, 0, 7, 0, ctypes.byref(s2))
    first[:] = -9.0                                      # the payload was copied at Isend
    mpi.MPI_Wait(ctypes.byref(s1), None)                 # a buffered send completes locally
    a, b = np.zeros(3), np.zeros(3)
    mpi.MPI_Irecv(_p(a), 3, MPI_DOUBLE, 0, 7, 0, ctypes.byref(r1))
    mpi.MPI_Irecv(_p(b), 3, MPI_DOUBLE, 0, 7, 0, ctypes.byref(r2))
    mpi.MPI_Wait(ctypes.byref(r1), None); mpi.MPI_Wait(ctypes.byref(r2), None); mpi.MPI_Wait(ctypes.byref(s2), None)
    assert np.all(a == 1.0) and np.all(b == 2.0)


def test_waitany_hands_out_every_completed_request_once(mpi):
    """HALO_EXCHANGE-Seq.cpp:34-116 unpacks in MPI_Waitany order."""
    n = 5
    recv = [np.zeros(2) for _ in range(n)]
    rreq = (ctypes.c_int * n)()
    for t in range(n):
        mpi.MPI_Irecv(_p(recv[t]), 2, MPI_DOUBLE, 0, 100 + t, 0, ctypes.byref(rreq, 4 * t))
    sreq = ctypes.c_int()
    for t in (3, 0, 4, 1, 2):
        payload = np.full(2, float(t))
        mpi.MPI_Isend(_p(payload), 2, MPI_DOUBLE, 0, 100 + t, 0, ctypes.byref(sreq))
        mpi.MPI_Wait(ctypes.byref(sreq), None)
    seen = []
    for _ in range(n):
        idx = ctypes.c_int(-1)
        mpi.MPI_Waitany(n, rreq, ctypes.byref(idx), None)
        seen.append(idx.value)
    assert sorted(seen) == list(range(n))
    idx = ctypes.c_int(0)
    mpi.MPI_Waitany(n, rreq, ctypes.byref(idx), None)
    assert idx.value == MPI_UNDEFINED
    for t in range(n):
        assert np.all(recv[t] == float(t))


_RANK_PROGRAM = r"""
import ctypes, sys
import numpy as np
L = ctypes.CDLL(sys.argv[1])
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
L.MPI_Init(None, None)
r, s = ctypes.c_int(), ctypes.c_int()
L.MPI_Comm_rank(0, ctypes.byref(r)); L.MPI_Comm_size(0, ctypes.byref(s))
r, s = r.value, s.value
# ring: two messages with the same tag to the right neighbour (must not overtake), one with another tag to the left
right, left = (r + 1) % s, (r - 1) % s
a1, a2, b = np.full(4, 10.0 * r + 1), np.full(4, 10.0 * r + 2), np.full(2, 100.0 + r)
req = (ctypes.c_int * 6)()
ra1, ra2, rb = np.zeros(4), np.zeros(4), np.zeros(2)
L.MPI_Irecv(p(ra1), 4, 8, left, 5, 0, ctypes.byref(req, 0))
L.MPI_Irecv(p(ra2), 4, 8, left, 5, 0, ctypes.byref(req, 4))
L.MPI_Irecv(p(rb), 2, 8, right, 6, 0, ctypes.byref(req, 8))
L.MPI_Isend(p(a1), 4, 8, right, 5, 0, ctypes.byref(req, 12))
L.MPI_Isend(p(a2), 4, 8, right, 5, 0, ctypes.byref(req, 16))
L.MPI_Isend(p(b), 2, 8, left, 6, 0, ctypes.byref(req, 20))
done = []
for _ in range(3):
    i = ctypes.c_int(-1)
    L.MPI_Waitany(3, req, ctypes.byref(i), None)
    done.append(i.value)
L.MPI_Waitall(6, req, None)
ok = sorted(done) == [0, 1, 2] and np.all(ra1 == 10.0 * left + 1) and np.all(ra2 == 10.0 * left + 2) and np.all(rb == 100.0 + right)
# collectives: long double sum / max in rank order, gather of bytes, broadcast
x = np.array([r + 0.25], dtype=np.longdouble); y = np.zeros(1, dtype=np.longdouble)
L.MPI_Allreduce(p(x), p(y), 1, 9, 1, 0)
ok = ok and y[0] == sum(q + 0.25 for q in range(s))
L.MPI_Allreduce(p(x), p(y), 1, 9, 3, 0)
ok = ok and y[0] == s - 1 + 0.25
g = np.zeros(8 * s, dtype=np.uint8); mine = np.array([float(r)])
L.MPI_Gather(p(mine), 8, 1, p(g), 8, 1, 0, 0)
if r == 0:
    ok = ok and g.view(np.float64).tolist() == [float(q) for q in range(s)]
ag = np.zeros(s, dtype=np.float64)
L.MPI_Allgather(p(mine), 8, 1, p(ag), 8, 1, 0)
ok = ok and ag.tolist() == [float(q) for q in range(s)]
v = np.array([42.0 if r == 0 else -1.0])
L.MPI_Bcast(p(v), 1, 8, 0, 0)
ok = ok and v[0] == 42.0
L.MPI_Barrier(0)
L.MPI_Finalize()
sys.exit(0 if ok else 3)
"""


@pytest.mark.parametrize("nranks", [2, 5])
def test_shared_memory_transport_ring_and_collectives(mpi, nranks, tmp_path):
    """P processes over the stub's shared-memory arena: tag + source matching without overtaking, Waitany, and the
    collectives the suite's reports use (Executor.cpp:70-118)."""
    import subprocess
    import sys
    prog = tmp_path / "rank.py"
    prog.write_text(_RANK_PROGRAM)
    arena = f"/dev/shm/rpb_mpi_test_{os.getpid()}_{nranks}"
    with open(arena, "wb") as f:
        f.truncate(64 << 20)
    try:
        procs = [subprocess.Popen([sys.executable, str(prog), SO],
                                  env=dict(os.environ, RPB_MPI_SIZE=str(nranks), RPB_MPI_RANK=str(r), RPB_MPI_SHM=arena))
                 for r in range(nranks)]
        rcs = [p.wait(timeout=120) for p in procs]
    finally:
        os.unlink(arena)
    assert rcs == [0] * nranks


_RING_PROGRAM = r"""
import ctypes, sys
import numpy as np
L = ctypes.CDLL(sys.argv[1])
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
L.MPI_Init(None, None)
r, s = ctypes.c_int(), ctypes.c_int()
L.MPI_Comm_rank(0, ctypes.byref(r)); L.MPI_Comm_size(0, ctypes.byref(s))
r, s = r.value, s.value
right, left = (r + 1) % s, (r - 1) % s
n = 1 << 19                                   # 4 MiB messages
ok = True
for rep in range(40):                         # 40 reps x s ranks x 2 messages x 4 MiB >> the 48 MiB of payload room
    out1, out2 = np.full(n, 1000.0 * rep + r), np.full(n, -1000.0 * rep - r)
    in1, in2 = np.zeros(n), np.zeros(n)
    req = (ctypes.c_int * 4)()
    L.MPI_Irecv(p(in1), n, 8, left, 1, 0, ctypes.byref(req, 0))
    L.MPI_Irecv(p(in2), n, 8, right, 2, 0, ctypes.byref(req, 4))
    L.MPI_Isend(p(out1), n, 8, right, 1, 0, ctypes.byref(req, 8))
    L.MPI_Isend(p(out2), n, 8, left, 2, 0, ctypes.byref(req, 12))
    L.MPI_Waitall(4, req, None)               # no barrier between reps: ranks drift apart, messages of several reps are in flight
    ok = ok and np.all(in1 == 1000.0 * rep + left) and np.all(in2 == -1000.0 * rep - right)
L.MPI_Finalize()
sys.exit(0 if ok else 3)
"""


def test_shared_memory_arena_is_a_ring(mpi, tmp_path):
    """More bytes than the arena holds pass through it when ranks are not in lock step: the message log and the payload heap
    are rings, and a sender that finds no room waits for a receiver (round 2: a bump allocator that was only rewound when
    nothing was in flight ran out under 8 GPU ranks and hung the reference driver)."""
    import subprocess
    import sys
    nranks = 4
    prog = tmp_path / "ring.py"
    prog.write_text(_RING_PROGRAM)
    arena = f"/dev/shm/rpb_mpi_test_ring_{os.getpid()}"
    with open(arena, "wb") as f:
        f.truncate(96 << 20)                   # ~45 MiB of bookkeeping + ~50 MiB of payload room: 12 messages
    try:
        procs = [subprocess.Popen([sys.executable, str(prog), SO],
                                  env=dict(os.environ, RPB_MPI_SIZE=str(nranks), RPB_MPI_RANK=str(r), RPB_MPI_SHM=arena))
                 for r in range(nranks)]
        rcs = [p.wait(timeout=180) for p in procs]
    finally:
        os.unlink(arena)
    assert rcs == [0] * nranks
