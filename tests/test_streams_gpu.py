"""GPU: the ABI's stream and graph contract (include/rpb200.h, "STREAMS").

* one context, several streams, kernels that own scratch (DOT, REDUCE_SUM, SCAN, INDEXLIST) in flight CONCURRENTLY: every
  result still equals the oracle's (per-stream partials / tickets / look-back descriptors / epochs);
* a CUDA graph holding ONE scan (or index list) is replayed several times with different inputs: the look-back epoch is read
  from device memory at every replay, so no replay can see an earlier replay's descriptors as ready;
* the three PA kernels on different streams with DIFFERENT basis matrices (their __constant__ tables are per device: the calls
  serialise instead of racing);
* more streams than pre-allocated scratch sets; attach / detach.

Integer-valued inputs make every association exact, so the comparisons are bit-exact.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bits(a):
    return np.ascontiguousarray(a).view(np.int64)


SIZES = [100003, 300007, 1000003, 3000017]          # small enough that kernels of different streams are co-resident


def _inputs(seed):
    rng = np.random.default_rng(seed)
    out = []
    for n in SIZES:
        a = rng.integers(-20, 20, n).astype(np.float64)
        b = rng.integers(-20, 20, n).astype(np.float64)
        out.append((a, b))
    return out


def test_four_streams_reductions_scans_indexlists_concurrently(ctx):
    L = oracle.lib()
    data = _inputs(1)
    streams = [torch.cuda.Stream() for _ in SIZES]
    d_a = [dev(a) for a, _ in data]
    d_b = [dev(b) for _, b in data]
    f64 = dict(dtype=torch.float64, device="cuda")
    dot = [torch.zeros(1, **f64) for _ in SIZES]
    rsum = [torch.zeros(1, **f64) for _ in SIZES]
    scan = [torch.empty(n, **f64) for n in SIZES]
    lst = [torch.full((n,), -1, dtype=torch.int32, device="cuda") for n in SIZES]
    ln = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in SIZES]
    torch.cuda.synchronize()
    for _ in range(6):                               # many rounds: a race does not show every time
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                ctx.stream_dot(d_a[i], d_b[i], dot[i])
                ctx.reduce_sum(d_a[i], rsum[i])
                ctx.scan_exclusive(d_b[i], scan[i])
                ctx.indexlist(d_a[i], lst[i], ln[i])
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(data):
        n = a.size
        assert dot[i].item() == L.orc_stream_dot(a, b, n, 0.0), i
        assert rsum[i].item() == L.orc_reduce_sum(a, n, 0.0), i
        ref = np.empty(n); L.orc_scan_exclusive(b, ref, n)
        assert np.array_equal(bits(scan[i].cpu().numpy()), bits(ref)), i
        want = np.full(n, -1, dtype=np.int32)
        k = L.orc_indexlist(a, want, n)
        assert ln[i].item() == k and np.array_equal(lst[i].cpu().numpy(), want), i
    for st in streams:
        ctx.stream_detach(st.cuda_stream)


@pytest.mark.parametrize("n", [200003, (1 << 23) + 48])     # register-staged kernels; TMA-staged kernels (n >= 2 * 8192 * SMs)
def test_one_scan_graph_replayed_with_new_inputs(ctx, n):
    L = oracle.lib()
    rng = np.random.default_rng(n)
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    lst = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    ln = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.scan_reserve(n)                              # a capture may not allocate (rpb200.h)
    ctx.indexlist_reserve(n)
    # the capture stream gets its scratch set BEFORE the capture starts: this session's context has served more streams than
    # it pre-allocates sets for, and a capture may not allocate (rpb200.h, rpb200_stream_attach)
    cap = torch.cuda.Stream()
    ctx.stream_attach(cap.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        ctx.scan_exclusive(x, y)                     # ONE scan and ONE index list in the graph
        ctx.indexlist(x, lst, ln)
    for rep in range(4):
        h = rng.integers(-9, 9, n).astype(np.float64)
        x.copy_(dev(h))
        y.fill_(-1.0); lst.fill_(-1)
        g.replay()
        torch.cuda.synchronize()
        ref = np.empty(n); L.orc_scan_exclusive(h, ref, n)
        assert np.array_equal(bits(y.cpu().numpy()), bits(ref)), rep
        want = np.full(n, -1, dtype=np.int32)
        k = L.orc_indexlist(h, want, n)
        assert ln.item() == k and np.array_equal(lst.cpu().numpy(), want), rep
    # and the same context keeps working outside the graph afterwards
    ctx.scan_exclusive(x, y)
    torch.cuda.synchronize()
    assert np.array_equal(bits(y.cpu().numpy()), bits(ref))
    del g
    ctx.stream_detach(cap.cuda_stream)


def test_reduction_graph_replayed_with_new_inputs(ctx):
    L = oracle.lib()
    n = 777781
    rng = np.random.default_rng(5)
    a = torch.zeros(n, dtype=torch.float64, device="cuda"); b = torch.zeros_like(a)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    cap = torch.cuda.Stream()
    ctx.stream_attach(cap.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        for _ in range(3):
            ctx.stream_dot(a, b, out, accumulate=True)     # DOT-Seq.cpp:45: m_dot += dot, three reps in one graph
    for rep in range(3):
        ha, hb = rng.integers(-9, 9, n).astype(np.float64), rng.integers(-9, 9, n).astype(np.float64)
        a.copy_(dev(ha)); b.copy_(dev(hb)); out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert out.item() == 3.0 * L.orc_stream_dot(ha, hb, n, 0.0)
    del g
    ctx.stream_detach(cap.cuda_stream)


def test_pa_kernels_on_two_streams_with_different_bases(ctx):
    """MASS3DPA on stream 0 with B = 1 and on stream 1 with B = 2 (integer-valued => exact): the __constant__ tables are one
    set per device, so the calls must serialise; each result must be the one its own basis defines."""
    L = oracle.lib()
    NE = 4096
    f64 = dict(dtype=torch.float64, device="cuda")
    res = []
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bufs = []
    for i, scale in enumerate((1.0, 2.0)):
        B = np.full(20, scale); Bt = np.full(20, scale)
        D = np.ones(125 * NE); X = np.ones(64 * NE); Y = np.zeros(64 * NE)
        bufs.append((B, Bt, D, X, Y, dev(B), dev(Bt), dev(D), dev(X), torch.zeros(64 * NE, **f64)))
    torch.cuda.synchronize()
    for _ in range(5):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                _, _, _, _, _, dB, dBt, dD, dX, dY = bufs[i]
                ctx.mass3dpa(dB, dBt, dD, dX, dY, NE)
    torch.cuda.synchronize()
    for i in range(2):
        B, Bt, D, X, Y, *_ , dY = bufs[i]
        ref = np.zeros(64 * NE)
        for _ in range(5):
            L.orc_mass3dpa(B, Bt, D, X, ref, NE)
        assert np.array_equal(bits(dY.cpu().numpy()), bits(ref)), i
    for st in streams:
        ctx.stream_detach(st.cuda_stream)


def test_more_streams_than_preallocated_sets_and_detach(ctx):
    L = oracle.lib()
    n = 50021
    a = np.random.default_rng(9).integers(-5, 5, n).astype(np.float64)
    d_a = dev(a)
    want = L.orc_reduce_sum(a, n, 0.0)
    streams = [torch.cuda.Stream() for _ in range(11)]          # more than the 8 pre-allocated scratch sets
    outs = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in streams]
    torch.cuda.synchronize()
    for st, o in zip(streams, outs):
        with torch.cuda.stream(st):
            ctx.reduce_sum(d_a, o)
    torch.cuda.synchronize()
    assert all(o.item() == want for o in outs)
    for st in streams:
        ctx.stream_detach(st.cuda_stream)
    with pytest.raises(Exception):
        ctx.stream_detach(streams[0].cuda_stream)          # not attached any more
    # attach explicitly, then capture on that (5th+) stream without any allocation inside the capture
    st = torch.cuda.Stream()
    ctx.stream_attach(st.cuda_stream)
    o = torch.zeros(1, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        ctx.reduce_sum(d_a, o)
    torch.cuda.synchronize()
    assert o.item() == want
    ctx.stream_detach(st.cuda_stream)


def test_create_leaves_the_current_device_alone_and_calls_check_it():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from rajaperf_b200 import Context, cabi
    torch.cuda.set_device(0)
    c1 = Context(1)
    assert torch.cuda.current_device() == 0
    x = torch.ones(1000, dtype=torch.float64, device="cuda:1")
    o = torch.zeros(1, dtype=torch.float64, device="cuda:1")
    rc = c1.lib.rpb200_reduce_sum(c1.h, x.data_ptr(), 1000, 0.0, o.data_ptr(), None)
    assert rc == -19                                       # RPB200_EDEVICE: the current device is 0
    torch.cuda.set_device(1)
    c1.reduce_sum(x, o)
    torch.cuda.synchronize()
    assert o.item() == 1000.0
    c1.close()
    torch.cuda.set_device(0)
