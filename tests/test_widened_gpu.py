"""GPU parity for the widened rows: Basic_INDEXLIST / INDEXLIST_3LOOP (bit-exact integer
output) and Polybench_GEMM (FP64 tolerance class) through the C ABI vs the CPU oracle and the reference's
golden checksums."""
import json
import os

import numpy as np
import pytest

import oracle
import suite_data as sd

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLD = {(c["kernel"], c["size"], c["reps"]): c["checksum"]
        for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_checksums.json")))["cases"]
        if not c["flags"]}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------- INDEXLIST
@pytest.mark.parametrize("n", [1, 2, 31, 33, 1000, 8191, 8192, 8193, 123457, 1000000, (1 << 22) + 5])
def test_indexlist_matches_oracle_bit_exact(ctx, n):
    d = sd.indexlist(n)
    lst = dev(d["list"]); ln = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    ctx.indexlist(dev(d["x"]), lst, ln)
    ref = d["list"].copy()
    ref_len = oracle.lib().orc_indexlist(d["x"], ref, n)
    assert int(ln.item()) == ref_len
    assert np.array_equal(lst.cpu().numpy(), ref)          # selected prefix AND the untouched tail
    ref3 = d["list"].copy()
    assert oracle.lib().orc_indexlist_3loop(d["x"], ref3, n) == ref_len and np.array_equal(ref3, ref)


@pytest.mark.parametrize("kernel", ["Basic_INDEXLIST", "Basic_INDEXLIST_3LOOP"])
@pytest.mark.parametrize("size,reps", [(0, 1), (0, 3), (1, 1), (1000, 2), (123457, 2)])
def test_indexlist_suite_checksum_matches_reference_golden(ctx, kernel, size, reps):
    n = size or 1000000
    d = sd.indexlist(n)
    x, lst = dev(d["x"]), dev(d["list"])
    ln = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    for _ in range(reps):
        ctx.indexlist(x, lst, ln)
    got = oracle.checksum_int(lst.cpu().numpy()) + np.longdouble(int(ln.item()))     # INDEXLIST.cpp:69-73
    ref = np.longdouble(GOLD[(kernel, size, reps)])
    assert abs(got - ref) <= abs(ref) * np.longdouble(2e-19), (got, ref)


def test_indexlist_edge_cases(ctx):
    ln = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    ctx.indexlist(torch.empty(0, dtype=torch.float64, device="cuda"), torch.empty(0, dtype=torch.int32, device="cuda"), ln, n=0)
    assert int(ln.item()) == 0
    n = 100003
    for fill, want in ((-1.0, n), (1.0, 0), (0.0, 0), (-0.0, 0)):       # x < 0.0 is false for -0.0
        x = torch.full((n,), fill, dtype=torch.float64, device="cuda")
        lst = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        ctx.indexlist(x, lst, ln)
        assert int(ln.item()) == want
        got = lst.cpu().numpy()
        assert np.array_equal(got[:want], np.arange(want, dtype=np.int32)) and np.all(got[want:] == -7)
    # unaligned x (sub-range starting at an odd element): scalar path
    x = torch.randn(n + 1, dtype=torch.float64, device="cuda")[1:]
    lst = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    ctx.indexlist(x, lst, ln)
    ref = torch.nonzero(x < 0).flatten().to(torch.int32)
    assert int(ln.item()) == ref.numel() and torch.equal(lst[:ref.numel()], ref)


def test_indexlist_full_size_properties(ctx):
    """2^27 elements (the Algorithm-group size): ascending, complete, consistent with a torch count."""
    n = 1 << 27
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    lst = torch.empty(n, dtype=torch.int32, device="cuda")
    ln = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(2):                                   # second call: epoch-tagged state is reused
        ctx.indexlist(x, lst, ln)
    m = int(ln.item())
    assert m == int((x < 0).sum().item())
    sel = lst[:m].long()
    assert bool((sel[1:] > sel[:-1]).all()) and bool((x[sel] < 0).all())


# ------------------------------------------------------------------------------------- MEMSET / MEMCPY
@pytest.mark.parametrize("n", [1, 3, 4, 5, 1000, 123457, (1 << 22) + 3])
@pytest.mark.parametrize("offset", [0, 1, 3])
def test_memset_writes_exactly_the_range(ctx, n, offset):
    """x[i] = val on [offset, offset + n) of a sentinel-filled array: every element inside, nothing outside (unaligned heads)."""
    buf = torch.full((n + 8,), -1.234567e89, dtype=torch.float64, device="cuda")
    ctx.memset_f64(buf[offset:offset + n], 0.25, n)
    got = buf.cpu().numpy()
    assert np.all(got[offset:offset + n] == 0.25) and np.all(got[:offset] == -1.234567e89) and np.all(got[offset + n:] == -1.234567e89)


@pytest.mark.parametrize("kernel", ["Algorithm_MEMSET", "Algorithm_MEMCPY"])
@pytest.mark.parametrize("size,reps", [(0, 1), (1, 1), (123457, 2)])
def test_calibration_streams_match_reference_golden(ctx, kernel, size, reps):
    n = size or 1000000
    y = torch.full((n,), -1.234567e89, dtype=torch.float64, device="cuda")         # MEMCPY.cpp:59-63, MEMSET.cpp:59-63
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        if kernel == "Algorithm_MEMSET":
            ctx.memset_f64(y, 0.0)
        else:
            ctx.stream_copy(y, x)
    got = oracle.checksum(y.cpu().numpy())
    assert got == np.longdouble(GOLD[(kernel, size, reps)]) == 0      # 0 iff every sentinel was overwritten


# ------------------------------------------------------------------------------------- POLYBENCH_GEMM
@pytest.mark.parametrize("ni,nj,nk", [(1, 1, 1), (3, 5, 7), (8, 8, 4), (64, 64, 16), (65, 63, 17), (100, 100, 120),
                                      (129, 257, 33), (200, 130, 0), (333, 334, 401)])
def test_gemm_matches_oracle(ctx, ni, nj, nk):
    rng = np.random.default_rng(ni * 1000 + nj)
    A, B = rng.random(ni * nk), rng.random(nk * nj)
    C = dev(np.full(ni * nj, 3.0))
    ctx.polybench_gemm(dev(A) if nk else None, dev(B) if nk else None, C, ni, nj, nk, 0.62, 1.002)
    ref = np.full(ni * nj, 3.0)
    oracle.lib().orc_polybench_gemm(A, B, ref, ni, nj, nk, 0.62, 1.002)
    got = C.cpu().numpy()
    # different association + FMA: a few ulp of the running sum (tolerance class)
    assert np.all(np.abs(got - ref) <= 1e-13 * np.maximum(np.abs(ref), 1.0)), np.abs(got - ref).max()


@pytest.mark.parametrize("tile", [64, 128])
def test_gemm_integer_valued_is_bit_exact_both_tiles(ctx, tile):
    """Integer-valued operands, alpha = 1: every association is exact."""
    ni, nj, nk = 300, 260, 150
    rng = np.random.default_rng(3)
    A = rng.integers(-8, 9, ni * nk).astype(np.float64); B = rng.integers(-8, 9, nk * nj).astype(np.float64)
    C = torch.zeros(ni * nj, dtype=torch.float64, device="cuda")
    ctx.set_tuning("Polybench_GEMM", tile, -1, -1)
    try:
        ctx.polybench_gemm(dev(A), dev(B), C, ni, nj, nk, 1.0)
    finally:
        ctx.set_tuning("Polybench_GEMM", 256, -1, -1)
    ref = (A.reshape(ni, nk) @ B.reshape(nk, nj)).reshape(-1)
    assert np.array_equal(C.cpu().numpy(), ref)


@pytest.mark.parametrize("size,reps", [(0, 1), (0, 2), (1, 1), (10000, 2), (54321, 1)])
def test_gemm_suite_checksum_matches_reference_golden(ctx, size, reps):
    d = sd.polybench_gemm(size)
    A, B, C = dev(d["A"]), dev(d["B"]), dev(d["C"])
    for _ in range(reps):
        ctx.polybench_gemm(A, B, C, d["ni"], d["nj"], d["nk"], d["alpha"], d["beta"])
    got = oracle.checksum(C.cpu().numpy(), d["scale"])
    ref = np.longdouble(GOLD[("Polybench_GEMM", size, reps)])
    assert abs(got - ref) < 1e-7, (got, ref)          # test/test-raja-perf-suite.cpp:167


def test_gemm_large_linearity(ctx):
    """4096 x 4096 x 4912 (both tile kernels' big-grid path): C(A, B1 + B2) == C(A, B1) + C(A, B2) to rounding,
    and a sampled set of entries against float64 dot products."""
    ni = nj = 4096; nk = 4912
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand(ni * nk, dtype=torch.float64, device="cuda", generator=g)
    B1 = torch.rand(nk * nj, dtype=torch.float64, device="cuda", generator=g)
    B2 = torch.rand(nk * nj, dtype=torch.float64, device="cuda", generator=g)
    C1, C2, C12 = (torch.empty(ni * nj, dtype=torch.float64, device="cuda") for _ in range(3))
    ctx.polybench_gemm(A, B1, C1, ni, nj, nk, 0.62)
    ctx.polybench_gemm(A, B2, C2, ni, nj, nk, 0.62)
    ctx.polybench_gemm(A, B1 + B2, C12, ni, nj, nk, 0.62)
    assert float(((C1 + C2) - C12).abs().max()) <= 1e-12 * float(C12.abs().max())
    rows = torch.tensor([0, 1, 777, 4095], device="cuda"); cols = torch.tensor([0, 5, 2048, 4095], device="cuda")
    want = 0.62 * (A.view(ni, nk)[rows] @ B1.view(nk, nj)[:, cols])
    got = C1.view(ni, nj)[rows][:, cols]
    assert float((want - got).abs().max()) <= 1e-12 * float(want.abs().max())
