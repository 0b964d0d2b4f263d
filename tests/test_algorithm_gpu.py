"""GPU parity: Algorithm group (SCAN, SORT, SORTPAIRS) through the C ABI vs the CPU oracle and
the reference's golden checksums.  SORT / SORTPAIRS are bit-exact; SCAN is in the tolerance
class: 1e-7 absolute on the suite checksum."""
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


def bits(a):
    return np.ascontiguousarray(a).view(np.int64)


# ------------------------------------------------------------------------------------- SCAN
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 4096, 4097, 123457, 1000000, (1 << 22) + 3])
def test_scan_matches_oracle(ctx, n):
    d = sd.scan(n)
    y = dev(d["y"])
    ctx.scan_exclusive(dev(d["x"]), y)
    ref = np.empty(n); oracle.lib().orc_scan_exclusive(d["x"], ref, n)
    got = y.cpu().numpy()
    assert got[0] == 0.0
    # association differs from the serial sum: relative 1e-12 of the running total is far inside
    # the suite's tolerance (its own OpenMP scan drifts 2e-13 relative at 2^27)
    assert np.all(np.abs(got - ref) <= 1e-12 * np.maximum(np.abs(ref), 1.0))


@pytest.mark.parametrize("size,reps", [(0, 1), (0, 3), (1, 1), (1000, 2), (123457, 2)])
def test_scan_suite_checksum_matches_reference_golden(ctx, size, reps):
    n = size or 1000000
    d = sd.scan(n)
    x, y = dev(d["x"]), dev(d["y"])
    for _ in range(reps):
        ctx.scan_exclusive(x, y)
    got = oracle.checksum(y.cpu().numpy(), sd.scan_scale(n))
    ref = np.longdouble(GOLD[("Algorithm_SCAN", size, reps)])
    assert abs(got - ref) < 1e-7, (got, ref)          # test/test-raja-perf-suite.cpp:167


def test_scan_integer_valued_input_is_bit_exact(ctx):
    """With integer-valued doubles every association is exact: the scan must be bit-identical."""
    n = 3000017
    rng = np.random.default_rng(7)
    x = rng.integers(0, 1000, n).astype(np.float64)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.scan_exclusive(dev(x), y)
    ref = np.empty(n); oracle.lib().orc_scan_exclusive(x, ref, n)
    assert np.array_equal(bits(y.cpu().numpy()), bits(ref))


@pytest.mark.parametrize("tune", [(128, 1, 1), (256, 2, 2), (512, 4, 4), (64, 8, 4)])
def test_scan_every_tuning_and_repeated_calls(ctx, tune):
    n = 777777
    ctx.set_tuning("Algorithm_SCAN", *tune)
    x = np.random.default_rng(3).integers(0, 50, n).astype(np.float64)
    ref = np.empty(n); oracle.lib().orc_scan_exclusive(x, ref, n)
    xd = dev(x); y = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(4):                   # epoch-tagged look-back state: no clearing between calls
        y.fill_(-1.0)
        ctx.scan_exclusive(xd, y)
        assert np.array_equal(bits(y.cpu().numpy()), bits(ref))
    ctx.set_tuning("Algorithm_SCAN", 256, 4, 4)


def test_scan_unaligned_and_growing_sizes(ctx):
    x = np.random.default_rng(5).integers(0, 9, 200001).astype(np.float64)
    xd = dev(x)
    for n in (100, 200000, 5):           # look-back state grows, then is reused for a smaller scan
        y = torch.empty(n + 1, dtype=torch.float64, device="cuda")
        ctx.scan_exclusive(xd[1:], y[1:], n=n)           # 8-byte aligned only
        ref = np.empty(n); oracle.lib().orc_scan_exclusive(x[1:n + 1].copy(), ref, n)
        assert np.array_equal(bits(y[1:].cpu().numpy()), bits(ref))


def test_scan_full_size_properties(ctx):
    """BASELINE config #3 size (2^27): y[0] = 0, differences reproduce x, last = total."""
    n = 1 << 27
    x = torch.randint(0, 4, (n,), device="cuda").to(torch.float64)     # exact arithmetic
    y = torch.empty_like(x)
    ctx.scan_exclusive(x, y)
    assert y[0].item() == 0.0
    assert torch.equal(y[1:] - y[:-1], x[:-1])
    assert y[-1].item() + x[-1].item() == x.sum().item()


# ------------------------------------------------------------------------------------- SORT
def _scratch(ctx, n, pairs):
    nb = ctx.sort_scratch_bytes(n, pairs)
    return torch.empty(max(nb, 8) // 8 + 1, dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("n", [1, 2, 3, 100, 4096, 4097, 65537, 1000000, (1 << 21) + 11])
def test_sort_keys_bit_exact(ctx, n):
    x = oracle.init_rand_value(n)
    k = dev(x)
    ctx.sort_keys(k, _scratch(ctx, n, False))
    ref = x.copy(); oracle.lib().orc_sort(ref, n)
    assert np.array_equal(bits(k.cpu().numpy()), bits(ref))


def test_sort_negative_zero_denormal_inf_keys(ctx):
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.standard_normal(50000) * 1e300, rng.standard_normal(50000) * 1e-310,
                        [0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324, 1.0, -1.0] * 10])
    rng.shuffle(x)
    n = x.size
    k = dev(x)
    ctx.sort_keys(k, _scratch(ctx, n, False))
    got = k.cpu().numpy()
    ref = np.sort(x)                      # -0.0 and +0.0 compare equal for a comparison sort
    assert np.array_equal(got, ref)
    assert np.all(got[1:] >= got[:-1])
    # radix order is the IEEE total order: every -0.0 before every +0.0
    z = got[got == 0.0]
    sgn = np.signbit(z)
    assert np.all(sgn[:-1] >= sgn[1:])


def _adversarial_keys(kind, n, rng):
    """Key sets aimed at the parts of the onesweep pass: uniform tiles (every key of an 8192-key tile shares the pass's digit:
    the tile skips ranking and moves as one block), tiles that are ALMOST uniform, and the look-back across runs of both."""
    if kind == "all_equal":                      # every tile uniform in every pass
        return np.full(n, 0.37)
    if kind == "sorted":                         # high passes: long runs of uniform tiles with different digits
        return np.sort(rng.random(n))
    if kind == "reversed":
        return np.sort(rng.random(n))[::-1].copy()
    if kind == "suite_like":                     # rand()/RAND_MAX in [0, 1]: top byte 0x3f except for a handful of tiny keys
        return rng.integers(0, 2**31 - 1, n).astype(np.float64) / 2147483647.0
    if kind == "one_stranger_per_tile":          # uniform but for ONE key per tile, at varying positions
        x = np.full(n, 0.5)
        pos = np.arange(0, n, 8192) + (np.arange(0, (n + 8191) // 8192) * 977) % 8192
        x[pos[pos < n]] = -3.0
        return x
    if kind == "two_values_blocks":              # alternating uniform tiles of two digits, block = one tile
        return np.where((np.arange(n) // 8192) % 2 == 0, 1.0, 2.0**40)
    if kind == "mixed_signs_runs":
        x = rng.standard_normal(n)
        x[: n // 3] = np.abs(x[: n // 3]); x[n // 3: 2 * n // 3] = -np.abs(x[n // 3: 2 * n // 3])
        return x
    raise KeyError(kind)


ADVERSARIAL = ["all_equal", "sorted", "reversed", "suite_like", "one_stranger_per_tile", "two_values_blocks", "mixed_signs_runs"]


@pytest.mark.parametrize("hist", ["lanes", "shared_bins", "no_uniform_test", "lookback_1"])
@pytest.mark.parametrize("kind", ADVERSARIAL)
@pytest.mark.parametrize("n", [8192, 8193, 3 * 8192, 1000003])
def test_sort_keys_adversarial_sets_bit_exact(ctx, n, kind, hist):
    rng = np.random.default_rng(n + len(kind))
    x = _adversarial_keys(kind, n, rng)
    ctx.set_tuning("Algorithm_SORT", ctas_per_sm=1 if hist == "lookback_1" else 4, unroll={"shared_bins": 8, "no_uniform_test": 7}.get(hist, 4))
    try:
        k = dev(x)
        ctx.sort_keys(k, _scratch(ctx, n, False))
        got = k.cpu().numpy()
    finally:
        ctx.reset_tuning("Algorithm_SORT")
    ref = x.copy(); oracle.lib().orc_sort(ref, n)
    assert np.array_equal(bits(got), bits(ref))


@pytest.mark.parametrize("byte", range(8))
def test_sort_keys_one_byte_varies(ctx, byte):
    """Keys that differ in ONE byte only: seven passes see nothing but uniform tiles, one pass sees random digits."""
    n = 5 * 8192 + 17
    rng = np.random.default_rng(byte)
    base = np.float64(1.2345678901234567).view(np.uint64)
    b = rng.integers(0, 128 if byte == 7 else 256, n).astype(np.uint64)          # byte 7: keep the sign bit clear
    keys = ((base & ~(np.uint64(0xff) << np.uint64(8 * byte))) | (b << np.uint64(8 * byte))).view(np.float64)
    keys = keys[np.isfinite(keys)] if byte >= 6 else keys
    keys = keys[~np.isnan(keys)]
    n = keys.size
    k = dev(keys)
    ctx.sort_keys(k, _scratch(ctx, n, False))
    ref = keys.copy(); oracle.lib().orc_sort(ref, n)
    assert np.array_equal(bits(k.cpu().numpy()), bits(ref))


@pytest.mark.parametrize("kind", ["all_equal", "sorted", "suite_like", "one_stranger_per_tile", "two_values_blocks"])
@pytest.mark.parametrize("n", [8192, 3 * 8192 + 5, 1000003])
def test_sort_pairs_adversarial_sets_stable(ctx, n, kind):
    """The uniform-tile path must keep the order of equal keys (the values are the original positions)."""
    rng = np.random.default_rng(n)
    keys = _adversarial_keys(kind, n, rng)
    vals = np.arange(n, dtype=np.float64)
    k, v = dev(keys), dev(vals)
    ctx.sort_pairs(k, v, _scratch(ctx, n, True))
    order = np.argsort(keys.view(np.int64) ^ ((keys.view(np.int64) >> 63) & 0x7fffffffffffffff), kind="stable")   # IEEE total order
    assert np.array_equal(bits(k.cpu().numpy()), bits(keys[order]))
    assert np.array_equal(v.cpu().numpy(), vals[order])


@pytest.mark.parametrize("size,reps", [(0, 1), (0, 3), (1, 1), (1000, 2), (123457, 2)])
def test_sort_suite_checksum_matches_reference_golden(ctx, size, reps):
    """SORT.cpp:56-61: one rand() stream of n*reps keys, rep r sorts segment r in place."""
    n = size or 1000000
    x = dev(oracle.init_rand_value(n * reps))
    scratch = _scratch(ctx, n, False)
    for r in range(reps):
        ctx.sort_keys(x[n * r:], scratch, n=n)
    got = oracle.checksum(x.cpu().numpy(), 1.0)
    ref = np.longdouble(GOLD[("Algorithm_SORT", size, reps)])
    assert abs(got - ref) <= abs(ref) * np.longdouble(2e-19), (got, ref)


@pytest.mark.parametrize("n", [1, 2, 100, 4097, 65537, 1000000])
def test_sort_pairs_bit_exact_and_stable(ctx, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, max(2, n // 8), n).astype(np.float64) / 7.0     # many duplicate keys
    vals = np.arange(n, dtype=np.float64)
    k, v = dev(keys), dev(vals)
    ctx.sort_pairs(k, v, _scratch(ctx, n, True))
    rk, rv = keys.copy(), vals.copy(); oracle.lib().orc_sort_pairs(rk, rv, n)   # stable merge sort
    assert np.array_equal(bits(k.cpu().numpy()), bits(rk))
    assert np.array_equal(bits(v.cpu().numpy()), bits(rv))


@pytest.mark.parametrize("size,reps", [(0, 1), (0, 3), (1000, 2), (123457, 2)])
def test_sortpairs_suite_checksum_matches_reference_golden(ctx, size, reps):
    n = size or 1000000
    x = dev(oracle.init_rand_value(n * reps)); i = dev(oracle.init_rand_value(n * reps))
    scratch = _scratch(ctx, n, True)
    for r in range(reps):
        ctx.sort_pairs(x[n * r:], i[n * r:], scratch, n=n)
    got = oracle.checksum(x.cpu().numpy(), 1.0) + oracle.checksum(i.cpu().numpy(), 1.0)
    ref = np.longdouble(GOLD[("Algorithm_SORTPAIRS", size, reps)])
    assert abs(got - ref) <= abs(ref) * np.longdouble(4e-19), (got, ref)


def test_sort_full_size_properties(ctx):
    """BASELINE config #3 size (2^27 keys): sortedness + multiset preserved (sum and xor of the
    bit patterns are permutation-invariant) + idempotence."""
    n = 1 << 27
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    xi = x.view(torch.int64)
    s0 = xi.sum().item()
    x0 = 0
    for chunk in xi.split(1 << 24):
        x0 ^= int(np.bitwise_xor.reduce(chunk.cpu().numpy()))
    scratch = _scratch(ctx, n, False)
    ctx.sort_keys(x, scratch)
    assert bool((x[1:] >= x[:-1]).all())
    assert xi.sum().item() == s0
    x1 = 0
    for chunk in xi.split(1 << 24):
        x1 ^= int(np.bitwise_xor.reduce(chunk.cpu().numpy()))
    assert x1 == x0
    y = x.clone()
    ctx.sort_keys(y, scratch)
    assert torch.equal(x, y)
