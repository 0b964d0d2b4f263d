"""GPU parity: Apps group (MASS3DPA, DIFFUSION3DPA, CONVECTION3DPA, LTIMES) through the C ABI vs
the CPU oracle and the reference's golden checksums.

PA kernels: the suite's data is integer-valued, so every summation order is exact -> bit-exact
class.  With random data the regrouped FMA contractions differ from Base_Seq by
rounding only; the tolerance (relative 5e-13 of the largest |Y|) is stated in the tests.
LTIMES: tolerance class, 1e-7 absolute on the suite checksum (test/test-raja-perf-suite.cpp:167)."""
import json
import os

import numpy as np
import pytest

import oracle
import suite_data as sd

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLD = {(c["kernel"], c["size"], c["reps"], tuple(c["flags"])): c["checksum"]
        for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_checksums.json")))["cases"]}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bits(a):
    return np.ascontiguousarray(a).view(np.int64)


def run_pa(ctx, name, d, reps=1):
    Y = dev(d["Y"])
    if name == "mass":
        args = [dev(d["B"]), dev(d["Bt"]), dev(d["D"]), dev(d["X"]), Y]
        for _ in range(reps): ctx.mass3dpa(*args, d["NE"])
    elif name == "diffusion":
        args = [dev(d["B"]), dev(d["G"]), dev(d["D"]), dev(d["X"]), Y]
        for _ in range(reps): ctx.diffusion3dpa(*args, d["NE"], d.get("symmetric", True))
    else:
        args = [dev(d["B"]), dev(d["Bt"]), dev(d["G"]), dev(d["D"]), dev(d["X"]), Y]
        for _ in range(reps): ctx.convection3dpa(*args, d["NE"])
    return Y.cpu().numpy()


def run_pa_oracle(name, d, reps=1):
    L = oracle.lib()
    Y = d["Y"].copy()
    for _ in range(reps):
        if name == "mass":
            L.orc_mass3dpa(d["B"], d["Bt"], d["D"], d["X"], Y, d["NE"])
        elif name == "diffusion":
            L.orc_diffusion3dpa(d["B"], d["G"], d["D"], d["X"], Y, d["NE"], 1 if d.get("symmetric", True) else 0)
        else:
            L.orc_convection3dpa(d["B"], d["Bt"], d["G"], d["D"], d["X"], Y, d["NE"])
    return Y


MAKE = {"mass": sd.mass3dpa, "diffusion": sd.diffusion3dpa, "convection": sd.convection3dpa}
FULL = {"mass": "Apps_MASS3DPA", "diffusion": "Apps_DIFFUSION3DPA", "convection": "Apps_CONVECTION3DPA"}


@pytest.mark.parametrize("name", ["mass", "diffusion", "convection"])
@pytest.mark.parametrize("size,reps", [(0, 1), (0, 3), (1, 1), (5000, 2)])
def test_pa_suite_checksum_matches_reference_golden(ctx, name, size, reps):
    d = MAKE[name](size)
    got = oracle.checksum(run_pa(ctx, name, d, reps), 1.0)
    ref = np.longdouble(GOLD[(FULL[name], size, reps, ())])
    assert abs(got - ref) <= abs(ref) * np.longdouble(2e-19), (got, ref)


@pytest.mark.parametrize("name", ["mass", "diffusion", "convection"])
@pytest.mark.parametrize("NE", [1, 2, 31, 32, 33, 1000])
def test_pa_suite_data_bit_exact_vs_oracle(ctx, name, NE):
    unit = 125 if name == "mass" else 64
    d = MAKE[name](NE * unit)
    assert d["NE"] == NE
    assert np.array_equal(bits(run_pa(ctx, name, d, 2)), bits(run_pa_oracle(name, d, 2)))


@pytest.mark.parametrize("name", ["mass", "diffusion", "convection"])
@pytest.mark.parametrize("NE", [1, 37, 1025])
def test_pa_integer_valued_random_data_bit_exact(ctx, name, NE):
    """Small random integers in every array (basis included): all products and sums are exact, so
    the regrouped contractions must agree bit for bit -- exercises every index of every table,
    including DIFFUSION3DPA's aliased half-stored basis fill order."""
    rng = np.random.default_rng(NE)
    unit = 125 if name == "mass" else 64
    d = MAKE[name](NE * unit)
    for k in d:
        if k not in ("NE",):
            d[k] = rng.integers(-3, 4, d[k].size).astype(np.float64)
    assert np.array_equal(bits(run_pa(ctx, name, d, 1)), bits(run_pa_oracle(name, d, 1)))


@pytest.mark.parametrize("name,variant", [("mass", 10), ("mass", 12), ("mass", 13), ("mass", 16), ("mass", 17), ("mass", 18), ("mass", 19), ("mass", 21), ("mass", 23), ("mass", 25), ("mass", 30),
                                          ("convection", 10), ("convection", 11), ("convection", 14), ("convection", 16)])
def test_pa_launch_shape_tunings_bit_exact(ctx, name, variant):
    """Every selectable launch shape (elements per CTA / threads / ring stages / CTAs per SM, Y staged or not) on
    integer-valued random data, two reps, with a ragged last batch."""
    NE = 1029
    rng = np.random.default_rng(variant)
    d = MAKE[name](NE * (125 if name == "mass" else 64))
    for k in d:
        if k not in ("NE",):
            d[k] = rng.integers(-3, 4, d[k].size).astype(np.float64)
    kernel = "Apps_MASS3DPA" if name == "mass" else "Apps_CONVECTION3DPA"
    ctx.set_tuning(kernel, -1, -1, variant)
    try:
        got = run_pa(ctx, name, d, 2)
    finally:
        ctx.set_tuning(kernel, -1, -1, 1)
    assert np.array_equal(bits(got), bits(run_pa_oracle(name, d, 2)))


@pytest.mark.parametrize("name", ["mass", "diffusion", "convection"])
def test_pa_random_real_data_rounding_level(ctx, name):
    rng = np.random.default_rng(5)
    NE = 777
    unit = 125 if name == "mass" else 64
    d = MAKE[name](NE * unit)
    for k in d:
        if k not in ("NE",):
            d[k] = rng.standard_normal(d[k].size)
    got, ref = run_pa(ctx, name, d, 1), run_pa_oracle(name, d, 1)
    assert np.max(np.abs(got - ref)) <= 5e-13 * np.max(np.abs(ref))


def test_diffusion_nonsymmetric_path(ctx):
    """symmetric=false reads slabs 3..8 with the SYM=6 element stride (DIFFUSION3DPA.hpp:389-397),
    i.e. runs into the next element's storage; give D three slabs of slack like the oracle needs."""
    rng = np.random.default_rng(9)
    NE = 65
    d = sd.diffusion3dpa(NE * 64)
    d["D"] = rng.integers(-2, 3, 64 * 6 * NE + 3 * 64).astype(np.float64)
    d["X"] = rng.integers(-2, 3, 27 * NE).astype(np.float64)
    d["symmetric"] = False
    assert np.array_equal(bits(run_pa(ctx, "diffusion", d, 1)), bits(run_pa_oracle("diffusion", d, 1)))


# ------------------------------------------------------------------------------------------ LTIMES
def run_ltimes(ctx, d, reps=1):
    phi = dev(d["phi"]); ell = dev(d["ell"]); psi = dev(d["psi"])
    for _ in range(reps):
        ctx.ltimes(phi, ell, psi, d["nd"], d["ng"], d["nm"], d["nz"])
    return phi.cpu().numpy()


@pytest.mark.parametrize("size,reps,flags", [(0, 1, ()), (0, 3, ()), (5000, 2, ()),
                                             (100000, 1, ("--ltimes_num_d", "32", "--ltimes_num_g", "8", "--ltimes_num_m", "17"))])
def test_ltimes_suite_checksum_matches_reference_golden(ctx, size, reps, flags):
    f = dict(zip(flags[0::2], flags[1::2]))
    d = sd.ltimes(size, int(f.get("--ltimes_num_d", 64)), int(f.get("--ltimes_num_g", 32)), int(f.get("--ltimes_num_m", 25)))
    got = oracle.checksum(run_ltimes(ctx, d, reps), d["scale"])
    ref = np.longdouble(GOLD[("Apps_LTIMES", size, reps, flags)])
    assert abs(got - ref) < 1e-7, (got, ref)


@pytest.mark.parametrize("nz", [1, 3, 7, 100])
def test_ltimes_matches_oracle_elementwise(ctx, nz):
    d = sd.ltimes(nz * 2048)
    assert d["nz"] == nz
    got = run_ltimes(ctx, d, 2)
    ref = d["phi"].copy()
    for _ in range(2):
        oracle.lib().orc_ltimes(ref, d["ell"], d["psi"], 64, 32, 25, nz)
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))     # dot of length 64, reordered


@pytest.mark.parametrize("variant", [1, 24, 31, 33, 34])
@pytest.mark.parametrize("NE", [8, 1024, 40000, 1029])
def test_mass_line_major_accesses_bit_exact(ctx, variant, NE):
    """MASS3DPA with line-major X / Y accesses (pieces <-> slabs through a shared-memory tile): whole batches take the
    line-major kernel (the default, 1; 31 = 12 CTAs per SM), a ragged element count (1029) falls back to the
    slab-per-thread kernel (24); random integer-valued data and basis, two reps (Y accumulates)."""
    rng = np.random.default_rng(NE + variant)
    d = sd.mass3dpa(NE * 125)
    assert d["NE"] == NE
    for k in ("D", "X", "Y"):
        d[k] = rng.integers(-3, 4, d[k].size).astype(np.float64)
    d["B"] = rng.integers(-2, 3, d["B"].size).astype(np.float64)
    d["Bt"] = rng.integers(-2, 3, d["Bt"].size).astype(np.float64)
    ctx.set_tuning("Apps_MASS3DPA", -1, -1, variant)
    try:
        got = run_pa(ctx, "mass", d, 2)
    finally:
        ctx.set_tuning("Apps_MASS3DPA", -1, -1, 1)
    assert np.array_equal(bits(got), bits(run_pa_oracle("mass", d, 2)))


@pytest.mark.parametrize("variant", [5, 6, 8, 10])
def test_ltimes_staged_variants_integer_valued_bit_exact(ctx, variant):
    """The opt-in psi-ring kernels and the row-chunk fragment mapping (10; the default is line-major): other permutations of d, applied to psi and
    ell alike, on integer-valued data."""
    nz, nd, ng, nm = 37, 64, 32, 25
    rng = np.random.default_rng(variant)
    phi = rng.integers(-5, 6, nz * ng * nm).astype(np.float64)
    ell = rng.integers(-3, 4, nm * nd).astype(np.float64)
    psi = rng.integers(-3, 4, nz * ng * nd).astype(np.float64)
    ref = phi.copy()
    oracle.lib().orc_ltimes(ref, ell, psi, nd, ng, nm, nz)
    d_phi = dev(phi)
    ctx.set_tuning("Apps_LTIMES", -1, -1, variant)
    try:
        ctx.ltimes(d_phi, dev(ell), dev(psi), nd, ng, nm, nz)
    finally:
        ctx.set_tuning("Apps_LTIMES", -1, -1, 4)
    assert np.array_equal(bits(d_phi.cpu().numpy()), bits(ref))


def test_ltimes_integer_valued_data_bit_exact(ctx):
    rng = np.random.default_rng(2)
    d = sd.ltimes(40 * 2048)
    d["ell"] = rng.integers(-4, 5, d["ell"].size).astype(np.float64)
    d["psi"] = rng.integers(-4, 5, d["psi"].size).astype(np.float64)
    d["phi"] = rng.integers(-4, 5, d["phi"].size).astype(np.float64)
    got = run_ltimes(ctx, d, 1)
    ref = d["phi"].copy(); oracle.lib().orc_ltimes(ref, d["ell"], d["psi"], 64, 32, 25, d["nz"])
    assert np.array_equal(bits(got), bits(ref))


def test_ltimes_other_shapes_take_the_generic_kernel(ctx):
    rng = np.random.default_rng(4)
    for nd, ng, nm, nz in [(32, 8, 17, 11), (64, 3, 25, 5), (7, 5, 3, 2)]:
        d = dict(nd=nd, ng=ng, nm=nm, nz=nz, phi=rng.integers(-3, 4, nm * ng * nz).astype(np.float64),
                 ell=rng.integers(-3, 4, nd * nm).astype(np.float64), psi=rng.integers(-3, 4, nd * ng * nz).astype(np.float64))
        got = run_ltimes(ctx, d, 1)
        ref = d["phi"].copy(); oracle.lib().orc_ltimes(ref, d["ell"], d["psi"], nd, ng, nm, nz)
        assert np.array_equal(bits(got), bits(ref))


def test_apps_full_size_properties(ctx):
    """BASELINE config #4 sizes.  Elements are independent, so (1) the suite's all-ones data has a
    closed form per dof (MASS 8000, CONVECTION 5184, DIFFUSION the 27-entry pattern),
    checked over the whole array; (2) a random slice of elements is compared with the oracle."""
    NE = 4000000
    one = lambda n: torch.ones(n, dtype=torch.float64, device="cuda")
    Y = torch.zeros(64 * NE, dtype=torch.float64, device="cuda")
    ctx.mass3dpa(one(20), one(20), one(125 * NE), one(64 * NE), Y, NE)
    ctx.mass3dpa(one(20), one(20), one(125 * NE), one(64 * NE), Y, NE)
    assert bool((Y == 16000.0).all())
    del Y
    Y = torch.zeros(27 * NE, dtype=torch.float64, device="cuda")
    ctx.convection3dpa(one(12), one(12), one(12), one(192 * NE), one(27 * NE), Y, NE)
    assert bool((Y == 5184.0).all())
    Y.zero_()
    ctx.diffusion3dpa(one(12), one(12), one(384 * NE), one(27 * NE), Y, NE, True)
    pat = torch.tensor([0, 576, 0, 576, 1152, 576, 0, 576, 0, 576, 1152, 576, 1152, 1728, 1152, 576, 1152, 576,
                        0, 576, 0, 576, 1152, 576, 0, 576, 0], dtype=torch.float64, device="cuda")
    assert bool((Y.view(NE, 27) == pat).all())
    del Y
    # LTIMES at num_z = 500000: rows are independent; compare the first and last 64 zones with the oracle
    nz = 500000
    d = sd.ltimes(2048 * 50)          # ell from the suite; psi random at full size
    psi = torch.rand(2048 * nz, dtype=torch.float64, device="cuda")
    phi = torch.zeros(800 * nz, dtype=torch.float64, device="cuda")
    ctx.ltimes(phi, dev(d["ell"]), psi, 64, 32, 25, nz)
    for z0 in (0, nz - 64):
        p = psi[2048 * z0: 2048 * (z0 + 64)].cpu().numpy()
        ref = np.zeros(800 * 64); oracle.lib().orc_ltimes(ref, d["ell"], p, 64, 32, 25, 64)
        got = phi[800 * z0: 800 * (z0 + 64)].cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
