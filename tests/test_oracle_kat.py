"""Pin the CPU oracle to the reference: every case in tests/golden/ref_checksums.json is a
checksum printed by the reference's own raja-perf.exe (Base_Seq); the oracle's restatement of
setUp -> reps -> updateChecksum must reproduce it to the printed 20 significant digits."""
import json
import os

import numpy as np
import pytest

import oracle

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_checksums.json")))
# the kernels the reference compiles only WITH MPI, from the unmodified sources linked against the one-rank MPI
# stand-in of oracle/mpi_stub (oracle/build_ref_mpi.sh; generator: tests/golden/make_golden.py --mpi)
GOLD_MPI1 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_checksums_mpi1.json")))
# seeded random draws from both reference binaries (make_golden.py --fuzz): ragged sizes around the tile boundaries, odd
# LTIMES shapes, halo widths 1-3 with 1-6 variables, random rank grids of 1-8 ranks for the MPI-only kernels
GOLD_FUZZ = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_checksums_fuzz.json")))


def _iparams(kernel, flags, ranks=1):
    flags = list(flags)
    division = None
    if "--mpi_3d_division" in flags:                       # the one option with three values
        i = flags.index("--mpi_3d_division")
        division = [int(v) for v in flags[i + 1:i + 4]]
        del flags[i:i + 4]
    f = dict(zip(flags[0::2], flags[1::2]))
    if kernel == "Apps_LTIMES":
        return [int(f.get("--ltimes_num_d", 64)), int(f.get("--ltimes_num_g", 32)), int(f.get("--ltimes_num_m", 25))]
    if kernel.startswith("Comm_HALO"):
        if division is None:                               # the suite's default factorisation (RunParams.cpp:1211-1251)
            from rajaperf_b200.dist import rank_grid
            division = rank_grid(ranks)
        assert division[0] * division[1] * division[2] == ranks
        return [int(f.get("--halo_width", 1)), int(f.get("--halo_num_vars", 3))] + division
    return None


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"{c['kernel']}-s{c['size']}-r{c['reps']}")
def test_oracle_reproduces_reference_checksum(case):
    got = oracle.kat(case["kernel"], case["size"], case["reps"], _iparams(case["kernel"], case["flags"]))
    ref = np.longdouble(case["checksum"])
    # 20 printed digits of an 80-bit long double: allow 2 units in the 19th significant digit
    assert abs(got - ref) <= abs(ref) * np.longdouble(2e-19) + np.longdouble(0), (got, ref)


@pytest.mark.parametrize("case", GOLD_MPI1["cases"],
                         ids=lambda c: f"{c['kernel']}-s{c['size']}-r{c['reps']}-p{c['ranks']}-{'_'.join(x for x in c['flags'] if not x.startswith('--'))}")
def test_oracle_reproduces_reference_exchange_checksums(case):
    """HALO_EXCHANGE, HALO_EXCHANGE_FUSED, HALO_SENDRECV: the oracle's in-process delivery against the reference's own
    Irecv / pack / Isend / Waitall / unpack code, run on 1 rank (every neighbour is the rank itself) and on 2 / 4 / 6 / 8
    ranks (P processes over the MPI stand-in's shared-memory transport; the report averages the P checksums)."""
    got = oracle.kat(case["kernel"], case["size"], case["reps"], _iparams(case["kernel"], case["flags"], case["ranks"]))
    ref = np.longdouble(case["checksum"])
    tol = np.longdouble(2e-19) if case["ranks"] == 1 else np.longdouble(1e-18)      # + the rounding of the rank average
    assert abs(got - ref) <= abs(ref) * tol, (got, ref)


@pytest.mark.parametrize("case", GOLD_FUZZ["cases"],
                         ids=lambda c: f"{c['kernel']}-s{c['size']}-r{c['reps']}-p{c['ranks']}-{'_'.join(x for x in c['flags'] if not x.startswith('--'))}")
def test_oracle_reproduces_reference_checksums_on_random_inputs(case):
    """184 more pins: the oracle against the unmodified reference on inputs nobody picked by hand."""
    ranks = max(case["ranks"], 1)
    got = oracle.kat(case["kernel"], case["size"], case["reps"], _iparams(case["kernel"], case["flags"], ranks))
    ref = np.longdouble(case["checksum"])
    tol = np.longdouble(2e-19) if ranks == 1 else np.longdouble(1e-18)               # + the rounding of the rank average
    assert abs(got - ref) <= abs(ref) * tol, (got, ref)


def test_reference_multi_rank_checksums_equal_its_one_rank_checksums():
    """What SURVEY 8c derived on paper, now from the reference's own runs: every rank holds identical data, so the P-rank
    result equals the periodic self-exchange of one rank, for any rank grid."""
    one = {(c["kernel"], c["size"], c["reps"], _key(c["flags"])): np.longdouble(c["checksum"]) for c in GOLD_MPI1["cases"] if c["ranks"] == 1}
    multi = [c for c in GOLD_MPI1["cases"] if c["ranks"] > 1]
    hit = 0
    for c in multi:
        k = (c["kernel"], c["size"], c["reps"], _key(c["flags"]))
        if k in one:
            hit += 1
            assert abs(np.longdouble(c["checksum"]) - one[k]) <= abs(one[k]) * np.longdouble(1e-18), c
    assert len(multi) >= 15 and hit >= 9


def _key(flags):
    flags = list(flags)
    if "--mpi_3d_division" in flags:
        i = flags.index("--mpi_3d_division")
        del flags[i:i + 4]
    return tuple(flags)


def test_mpi_stub_build_agrees_with_the_plain_build_on_the_pack_kernels():
    """The two pack kernels exist in both reference builds: identical checksums, so the MPI-stub binary is the same suite."""
    plain = {(c["kernel"], c["size"], c["reps"], tuple(c["flags"])): c["checksum"] for c in GOLD["cases"]}
    both = [c for c in GOLD_MPI1["cases"] if (c["kernel"], c["size"], c["reps"], tuple(c["flags"])) in plain]
    assert len(both) >= 3
    for c in both:
        assert c["checksum"] == plain[(c["kernel"], c["size"], c["reps"], tuple(c["flags"]))]


def test_halo_exchange_single_rank_matches_any_rank_grid():
    """All ranks hold identical vars, so the P-rank exchange equals the 1-rank periodic
    self-exchange; and pack results match HALO_PACKING_FUSED's send buffers."""
    base = oracle.kat("Comm_HALO_EXCHANGE_FUSED", 8000, 2, [1, 3, 1, 1, 1])
    for pd in ([2, 1, 1], [2, 2, 1], [2, 2, 2], [3, 1, 2]):
        got = oracle.kat("Comm_HALO_EXCHANGE_FUSED", 8000, 2, [1, 3] + pd)
        # the report averages P identical checksums in long double: allow that rounding
        assert abs(got - base) <= abs(base) * np.longdouble(1e-18), (pd, got, base)


def test_diffusion_basis_fill_covers_all_entries():
    """The reference's aliased 12-double basis array starts uninitialised; fill #1 must write
    every entry (else Base_Seq would read garbage).  SURVEY appendix A.3 tables for b=g=1."""
    L = oracle.lib()
    b = np.ones(12)
    g = np.ones(12)
    f1 = np.empty(12)
    f2 = np.empty(12)
    L.orc_diffusion3dpa_tables(b, g, f1, f2)
    assert not np.isnan(f1).any() and not np.isnan(f2).any()
    assert f1.tolist() == [1, 1, 1, -1, 1, 1, -1, 1, 1, -1, 1, 1]
    assert f2.tolist() == [1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1]


def test_init_counter_parity():
    """DataUtils.cpp:504-513: factor 0.2 on even call counts, 0.1 on odd."""
    a0 = oracle.init_real(4, 0)
    a1 = oracle.init_real(4, 1)
    assert a0[0] == 0.2 * 1.1 / 1.12345 and a1[0] == 0.1 * 1.1 / 1.12345
