"""CPU checks of the oracle's restatement of the widened rows against independent numpy formulations;
the whole-kernel checksums are pinned to the reference binary in test_oracle_kat.py."""
import numpy as np
import pytest

import oracle
import suite_data as sd


@pytest.mark.parametrize("n", [1, 2, 17, 1000, 65537])
def test_indexlist_oracles_agree_with_numpy(n):
    d = sd.indexlist(n)
    L = oracle.lib()
    a, b = d["list"].copy(), d["list"].copy()
    la, lb = L.orc_indexlist(d["x"], a, n), L.orc_indexlist_3loop(d["x"], b, n)
    want = np.nonzero(d["x"] < 0.0)[0].astype(np.int32)
    assert la == lb == want.size
    assert np.array_equal(a[:la], want) and np.array_equal(b[:lb], want)
    assert np.array_equal(a[la:], d["list"][la:]) and np.array_equal(b[lb:], d["list"][lb:])    # tail untouched
    # initDataRandSign: about half negative, magnitudes from the 0.2 * (i + 1.1) / (i + 1.12345) sequence
    assert np.allclose(np.abs(d["x"]), 0.2 * (np.arange(n) + 1.1) / (np.arange(n) + 1.12345), rtol=0, atol=1e-15)


def test_indexlist_negative_zero_and_nan_are_not_selected():
    x = np.array([-0.0, 0.0, np.nan, -1e-300, 1.0, -np.inf])
    lst = np.full(6, -7, dtype=np.int32)
    assert oracle.lib().orc_indexlist(x, lst, 6) == 2 and lst.tolist() == [3, 5, -7, -7, -7, -7]


@pytest.mark.parametrize("target,dims", [(0, (1000, 1000, 1200)), (1, (1, 1, 1)), (10000, (100, 100, 120)), (54321, (233, 233, 279)),
                                         (16777216, (4096, 4096, 4915))])
def test_gemm_dims_follow_the_reference_constructor(target, dims):
    ni, nj, nk = (np.zeros(1, dtype=np.int64) for _ in range(3))
    oracle.lib().orc_polybench_gemm_dims(target, ni, nj, nk)
    assert (int(ni[0]), int(nj[0]), int(nk[0])) == dims


def test_gemm_oracle_matches_numpy_and_ignores_beta():
    rng = np.random.default_rng(1)
    ni, nj, nk = 37, 29, 53
    A, B = rng.standard_normal(ni * nk), rng.standard_normal(nk * nj)
    C1, C2 = np.full(ni * nj, 7.0), np.full(ni * nj, -3.0)
    oracle.lib().orc_polybench_gemm(A, B, C1, ni, nj, nk, 0.62, 1.002)
    oracle.lib().orc_polybench_gemm(A, B, C2, ni, nj, nk, 0.62, 123.0)
    assert np.array_equal(C1, C2)                                  # "C *= beta" is overwritten by "C = dot"
    want = 0.62 * (A.reshape(ni, nk) @ B.reshape(nk, nj)).reshape(-1)
    assert np.max(np.abs(C1 - want)) <= 1e-13 * np.max(np.abs(want))


def test_calibration_stream_checksums_are_zero():
    for k in ("Algorithm_MEMCPY", "Algorithm_MEMSET"):
        for size, reps in ((0, 1), (1, 1), (123457, 2)):
            assert oracle.kat(k, size, reps) == 0


def test_unfused_comm_kernels_share_the_fused_results():
    """HALO_PACKING == HALO_PACKING_FUSED and HALO_EXCHANGE == HALO_EXCHANGE_FUSED: same setUp, same data flow."""
    assert oracle.kat("Comm_HALO_PACKING", 27000, 2, [2, 5, 1, 1, 1]) == oracle.kat("Comm_HALO_PACKING_FUSED", 27000, 2, [2, 5, 1, 1, 1])
    assert oracle.kat("Comm_HALO_EXCHANGE", 8000, 2, [1, 3, 2, 2, 1]) == oracle.kat("Comm_HALO_EXCHANGE_FUSED", 8000, 2, [1, 3, 2, 2, 1])


def test_sendrecv_checksum_is_rank_grid_invariant_and_rep_idempotent():
    base = oracle.kat("Comm_HALO_SENDRECV", 8000, 1, [1, 3, 1, 1, 1])
    for pd in ([2, 1, 1], [2, 2, 2], [3, 1, 2]):
        assert oracle.kat("Comm_HALO_SENDRECV", 8000, 1, [1, 3] + pd) == base
    assert oracle.kat("Comm_HALO_SENDRECV", 8000, 4, [1, 3, 1, 1, 1]) == base        # the same payload every rep
    # message l lands where recv_tag == l: the checksum is NOT that of the send buffers in place
    L = oracle.lib()
    L.orc_reset_init_count()
    assert base > 0
