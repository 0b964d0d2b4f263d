"""Index arithmetic of the lane-private digit histogram (csrc/sort.cu sort_hist_lanes_kernel, the default since round 2)
restated in numpy: not a test of the CUDA code (tests/test_algorithm_gpu.py runs that, both histogram kernels), a test of the
counter-packing scheme and of the overflow bound its flush interval rests on."""
import numpy as np


def test_lane_private_packed_counters_reproduce_the_digit_histograms():
    """sort_hist_lanes_kernel: bin b of digit d is the 16-bit half (b & 1) of word col[d][(b >> 1) * 32 + lane]; a column is
    shared by the warps of the CTA, and everything is folded into the 64-bit histograms before a half can overflow."""
    rng = np.random.default_rng(5)
    warps, flush_keys = 32, 48          # the kernel folds every 2040 keys per thread; 48 here so that the fold path runs
    keys = rng.integers(0, 2**63, 200000, dtype=np.int64).astype(np.uint64)
    keys[:60000] = keys[0]                                            # worst case: every thread hits the same bins
    nthreads = warps * 32
    per_thread = -(-keys.size // nthreads)
    padded = np.concatenate([keys, np.zeros(per_thread * nthreads - keys.size, dtype=np.uint64)])
    valid = np.arange(padded.size) < keys.size
    g_hist = np.zeros((7, 256), dtype=np.int64)
    col = np.zeros((7, 128 * 32), dtype=np.uint32)
    since = 0

    def flush():
        for d in range(7):
            words = col[d].reshape(128, 32)
            g_hist[d, 0::2] += (words & 0xffff).sum(axis=1).astype(np.int64)
            g_hist[d, 1::2] += (words >> 16).sum(axis=1).astype(np.int64)
        col[:] = 0

    lane_of_thread = np.arange(nthreads) % 32
    for it in range(per_thread):                                      # one key per thread per round, all threads in lock step
        k = padded[it * nthreads:(it + 1) * nthreads]
        ok = valid[it * nthreads:(it + 1) * nthreads]
        for d in range(7):
            b = ((k >> np.uint64(8 * d)) & np.uint64(0xff)).astype(np.int64)
            idx = (b >> 1) * 32 + lane_of_thread
            inc = (1 << ((b & 1) << 4)).astype(np.uint32)
            before = col[d].copy()
            np.add.at(col[d], idx[ok], inc[ok])
            # no carry from the low half into the high half: each half changed by less than 2^16 - its old value
            assert np.all((col[d] & 0xffff) >= (before & 0xffff))
        since += 1
        if since + 8 > flush_keys:
            flush(); since = 0
    flush()
    for d in range(7):
        want = np.bincount(((keys >> np.uint64(8 * d)) & np.uint64(0xff)).astype(np.int64), minlength=256)
        assert np.array_equal(g_hist[d], want), d
    # the bound the kernel's flush interval rests on: 32 warps x 2040 keys < 2^16
    assert 32 * 2040 < 1 << 16
