"""Index arithmetic of the two opt-in kernels that were written after round 1's GPU budget was spent (csrc/scan_tma.cu LINE_ST,
csrc/sort.cu sort_hist_lanes_kernel), restated in numpy: not a test of the CUDA code, a test of the scheme it implements, so
that the first GPU run (tools/time_quick.py scan_line sort_hist) only has to confirm the transcription."""
import numpy as np


def test_four_lane_piece_transpose_gives_every_lane_one_piece_of_four_rows():
    """transpose4_pieces: lane 4G+r holds pieces 0..3 of row r; after two butterfly exchanges lane 4G+q holds piece q of
    rows 0..3, so a warp store instruction j covers the whole 128-byte rows 4G+j."""
    m = [[(lane, p) for p in range(4)] for lane in range(32)]          # m[lane][slot] = (row, piece)

    def exchange(keep_if_set, keep_if_clear, bit, mask):
        send = [m[l][keep_if_clear] if l & bit else m[l][keep_if_set] for l in range(32)]
        for l in range(32):
            if l & bit:
                m[l][keep_if_clear] = send[l ^ mask]
            else:
                m[l][keep_if_set] = send[l ^ mask]

    exchange(1, 0, 1, 1); exchange(3, 2, 1, 1)        # lane ^ 1 on piece bit 0
    exchange(2, 0, 2, 2); exchange(3, 1, 2, 2)        # lane ^ 2 on piece bit 1
    for lane in range(32):
        for j in range(4):
            assert m[lane][j] == ((lane & ~3) + j, lane & 3)
    # the store of instruction j: lane writes 32 bytes at row*128 + piece*32 -> per instruction 8 whole 128-byte lines
    for j in range(4):
        lines = {}
        for lane in range(32):
            row, piece = m[lane][j]
            lines.setdefault(row, set()).add(piece)
        assert len(lines) == 8 and all(v == {0, 1, 2, 3} for v in lines.values())


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
