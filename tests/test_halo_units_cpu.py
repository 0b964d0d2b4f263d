"""CPU: the unit lists the one-launch halo kernels walk (csrc/halo.cu: build_items_merged / build_items_xchg), through the
host-only hook rpb200_debug_halo_units -- no GPU is touched.

Checked for HALO_base geometries (the tuples of rpb200_halo_plan_bind: neighbour-major, variable-minor) and for ragged generic
tuples: every (tuple, chunk) of both lists appears EXACTLY once; units are well-formed; the four x-face chunks of a variable
form one unit; the variables of a message share one item (their index list is loaded once); the exchange order keeps every
pack unit before every unpack unit.  (Units are drawn from an atomic ticket by the kernel: any CTA may take any unit, so
"every unit once" is all the traversal needs.)"""
import ctypes

import numpy as np
import pytest

import suite_data as sd

CHUNK = 2048
UNPACK = 1 << 30


def first_of(seg):
    return seg & 0xff


def count_of(seg):
    return (seg >> 8) & 0xff


def units(pack, unpack, order):
    """pack / unpack: lists of (len, strided, msg, var) -> (items [(seg, chunk)], unit_first, n_pack_units)"""
    from rajaperf_b200 import cabi
    lib = cabi.load()

    def cols(t):
        n = len(t)
        return ((ctypes.c_int64 * max(n, 1))(*[x[0] for x in t]), (ctypes.c_int * max(n, 1))(*[x[1] for x in t]),
                (ctypes.c_int * max(n, 1))(*[x[2] for x in t]), (ctypes.c_int * max(n, 1))(*[x[3] for x in t]), n)
    p, u = cols(pack), cols(unpack)
    total = sum(-(-x[0] // CHUNK) for x in pack + unpack)
    items = (ctypes.c_int * (2 * total + 2))()
    first = (ctypes.c_int * (total + 2))()
    ni, nu, npu = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.rpb200_debug_halo_units(p[0], p[1], p[2], p[3], p[4], u[0], u[1], u[2], u[3], u[4], order, items, total + 1, first,
                                     total + 1, ctypes.byref(ni), ctypes.byref(nu), ctypes.byref(npu))
    assert rc == 0
    it = [(items[2 * i], items[2 * i + 1]) for i in range(ni.value)]
    return it, [first[i] for i in range(nu.value + 1)], npu.value


def plan_tuples(dims, hw, nv):
    """(len, strided, msg, var) of the plan's tuples: strided = the list's first two cells are not adjacent (csrc/halo.cu: classify)"""
    pack, unpack = sd.halo_lists(dims, hw)
    def side(lists):
        return [(int(lists[l].size), int(lists[l].size >= 2 and lists[l][1] - lists[l][0] != 1), l, v) for l in range(26) for v in range(nv)]
    return side(pack), side(unpack)


def check(pack, unpack, order):
    items, first, npu = units(pack, unpack, order)
    want = sorted([(s, c) for s, t in enumerate(pack) for c in range(-(-t[0] // CHUNK))] +
                  [(s | UNPACK, c) for s, t in enumerate(unpack) for c in range(-(-t[0] // CHUNK))])
    got = []
    for seg, c in items:                                              # an item moves chunk c of `count` consecutive tuples
        side = pack if not (seg & UNPACK) else unpack
        assert count_of(seg) >= 1
        tuples = side[first_of(seg):first_of(seg) + count_of(seg)]
        assert len({(t[0], t[2]) for t in tuples}) == 1               # same length and message: they share the index list
        got += [((first_of(seg) + j) | (seg & UNPACK), c) for j in range(count_of(seg))]
    assert sorted(got) == want                                        # every (tuple, chunk) exactly once
    assert first[0] == 0 and first[-1] == len(items) and all(a < b for a, b in zip(first, first[1:]))
    if order in (0, 6):                                               # exchange: pack units, then unpack units
        cut = first[npu]
        assert all(not (s & UNPACK) for s, _ in items[:cut]) and all(s & UNPACK for s, _ in items[cut:])
    return items, first, npu


@pytest.mark.parametrize("order", [1, 3, 5, 0, 6])
@pytest.mark.parametrize("dims,hw,nv", [((5, 5, 5), 1, 3), ((30, 30, 30), 2, 2), ((100, 100, 100), 1, 3), ((252, 252, 252), 1, 3)])
def test_plan_unit_lists_cover_every_chunk_once(dims, hw, nv, order):
    pack, unpack = plan_tuples(dims, hw, nv)
    items, first, npu = check(pack, unpack, order)
    xlen = max([t[0] for t in pack if t[1]] + [0])
    if xlen < CHUNK:
        return                                                        # no x units at this size: every unit is one item
    # the x faces: neighbours 0 (-x) and 1 (+x)
    sizes = [b - a for a, b in zip(first, first[1:])]
    if order in (1, 3):
        quads = [items[a:b] for a, b in zip(first, first[1:]) if b - a == 4]
        assert len(quads) == nv * -(-xlen // CHUNK) and set(sizes) <= {1, 4}
        for q in quads:                                               # {pack(-x), pack(+x), unpack(-x), unpack(+x)} of one variable, one chunk
            segs = [first_of(s) for s, _ in q]
            assert [s & UNPACK for s, _ in q] == [0, 0, UNPACK, UNPACK] and len({c for _, c in q}) == 1
            assert [s // nv for s in segs] == [0, 1, 0, 1] and len({s % nv for s in segs}) == 1 and all(count_of(s) == 1 for s, _ in q)
        singles = [items[a] for a, b in zip(first, first[1:]) if b - a == 1]
        assert all(count_of(s) == nv for s, _ in singles)             # every other item carries all the variables of its message
        if order == 3:                                                # x units first
            assert sizes[:len(quads)] == [4] * len(quads)
    else:
        pairs = [items[a:b] for a, b in zip(first, first[1:]) if b - a == 2]
        assert len(pairs) == 2 * nv * -(-xlen // CHUNK) and set(sizes) <= {1, 2}
        for q in pairs:
            assert len({s & UNPACK for s, _ in q}) == 1 and len({c for _, c in q}) == 1
            assert sorted(first_of(s) // nv for s, _ in q) == [0, 1]


def test_512_cubed_counts():
    """BASELINE config 5: every chunk once; the x units are 128 chunks x 3 variables."""
    pack, unpack = [], []
    n = 512
    for l, off in enumerate([(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)] + [(1, 1, 0)] * 12 + [(1, 1, 1)] * 8):
        ln = n ** (3 - sum(1 for o in off if o))
        strided = int(off[0] != 0 and ln >= 2)
        for v in range(3):
            pack.append((ln, strided, l, v)); unpack.append((ln, strided, l, v))
    items, first, _ = check(pack, unpack, 1)
    assert sum(1 for a, b in zip(first, first[1:]) if b - a == 4) == 3 * 128
    # the x units lie in the first 85 % of the list: the launch ends on light units
    last_x = max(u for u, (a, b) in enumerate(zip(first, first[1:])) if b - a == 4)
    assert last_x < 0.9 * (len(first) - 1)


@pytest.mark.parametrize("order", [1, 3, 5, 0, 6])
def test_ragged_generic_tuples(order):
    rng = np.random.default_rng(order)
    lens = [0, 1, 3, CHUNK - 1, CHUNK, CHUNK + 1, 3 * CHUNK + 5, 4 * CHUNK, 4 * CHUNK, 4 * CHUNK, 4 * CHUNK]
    pack = [(int(n), int(i >= 7), i % 5, i % 2) for i, n in enumerate(lens)]
    unpack = [(int(n), int(rng.integers(0, 2)), i % 3, i % 2) for i, n in enumerate(reversed(lens))]
    check(pack, unpack, order)
    check(pack, [], order)
    check([], unpack, order)


def test_random_plan_geometries_every_chunk_once_in_every_order():
    """Seeded random boxes (non-cubic, thin, wider halos, 1-4 variables -- 4 x 26 tuples is the most the one-launch kernels keep in
    shared memory; beyond that the library falls back to two launches): the property the kernels' traversal needs."""
    rng = np.random.default_rng(20261018)
    seen_x_units = 0
    for _ in range(30):
        hw = int(rng.choice([1, 1, 1, 2, 3]))
        dims = tuple(int(x) for x in rng.choice([hw, 2 * hw, 7, 33, 64, 129, 200], size=3))
        nv = int(rng.integers(1, 5))
        pack, unpack = plan_tuples(dims, hw, nv)
        for order in (1, 3, 5, 0, 6):
            items, first, _ = check(pack, unpack, order)
            if order == 1:
                seen_x_units += sum(1 for a, b in zip(first, first[1:]) if b - a == 4)
    assert seen_x_units > 0                                           # some of the boxes were big enough to have x units
