"""GPU parity: Comm group (HALO_base lists, HALO_PACKING_FUSED, HALO_EXCHANGE_FUSED) through the C ABI
vs the CPU oracle and the reference's golden checksums.  Everything here is copies and integer index
work => bit-exact."""
import json
import os

import numpy as np
import pytest

import oracle
import suite_data as sd

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLD = [c for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_checksums.json")))["cases"]
        if c["kernel"] == "Comm_HALO_PACKING_FUSED"]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def dev_list(plan, l, which):
    """Copy a plan-owned device index list back to the host (through the ABI's own memcpy helper)."""
    from rajaperf_b200 import cabi
    nb = plan.neighbors[l]
    n = nb["pack_len" if which == "pack" else "unpack_len"]
    src = nb["d_pack_list" if which == "pack" else "d_unpack_list"]
    host = np.empty(n, dtype=np.int32)
    cabi.check(cabi.load().rpb200_memcpy_d2h(host.ctypes.data, src, 4 * n, None), "memcpy_d2h")
    cabi.check(cabi.load().rpb200_device_synchronize(), "device_synchronize")
    return host


@pytest.mark.parametrize("dims,hw", [((5, 5, 5), 1), ((7, 4, 9), 2), ((30, 30, 30), 3), ((100, 100, 100), 1)])
def test_plan_index_lists_match_oracle(ctx, dims, hw):
    plan = ctx.halo_plan(dims, hw, 1)
    pack, unpack = sd.halo_lists(dims, hw)
    assert plan.var_size == int(np.prod(np.asarray(dims) + 2 * hw))
    for l in range(26):
        assert plan.neighbors[l]["pack_len"] == pack[l].size and plan.neighbors[l]["unpack_len"] == unpack[l].size
        assert np.array_equal(dev_list(plan, l, "pack"), pack[l]), l
        assert np.array_equal(dev_list(plan, l, "unpack"), unpack[l]), l
    plan.close()


@pytest.mark.parametrize("pdims", [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 2), (4, 2, 1)])
def test_plan_rank_grid_matches_oracle(ctx, pdims):
    P = pdims[0] * pdims[1] * pdims[2]
    for rank in range(P):
        plan = ctx.halo_plan((4, 4, 4), 1, 1, rank, pdims)
        r, st, rt = sd.halo_neighbors(rank, pdims)
        assert [nb["rank"] for nb in plan.neighbors] == r.tolist()
        assert [nb["send_tag"] for nb in plan.neighbors] == st.tolist()
        assert [nb["recv_tag"] for nb in plan.neighbors] == rt.tolist()
        plan.close()


def test_grid_dims_follow_the_reference_truncation():
    from rajaperf_b200 import cabi
    L = oracle.lib()
    for target in (1, 27, 1000, 27000, 999999, 1000000, 1000001, 134217728, 1073741824):
        d = np.zeros(3, dtype=np.int64)
        L.orc_halo_grid_dims(target, d)
        assert cabi.halo_grid_dims(target) == d.tolist()


FORMS = ["two_launches", "one_launch", "one_launch_x_first", "one_launch_two_phases", "one_launch_generic"]


def run_packing_fused(ctx, d, reps, form="two_launches"):
    """form: the pack launch + the unpack launch; ONE launch over the merged item list (rpb200_halo_plan_pack_unpack), with
    the x-face items mixed in or first; the same through the generic work-list API (what the reference-side stub calls)."""
    vars_ = [dev(v) for v in d["vars"]]
    pb = [dev(b) for b in d["pack_bufs"]]
    ub = [dev(b) for b in d["unpack_bufs"]]
    plan = ctx.halo_plan(d["dims"], d["hw"], d["nvars"])
    plan.bind(vars_, pb, ub)
    wls = None
    if form == "one_launch_generic":
        nv = d["nvars"]
        mk = lambda bufs, which: ctx.halo_worklist(
            [(bufs[l].data_ptr() + 8 * v * plan.neighbors[l][which + "_len"], plan.neighbors[l]["d_" + which + "_list"],
              vars_[v], plan.neighbors[l][which + "_len"], l) for l in range(26) for v in range(nv)])
        wls = (mk(pb, "pack"), mk(ub, "unpack"))
    if form in ("one_launch_x_first", "one_launch_two_phases"):
        ctx.set_tuning("Comm_HALO_PACKING_FUSED", unroll=3 if form == "one_launch_x_first" else 5)
    try:
        for _ in range(reps):
            if form == "two_launches":
                plan.pack()
                plan.unpack()
            elif form == "one_launch_generic":
                ctx.halo_pack_unpack(*wls)
            else:
                plan.pack_unpack()
        torch.cuda.synchronize()
    finally:
        ctx.reset_tuning("Comm_HALO_PACKING_FUSED")
    out = [v.cpu().numpy() for v in vars_], [b.cpu().numpy() for b in pb]
    if wls:
        wls[0].close(); wls[1].close()
    plan.close()
    return out


def oracle_packing_fused(d, reps):
    L = oracle.lib()
    vars_ = [v.copy() for v in d["vars"]]
    pb = [b.copy() for b in d["pack_bufs"]]
    nv = d["nvars"]
    for _ in range(reps):
        for l in range(26):
            n = d["pack_lists"][l].size
            for v in range(nv):
                L.orc_halo_pack(pb[l][v * n:(v + 1) * n], d["pack_lists"][l], vars_[v], n)
        for l in range(26):
            n = d["unpack_lists"][l].size
            for v in range(nv):
                L.orc_halo_unpack(vars_[v], d["unpack_lists"][l], d["unpack_bufs"][l][v * n:(v + 1) * n], n)
    return vars_, pb


@pytest.mark.parametrize("form", ["two_launches", "one_launch"])
@pytest.mark.parametrize("case", GOLD, ids=lambda c: f"s{c['size']}-r{c['reps']}-{'_'.join(c['flags'])}")
def test_halo_packing_fused_checksum_matches_reference_golden(ctx, case, form):
    f = dict(zip(case["flags"][0::2], case["flags"][1::2]))
    d = sd.halo_packing_fused(case["size"], int(f.get("--halo_width", 1)), int(f.get("--halo_num_vars", 3)))
    vars_, pb = run_packing_fused(ctx, d, case["reps"], form)
    ck = sum(oracle.checksum(v) for v in vars_) + sum(oracle.checksum(b) for b in pb)
    ref = np.longdouble(case["checksum"])
    assert abs(ck - ref) <= abs(ref) * np.longdouble(4e-19), (ck, ref)


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("target,hw,nv", [(1, 1, 1), (27, 1, 2), (1000, 3, 1), (50000, 2, 4), (300000, 1, 3), (16000000, 1, 3)])
def test_halo_packing_fused_bit_exact_vs_oracle(ctx, target, hw, nv, form):
    """(16000000: a 252^3 grid -- the x faces are longer than one chunk, so the X groups of the item list exist)"""
    d = sd.halo_packing_fused(target, hw, nv)
    vars_, pb = run_packing_fused(ctx, d, 2, form)
    rv, rp = oracle_packing_fused(d, 2)
    for a, b in zip(vars_ + pb, rv + rp):
        assert np.array_equal(a.view(np.int64), b.view(np.int64))


def test_generic_worklist_ragged_unaligned_and_empty_segments(ctx):
    """The tuple API on its own: ragged lengths (0, 1, 3, chunk-1, chunk, chunk+1, 3 chunks + 5),
    buffers that are only 8-byte aligned, lists that are only 4-byte aligned, repeated indices."""
    rng = np.random.default_rng(7)
    chunk = ctx.halo_chunk()
    lens = [0, 1, 3, chunk - 1, chunk, chunk + 1, 3 * chunk + 5, 4 * chunk]
    nvar = 100003
    var = rng.standard_normal(nvar)
    d_var = dev(var)
    big_buf = torch.zeros(sum(lens) + 8 * len(lens) + 8, dtype=torch.float64, device="cuda")
    big_list = torch.zeros(sum(lens) + 8 * len(lens) + 8, dtype=torch.int32, device="cuda")
    segs, lists, offs = [], [], []
    ob = ol = 0
    for k, n in enumerate(lens):
        lst = rng.integers(0, nvar, size=n).astype(np.int32)
        ob += (k % 4)          # 8-byte aligned only for most segments
        ol += (k % 3)          # 4-byte aligned only
        big_list[ol:ol + n] = torch.from_numpy(lst).cuda()
        segs.append((big_buf.data_ptr() + 8 * ob, big_list.data_ptr() + 4 * ol, d_var, n, k % 26))
        lists.append(lst); offs.append(ob)
        ob += n; ol += n
    wl = ctx.halo_worklist(segs)
    ctx.halo_pack(wl)
    torch.cuda.synchronize()
    got = big_buf.cpu().numpy()
    for lst, o in zip(lists, offs):
        assert np.array_equal(got[o:o + lst.size].view(np.int64), var[lst].view(np.int64))
    # unpack into a fresh variable: last writer wins only matters for repeated indices, so use a permutation
    perm = rng.permutation(nvar).astype(np.int32)
    d_var2 = torch.zeros(nvar, dtype=torch.float64, device="cuda")
    segs2, pos = [], 0
    src = rng.standard_normal(sum(lens))
    d_src = dev(src)
    d_perm = dev(perm)
    for k, n in enumerate(lens):
        segs2.append((d_src.data_ptr() + 8 * pos, d_perm.data_ptr() + 4 * pos, d_var2, n, 0))
        pos += n
    wl2 = ctx.halo_worklist(segs2)
    ctx.halo_unpack(wl2)
    torch.cuda.synchronize()
    ref = np.zeros(nvar)
    ref[perm[:pos]] = src[:pos]
    assert np.array_equal(d_var2.cpu().numpy().view(np.int64), ref.view(np.int64))
    # update() with the same lengths but another variable
    d_var3 = dev(var * 2.0)
    wl.update([(b, l, d_var3, n, m) for (b, l, _, n, m) in segs])
    ctx.halo_pack(wl)
    torch.cuda.synchronize()
    got = big_buf.cpu().numpy()
    for lst, o in zip(lists, offs):
        assert np.array_equal(got[o:o + lst.size], 2.0 * var[lst])
    wl.close(); wl2.close()


def simulate_exchange(dims, hw, nv, pdims, reps):
    """P ranks of HALO_EXCHANGE_FUSED on the CPU with the oracle's primitives
    (HALO_EXCHANGE_FUSED-Seq.cpp:35-116; delivery as in oracle/rpb_oracle.c)."""
    L = oracle.lib()
    P = pdims[0] * pdims[1] * pdims[2]
    pack, unpack = sd.halo_lists(dims, hw)
    var_size = int(np.prod(np.asarray(dims) + 2 * hw))
    vars_ = [[np.arange(var_size, dtype=np.float64) + v for v in range(nv)] for _ in range(P)]
    nbr = [sd.halo_neighbors(r, pdims) for r in range(P)]
    for _ in range(reps):
        sent = [[np.concatenate([vars_[r][v][pack[l]] for v in range(nv)]) for l in range(26)] for r in range(P)]
        for q in range(P):
            ranks, _, rtags = nbr[q]
            for l in range(26):
                buf = sent[ranks[l]][rtags[l]]
                n = unpack[l].size
                assert buf.size == nv * n
                for v in range(nv):
                    L.orc_halo_unpack(vars_[q][v], unpack[l], np.ascontiguousarray(buf[v * n:(v + 1) * n]), n)
    return vars_


@pytest.mark.parametrize("pdims", [(1, 1, 1), (2, 1, 1), (2, 2, 2), (3, 1, 2)])
@pytest.mark.parametrize("dims,hw,nv", [((6, 6, 6), 1, 3), ((20, 20, 20), 2, 2), ((64, 64, 64), 1, 3)])
def test_halo_exchange_fused_many_ranks_on_one_gpu_bit_exact(ctx, pdims, dims, hw, nv):
    """All ranks of a px*py*pz grid live in this process and share the GPU; their windows are
    connected by plain pointers.  Every rank packs (and signals), then every rank waits + unpacks."""
    P = pdims[0] * pdims[1] * pdims[2]
    reps = 3
    plans, dvars, wins = [], [], []
    for r in range(P):
        plan = ctx.halo_plan(dims, hw, nv, r, pdims)
        vs = [torch.arange(plan.var_size, dtype=torch.float64, device="cuda") + v for v in range(nv)]
        w, nbytes, _ = plan.window(vs, want_handle=False)
        plans.append(plan); dvars.append(vs); wins.append(w)
    for plan in plans:
        plan.connect_ptrs(wins)
    for _ in range(reps):
        for plan in plans:
            plan.exchange_pack()
        for plan in plans:
            plan.exchange_unpack()
    torch.cuda.synchronize()
    ref = simulate_exchange(dims, hw, nv, pdims, reps)
    for r in range(P):
        plans[r].status()
        for v in range(nv):
            assert np.array_equal(dvars[r][v].cpu().numpy().view(np.int64), ref[r][v].view(np.int64)), (r, v)
    for plan in plans:
        plan.close()


@pytest.mark.parametrize("pdims", [(1, 1, 1), (2, 1, 1), (2, 2, 2)])
@pytest.mark.parametrize("dims,hw,nv", [((6, 6, 6), 1, 3), ((20, 20, 20), 2, 2), ((64, 64, 64), 1, 3)])
def test_halo_exchange_unfused_per_tuple_launches_bit_exact(ctx, pdims, dims, hw, nv):
    """The unfused HALO_EXCHANGE: one pack launch and one unpack launch per (neighbour, variable); a rep of
    unfused launches may follow a rep of fused ones (the epoch and the credit counters are shared)."""
    P = pdims[0] * pdims[1] * pdims[2]
    reps = 3
    plans, dvars, wins = [], [], []
    for r in range(P):
        plan = ctx.halo_plan(dims, hw, nv, r, pdims)
        vs = [torch.arange(plan.var_size, dtype=torch.float64, device="cuda") + v for v in range(nv)]
        w, nbytes, _ = plan.window(vs, want_handle=False)
        plans.append(plan); dvars.append(vs); wins.append(w)
    for plan in plans:
        plan.connect_ptrs(wins)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(reps):
        if rep == 1:                       # one fused rep in the middle
            for plan in plans: plan.exchange_pack()
            for plan in plans: plan.exchange_unpack()
            continue
        for plan in plans:
            for l in range(26):
                for v in range(nv):
                    assert plan.lib.rpb200_halo_exchange_pack_seg(plan.h, l, v, st) == 0
        for plan in plans:
            for l in range(26):
                for v in range(nv):
                    assert plan.lib.rpb200_halo_exchange_unpack_seg(plan.h, l, v, 1 if (l == 25 and v == nv - 1) else 0, st) == 0
    torch.cuda.synchronize()
    ref = simulate_exchange(dims, hw, nv, pdims, reps)
    for r in range(P):
        plans[r].status()
        for v in range(nv):
            assert np.array_equal(dvars[r][v].cpu().numpy().view(np.int64), ref[r][v].view(np.int64)), (r, v)
    for plan in plans:
        plan.close()


@pytest.mark.parametrize("pdims", [(1, 1, 1), (2, 1, 1), (2, 2, 2), (3, 1, 2)])
@pytest.mark.parametrize("dims,hw,nv", [((6, 6, 6), 1, 3), ((20, 20, 20), 2, 2), ((64, 64, 64), 1, 3)])
def test_halo_sendrecv_many_ranks_on_one_gpu_bit_exact(ctx, pdims, dims, hw, nv):
    """HALO_SENDRECV: every rank's 26 send buffers land in the receive slot whose recv_tag matches
    (HALO_SENDRECV-Seq.cpp:34-52); rank-specific contents make a misrouted message visible."""
    P = pdims[0] * pdims[1] * pdims[2]
    plans, sends, wins = [], [], []
    for r in range(P):
        plan = ctx.halo_plan(dims, hw, nv, r, pdims)
        w, _, _ = plan.window(None, want_handle=False)                  # transport-only window
        plans.append(plan); wins.append(w)
    for plan in plans:
        plan.connect_ptrs(wins)
    for r, plan in enumerate(plans):
        sb = [torch.arange(nv * nb["pack_len"], dtype=torch.float64, device="cuda") + 1000.0 * l + 1e6 * r
              for l, nb in enumerate(plan.neighbors)]
        plan.sendrecv_bind(sb); sends.append(sb)
    for rep in range(3):
        for r, plan in enumerate(plans):
            for b in sends[r]:
                b += 0.5                                                 # new payload every rep: stale generations would show
        for plan in plans:                      # ranks share one stream here: all puts before the first wait spins
            plan.sendrecv_put()
        for plan in plans:
            plan.sendrecv_wait()
    torch.cuda.synchronize()
    for q, plan in enumerate(plans):
        plan.status()
        ranks, _, rtags = sd.halo_neighbors(q, pdims)
        for l in range(26):
            ptr, n = plan.recv_buffer(l)
            assert n == nv * plan.neighbors[l]["unpack_len"]
            got = np.empty(n)
            import ctypes
            assert plan.lib.rpb200_memcpy_d2h(got.ctypes.data_as(ctypes.c_void_p), ptr, 8 * n, None) == 0
            torch.cuda.synchronize()
            want = sends[ranks[l]][rtags[l]].cpu().numpy()
            assert np.array_equal(got, want), (q, l)
    for plan in plans:
        plan.close()


@pytest.mark.parametrize("size,reps,hw,nv", [(0, 1, 1, 3), (27000, 2, 2, 5)])
def test_halo_sendrecv_suite_checksum_single_rank(ctx, size, reps, hw, nv):
    """KernelBase flow of Comm_HALO_SENDRECV on one rank against the oracle's whole-kernel driver."""
    from rajaperf_b200 import cabi
    dims = cabi.halo_grid_dims(size or 1000000)
    plan = ctx.halo_plan(dims, hw, nv)
    plan.window(None, want_handle=False); plan.connect_ptrs([0])
    L = oracle.lib()
    L.orc_reset_init_count()
    dummy = np.zeros(1)
    for _ in range(52):
        L.orc_init_const(dummy, 0, 0.0)                                  # setUp_base: 52 list allocations
    sb = []
    for nb in plan.neighbors:
        a = np.empty(nv * nb["pack_len"]); L.orc_init_real(a, a.size); sb.append(torch.from_numpy(a).cuda())
    plan.sendrecv_bind(sb)
    for _ in range(reps):
        plan.sendrecv()
    torch.cuda.synchronize(); plan.status()
    ck = np.longdouble(0)
    import ctypes
    for l in range(26):
        ptr, n = plan.recv_buffer(l)
        got = np.empty(n)
        assert plan.lib.rpb200_memcpy_d2h(got.ctypes.data_as(ctypes.c_void_p), ptr, 8 * n, None) == 0
        torch.cuda.synchronize()
        ck += oracle.checksum(got)
    ref = oracle.kat("Comm_HALO_SENDRECV", size, reps, [hw, nv, 1, 1, 1])
    assert abs(ck - ref) <= abs(ref) * np.longdouble(4e-19), (ck, ref)
    plan.close()


def test_halo_exchange_seg_rejects_bad_arguments(ctx):
    plan = ctx.halo_plan((6, 6, 6), 1, 2)
    vs = [torch.zeros(plan.var_size, dtype=torch.float64, device="cuda") for _ in range(2)]
    st = torch.cuda.current_stream().cuda_stream
    assert plan.lib.rpb200_halo_exchange_pack_seg(plan.h, 0, 0, st) != 0          # not connected yet
    plan.window(vs, want_handle=False); plan.connect_ptrs([0])
    for l, v in ((-1, 0), (26, 0), (0, 2), (0, -1)):
        assert plan.lib.rpb200_halo_exchange_pack_seg(plan.h, l, v, st) != 0
        assert plan.lib.rpb200_halo_exchange_unpack_seg(plan.h, l, v, 0, st) != 0
    plan.close()


@pytest.mark.parametrize("size,reps,hw,nv", [(0, 1, 1, 3), (0, 3, 1, 3), (27000, 2, 2, 5)])
def test_halo_exchange_fused_suite_checksum_single_rank(ctx, size, reps, hw, nv):
    """KernelBase flow of Comm_HALO_EXCHANGE_FUSED on one rank (periodic self-exchange) against the
    oracle's whole-kernel driver (the reference cannot build this kernel without MPI)."""
    from rajaperf_b200 import cabi
    dims = cabi.halo_grid_dims(size or 1000000)
    plan = ctx.halo_plan(dims, hw, nv)
    vs = [torch.arange(plan.var_size, dtype=torch.float64, device="cuda") + v for v in range(nv)]
    plan.window(vs, want_handle=False)
    plan.connect_ptrs([0])
    for _ in range(reps):
        plan.exchange()
    torch.cuda.synchronize()
    plan.status()
    ck = sum(oracle.checksum(v.cpu().numpy()) for v in vs)
    ref = oracle.kat("Comm_HALO_EXCHANGE_FUSED", size, reps, [hw, nv, 1, 1, 1])
    assert abs(ck - ref) <= abs(ref) * np.longdouble(4e-19), (ck, ref)
    plan.close()


@pytest.mark.parametrize("unroll", [2, 1, 3], ids=["two_launches", "one_launch", "one_launch_progressive"])
def test_halo_full_size_properties(ctx, unroll):
    """512^3 per GPU (BASELINE config 5): after one exchange every ghost cell equals the periodic
    image of an owned cell; a second exchange is idempotent; owned cells are never modified."""
    n, hw, nv = 512, 1, 3
    ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", unroll=unroll)
    try:
        _halo_full_size_properties(ctx, n, hw, nv)
    finally:
        ctx.reset_tuning("Comm_HALO_EXCHANGE_FUSED")


def _halo_full_size_properties(ctx, n, hw, nv):
    plan = ctx.halo_plan((n, n, n), hw, nv)
    e = n + 2 * hw
    vs = [torch.arange(plan.var_size, dtype=torch.float64, device="cuda") + v for v in range(nv)]
    plan.window(vs, want_handle=False)
    plan.connect_ptrs([0])
    plan.exchange()
    torch.cuda.synchronize()
    plan.status()
    for v in range(nv):
        a = vs[v].view(e, e, e)
        idx = torch.arange(e, device="cuda")
        src = ((idx - hw) % n) + hw                         # periodic image inside the owned box
        want = (src.view(e, 1, 1) * e * e + src.view(1, e, 1) * e + src.view(1, 1, e)).to(torch.float64) + v
        assert torch.equal(a, want)
    before = [v.clone() for v in vs]
    plan.exchange()
    torch.cuda.synchronize()
    for a, b in zip(vs, before):
        assert torch.equal(a, b)
    plan.close()


@pytest.mark.parametrize("unroll", [1, 3], ids=["two_phases", "progressive"])
@pytest.mark.parametrize("dims,hw,nv", [((6, 6, 6), 1, 3), ((20, 20, 20), 2, 2), ((64, 64, 64), 1, 3), ((252, 252, 252), 1, 3)])
def test_halo_exchange_one_launch_form_bit_exact(ctx, dims, hw, nv, unroll):
    """Tunings `unroll` 1 / 3 of Comm_HALO_EXCHANGE_FUSED: the whole rep is ONE launch over the unit list (1: all pack units,
    signal, wait + unpack units; 3: one ticket over both kinds, messages signalled unit by unit).  One rank per GPU only (every CTA of a rank must be resident while its peers pack), so on one GPU
    this is the 1 x 1 x 1 rank grid: 26 self-messages, i.e. the periodic self-exchange of the oracle."""
    reps = 3
    plan = ctx.halo_plan(dims, hw, nv)
    vs = [torch.arange(plan.var_size, dtype=torch.float64, device="cuda") + v for v in range(nv)]
    plan.window(vs, want_handle=False)
    plan.connect_ptrs([0])
    ctx.set_tuning("Comm_HALO_EXCHANGE_FUSED", unroll=unroll)
    try:
        for _ in range(reps):
            plan.exchange()
        torch.cuda.synchronize()
    finally:
        ctx.reset_tuning("Comm_HALO_EXCHANGE_FUSED")
    plan.status()
    ref = simulate_exchange(dims, hw, nv, (1, 1, 1), reps)
    for v in range(nv):
        assert np.array_equal(vs[v].cpu().numpy().view(np.int64), ref[0][v].view(np.int64)), v
    plan.close()


def test_halo_packing_fused_full_size_one_launch_equals_two_launches(ctx):
    """512^3 per GPU, 3 variables (BASELINE config 5): the one-launch form leaves every variable and every pack buffer
    bit-identical to the pack launch followed by the unpack launch; and the pack buffers hold var[list]."""
    n, hw, nv = 512, 1, 3
    plan = ctx.halo_plan((n, n, n), hw, nv)
    f64 = dict(dtype=torch.float64, device="cuda")
    results = []
    for form in ("two", "one"):
        vs = [torch.arange(plan.var_size, **f64) * 0.5 + v for v in range(nv)]
        pb = [torch.zeros(nv * nb["pack_len"], **f64) for nb in plan.neighbors]
        gen = torch.Generator(device="cuda").manual_seed(11)
        ub = [torch.rand(nv * nb["unpack_len"], generator=gen, **f64) for nb in plan.neighbors]
        plan.bind(vs, pb, ub)
        for _ in range(2):
            if form == "two":
                plan.pack(); plan.unpack()
            else:
                plan.pack_unpack()
        torch.cuda.synchronize()
        results.append((vs, pb))
    for a, b in zip(results[0][0] + results[0][1], results[1][0] + results[1][1]):
        assert torch.equal(a, b)
    vs, pb = results[1]
    l = 1                                                       # the +x face: packed from owned cells i = n
    lst = torch.from_numpy(dev_list(plan, l, "pack")).cuda().long()
    for v in range(nv):
        assert torch.equal(pb[l][v * lst.numel():(v + 1) * lst.numel()], vs[v][lst])
    plan.close()
