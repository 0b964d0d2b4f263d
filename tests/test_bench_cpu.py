"""bench.py pieces that need no GPU: the reference arm's JSON line (the contract the driver parses), the parser of the
suite's timing report, and the per-kernel cpu_openmp leg -- on tiny samples, against the reference binary when it is here
(oracle/_ref/raja-perf.exe) and the OpenMP port otherwise."""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_suite_timing_report_parser(bench, tmp_path):
    p = tmp_path / "RAJAPerf-timing-Average.csv"
    p.write_text("Mean Runtime Report (sec.)  ,  , \n"
                 "Kernel          , Base_OpenMP , RAJA_OpenMP\n"
                 "Kernel          , default  , default    \n"
                 "Stream_TRIAD    , 0.000358 ,    0.038571\n"
                 "Algorithm_SORT  , Not run ,    0.055839\n")
    got = bench._read_suite_csv(str(p))
    assert got == {"Stream_TRIAD": {"Base_OpenMP-default": 0.000358, "RAJA_OpenMP-default": 0.038571},
                   "Algorithm_SORT": {"RAJA_OpenMP-default": 0.055839}}


def test_reference_arm_prints_the_contract_line(bench, monkeypatch):
    monkeypatch.setattr(bench, "STREAM_N", 1 << 20)
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference_arm(args, rank=0)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "GB/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    # both arms print the SAME config (it names the workload; the implementation is in `arm`)
    assert line["config"] == bench.workload_config(1 << 20, 1) and "Base_OpenMP" in line["arm"]
    assert "Base_B200" not in json.dumps(line["config"]) and "OpenMP" not in json.dumps(line["config"])
    # ranks other than 0 print nothing
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference_arm(args, rank=1)
    assert buf.getvalue() == ""


def test_cpu_openmp_leg_on_tiny_samples(bench, monkeypatch):
    if not os.path.exists(bench.REF_EXE):
        assert bench.cpu_kernels_reference() == {}          # no binary: the column is simply absent
        return
    monkeypatch.setattr(bench, "CPU_KERNEL_SAMPLES", [(["Algorithm_SCAN", "Algorithm_REDUCE_SUM"], 100000, 2),
                                                      (["Algorithm_SORT"], 50000, 1), (["Apps_LTIMES"], 20480, 1)])
    got = bench.cpu_kernels_reference()
    assert set(got) == {"Algorithm_SCAN", "Algorithm_REDUCE_SUM", "Algorithm_SORT", "Apps_LTIMES"}
    for k, v in got.items():
        assert v["ms"] > 0 and v["gbs"] > 0 and v["threads"] >= 1 and v["variant"].split("-")[0] in ("Base_OpenMP", "RAJA_OpenMP")
    assert got["Algorithm_SORT"]["variant"].startswith("RAJA_OpenMP")      # SORT has no Base_OpenMP (SORT.cpp:40-47)
    assert got["Algorithm_SCAN"]["size"] == 100000 and got["Algorithm_SCAN"]["reps"] == 2
    json.dumps(got)                                                         # goes into the bench line as is


def test_measured_peak_falls_back_to_the_profiling_guide_number(bench, monkeypatch, tmp_path):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    peak, src = bench.measured_peak()
    assert peak == 6650.0 and "fallback" in src
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6554.9}))
    peak, src = bench.measured_peak()
    assert peak == 6554.9 and "measured" in src


def test_ncu_traffic_is_refused_unless_it_is_of_the_launch_this_run_makes(bench, monkeypatch, tmp_path):
    """roofline.traffic comes from an ncu capture of the same kernel at the same size and launch shape, or it is null."""
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "r02_ncu_traffic.json").write_text(json.dumps({
        "Stream_TRIAD": {"kernel_name": "stream_ew_kernel<3, 2>", "n": 1 << 28, "tuning": [512, 0, 2],
                         "dram_bytes_per_launch": 6.4e9, "commit": "abc1234"}}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))

    class Ctx:
        def __init__(self, t): self.t = t
        def get_tuning(self, kernel): return self.t

    got, src = bench.ncu_traffic(Ctx((512, 0, 2)), "Stream_TRIAD", 1 << 28)
    assert got == 6.4e9 and "stream_ew_kernel<3, 2>" in src and "abc1234" in src
    got, why = bench.ncu_traffic(Ctx((256, 8, 4)), "Stream_TRIAD", 1 << 28)
    assert got is None and "launch shape" in why
    got, why = bench.ncu_traffic(Ctx((512, 0, 2)), "Stream_TRIAD", 1 << 20)
    assert got is None and "n =" in why
    got, why = bench.ncu_traffic(Ctx((512, 0, 2)), "Stream_COPY", 1 << 28)
    assert got is None and "no ncu capture" in why


def test_device_side_suite_data_is_bit_identical_to_the_oracle(bench):
    """bench.py generates the suite's input arrays on the device with torch arithmetic; the same expressions on CPU tensors must
    reproduce initData (common/DataUtils.cpp:504-513: (factor*(i+1.1))/(i+1.12345)) bit for bit, shards included."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    n = 300001
    for factor, count in ((0.2, 0), (0.1, 1)):
        want = oracle.init_real(n, count)
        got = bench.init_real_dev(torch, n, factor, "cpu").numpy()
        assert np.array_equal(got.view(np.int64), want.view(np.int64))
    # a shard of a larger array (the sharded REDUCE_SUM of the N > 1 path): elements [first, first + m) of the same sequence
    whole = oracle.init_real(n, 0)
    for first, m in ((0, 7), (12345, 100000), (n - 5, 5)):
        got = bench.init_real_dev_range(torch, first, m, 0.2, "cpu").numpy()
        assert np.array_equal(got.view(np.int64), whole[first:first + m].view(np.int64))
    # far into an 8-GPU array (rank 7 of 8 x 2^27): still the closed form, evaluated in double
    first = 7 * (1 << 27) + 3
    got = bench.init_real_dev_range(torch, first, 4, 0.2, "cpu").numpy()
    i = np.arange(first, first + 4, dtype=np.float64)
    assert np.array_equal(got, (0.2 * (i + 1.1)) / (i + 1.12345))
    assert bench.init_scalar(0.2) == (0.2 * 1.1) / 1.12345


@pytest.mark.parametrize("pdims", [(1, 1, 1), (2, 2, 2), (3, 1, 2)])
@pytest.mark.parametrize("g,hw,nv", [(6, 1, 3), (10, 2, 2)])
def test_halo_verification_accepts_the_simulated_exchange_and_rejects_a_wrong_cell(bench, pdims, g, hw, nv):
    """`halo_exchange.verified` rests on ghosts_are_periodic_images(): it must hold for every rank of the CPU simulation of the
    reference's exchange (HALO_EXCHANGE_FUSED-Seq.cpp:35-116) on any rank grid, fail before the exchange, and fail when a
    single ghost or owned cell is off."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_comm_gpu import simulate_exchange
    e = g + 2 * hw
    fresh = [torch.arange(e ** 3, dtype=torch.float64) + v for v in range(nv)]
    assert not bench.ghosts_are_periodic_images(torch, fresh, g, hw, "cpu")          # ghost cells not exchanged yet
    ranks = simulate_exchange((g, g, g), hw, nv, pdims, 2)
    for vs in ranks:
        tv = [torch.from_numpy(v.copy()) for v in vs]
        assert bench.ghosts_are_periodic_images(torch, tv, g, hw, "cpu")
    tv = [torch.from_numpy(v.copy()) for v in ranks[-1]]
    tv[nv - 1][0] += 1.0                                                           # a corner ghost cell
    assert not bench.ghosts_are_periodic_images(torch, tv, g, hw, "cpu")
    tv[nv - 1][0] -= 1.0
    assert bench.ghosts_are_periodic_images(torch, tv, g, hw, "cpu")
    centre = ((e // 2) * e + e // 2) * e + e // 2
    tv[0][centre] = -7.0                                                           # an owned cell
    assert not bench.ghosts_are_periodic_images(torch, tv, g, hw, "cpu")
