"""CPU-side checks of the drop-in boundary: librpb200.so loads without a GPU, exports every symbol
include/rpb200.h declares, the ctypes table covers the header 1:1, and nothing in the product
package touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rpb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rpb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ["rpb200_create", "rpb200_stream_triad", "rpb200_stream_dot", "rpb200_reduce_sum",
                 "rpb200_scan_exclusive", "rpb200_sort_keys_f64", "rpb200_sort_pairs_f64", "rpb200_mass3dpa",
                 "rpb200_diffusion3dpa", "rpb200_convection3dpa", "rpb200_ltimes", "rpb200_halo_pack",
                 "rpb200_halo_unpack", "rpb200_halo_plan_create", "rpb200_halo_exchange"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from rajaperf_b200 import cabi
    assert os.path.exists(cabi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(cabi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from rajaperf_b200 import cabi
    assert sorted(cabi.SIGNATURES) == declared_symbols()
    cabi.load()


def test_host_only_entry_points_work_without_a_gpu():
    from rajaperf_b200 import cabi
    lib = cabi.load()
    assert lib.rpb200_version().decode().startswith("rajaperf-b200")
    assert lib.rpb200_halo_chunk() > 0
    assert cabi.halo_grid_dims(1000000) == [100, 100, 100]
    assert lib.rpb200_sort_scratch_bytes(1 << 20, 0) >= 8 << 20
    assert lib.rpb200_error_string(cabi.ctypes.c_int(-22).value).decode() == "rpb200: invalid argument"


def test_no_gpu_means_create_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rajaperf_b200 import Context, RPB200Error
    with pytest.raises(RPB200Error):
        Context(0)


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under rajaperf_b200/ or include/ may name it."""
    bad = []
    for base in ("rajaperf_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"liboracle|rpb_oracle|orc_[a-z]|oracle/", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
