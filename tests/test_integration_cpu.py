"""rajaperf_b200/integration: the Base_B200 variant applied to a copy of the reference tree.  Runs only where the reference
sources are mounted (the authoring container); the script itself refuses to continue when an anchor does not match, so a
successful run means every edit of INTEGRATION.md sections 1-2 landed exactly once."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
INTEG = os.path.join(ROOT, "rajaperf_b200", "integration")
KERNELS = {"stream": ["ADD", "COPY", "DOT", "MUL", "TRIAD"], "algorithm": ["REDUCE_SUM", "SCAN", "SORT", "SORTPAIRS"],
           "apps": ["MASS3DPA", "DIFFUSION3DPA", "CONVECTION3DPA", "LTIMES"], "comm": ["HALO_PACKING_FUSED", "HALO_EXCHANGE_FUSED"]}


def test_every_stub_calls_only_declared_abi_entry_points():
    """The stubs are the reference-side binding: every rpb200_* name they use is declared in include/rpb200.h."""
    import re
    header = open(os.path.join(ROOT, "include", "rpb200.h")).read()
    declared = set(re.findall(r"\b(rpb200_[a-z0-9_]+)\s*\(", header)) | {"rpb200_ctx", "rpb200_stream_t", "rpb200_halo_seg", "rpb200_halo_worklist", "rpb200_halo_plan"}
    n = 0
    for group, names in KERNELS.items():
        for k in names:
            src = open(os.path.join(INTEG, "src", group, f"{k}-B200.cpp")).read()
            assert f"void {k}::runB200Variant(VariantID vid" in src and "startTimer();" in src and "stopTimer();" in src
            used = set(re.findall(r"\b(rpb200_[a-z0-9_]+)\b", src))
            assert used <= declared, (k, used - declared)
            n += 1
    assert n == 15


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference sources are not mounted here")
def test_apply_script_patches_a_copy_of_the_reference(tmp_path):
    dst = tmp_path / "ref_b200"
    r = subprocess.run([sys.executable, os.path.join(INTEG, "apply_base_b200.py"), "--src", REF, "--dst", str(dst)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    common = dst / "src" / "common"
    hpp = (common / "RAJAPerfSuite.hpp").read_text()
    assert hpp.index("Base_Seq = 0") < hpp.index("RAJA_CUDA,") < hpp.index("Base_B200,") < hpp.index("NumVariants")
    names = (common / "RAJAPerfSuite.cpp").read_text()
    assert names.index('std::string("RAJA_CUDA")') < names.index('std::string("Base_B200")') < names.index('std::string("Base_HIP")')
    kb = (common / "KernelBase.cpp").read_text()
    assert kb.count("case Base_B200 :") == 5 and "runB200Variant(vid, tune_idx);" in kb and "setB200TuningDefinitions(vid);" in kb
    assert "running_variant == Base_B200" in (common / "KernelBase.hpp").read_text()
    for group, ks in KERNELS.items():
        cm = (dst / "src" / group / "CMakeLists.txt").read_text()
        for k in ks:
            assert f"{k}-B200.cpp" in cm and (dst / "src" / group / f"{k}-B200.cpp").exists()
            assert "setVariantDefined( Base_B200 );" in (dst / "src" / group / f"{k}.cpp").read_text()
            assert "void runB200Variant(VariantID vid, size_t tune_idx);" in (dst / "src" / group / f"{k}.hpp").read_text()
    # nothing was written to the mounted reference
    assert "Base_B200" not in open(os.path.join(REF, "src", "common", "RAJAPerfSuite.hpp")).read()
