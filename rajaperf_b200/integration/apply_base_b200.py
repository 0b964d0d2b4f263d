#!/usr/bin/env python
"""Apply the Base_B200 integration to a COPY of the reference tree -- INTEGRATION.md, executed.

    python rajaperf_b200/integration/apply_base_b200.py --src /root/reference --dst /tmp/ref_b200
    (then: rajaperf_b200/integration/build_ref_b200.sh)

Nothing is written under --src and no reference source is stored in this repo: the script copies the tree to --dst, makes
the six `src/common` edits of INTEGRATION.md section 1 there by anchored text insertion (every anchor must match exactly
once, or exactly the stated number of times -- otherwise the script stops: the reference changed), adds
`setVariantDefined( Base_B200 )` + the `runB200Variant` declaration to the 15 kernel classes, drops in the stubs of
rajaperf_b200/integration/src/ (this repo's own files) and lists them in the group CMakeLists.  The result is the reference's
own driver, reports and checksum comparison with one more variant: `raja-perf.exe -k Stream -v Base_Seq Base_CUDA Base_B200`.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KERNELS = {
    "stream": ["ADD", "COPY", "DOT", "MUL", "TRIAD"],
    "algorithm": ["REDUCE_SUM", "SCAN", "SORT", "SORTPAIRS"],
    "apps": ["MASS3DPA", "DIFFUSION3DPA", "CONVECTION3DPA", "LTIMES"],
    "comm": ["HALO_PACKING_FUSED", "HALO_EXCHANGE_FUSED"],      # the exchange compiles only in an MPI build (it is #if'ed out otherwise)
}


def edit(path, old, new, count=1):
    s = open(path).read()
    n = s.count(old)
    if n != count:
        sys.exit(f"{path}: anchor {old!r} found {n} times, expected {count} -- the reference differs from the one this was written for")
    open(path, "w").write(s.replace(old, new))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default="/tmp/ref_b200")
    a = ap.parse_args()
    if os.path.exists(a.dst):
        shutil.rmtree(a.dst)
    shutil.copytree(a.src, a.dst, symlinks=True, ignore=shutil.ignore_patterns(".git"))
    for root, dirs, files in os.walk(a.dst):                      # the mounted reference is read-only: make the copy writable
        for n in dirs + files:
            p = os.path.join(root, n)
            if not os.path.islink(p):
                os.chmod(p, os.stat(p).st_mode | 0o200)
    src = os.path.join(a.dst, "src")
    common = os.path.join(src, "common")

    # 1. VariantID + VariantNames: right after RAJA_CUDA, so that Base_Seq stays 0 and the CUDA data-space arms can be shared
    edit(os.path.join(common, "RAJAPerfSuite.hpp"), "  RAJA_CUDA,\n", "  RAJA_CUDA,\n  Base_B200,\n")
    edit(os.path.join(common, "RAJAPerfSuite.cpp"), '  std::string("RAJA_CUDA"),\n', '  std::string("RAJA_CUDA"),\n  std::string("Base_B200"),\n')
    # 2. isVariantAvailable / isVariantGPU (both inside #if defined(RAJA_ENABLE_CUDA))
    edit(os.path.join(common, "RAJAPerfSuite.cpp"), "       vid == RAJA_CUDA ) {\n", "       vid == RAJA_CUDA ||\n       vid == Base_B200 ) {\n", 2)
    # 3. KernelBase.hpp: the two virtuals (non-pure: the other ~70 kernels simply do not define the variant), synchronize()
    edit(os.path.join(common, "KernelBase.hpp"), "  virtual void runCudaVariant(VariantID vid, size_t tune_idx) = 0;\n",
         "  virtual void runCudaVariant(VariantID vid, size_t tune_idx) = 0;\n"
         "  virtual void runB200Variant(VariantID vid, size_t /*tune_idx*/)\n"
         "  {\n"
         "    getCout() << \"\\n  \" << getName() << \" : no Base_B200 variant (id \" << vid << \")\" << std::endl;\n"
         "  }\n")
    edit(os.path.join(common, "KernelBase.hpp"), "  virtual void setCudaTuningDefinitions(VariantID vid)\n",
         "  virtual void setB200TuningDefinitions(VariantID vid)\n"
         "  { addVariantTuningName(vid, getDefaultTuningName()); }\n"
         "  virtual void setCudaTuningDefinitions(VariantID vid)\n")
    edit(os.path.join(common, "KernelBase.hpp"), "         running_variant == RAJA_CUDA ) {\n",
         "         running_variant == RAJA_CUDA ||\n         running_variant == Base_B200 ) {\n")
    # 4. KernelBase.cpp: tuning definitions, the three data spaces (device memory: also for the MPI buffers), runKernel
    kb = os.path.join(common, "KernelBase.cpp")
    edit(kb, "    case Base_CUDA :\n    case Lambda_CUDA :\n    case RAJA_CUDA :\n    {\n#if defined(RAJA_ENABLE_CUDA)\n      setCudaTuningDefinitions(vid);\n#endif\n      break;\n    }\n",
         "    case Base_CUDA :\n    case Lambda_CUDA :\n    case RAJA_CUDA :\n    {\n#if defined(RAJA_ENABLE_CUDA)\n      setCudaTuningDefinitions(vid);\n#endif\n      break;\n    }\n\n"
         "    case Base_B200 :\n    {\n#if defined(RAJA_ENABLE_CUDA)\n      setB200TuningDefinitions(vid);\n#endif\n      break;\n    }\n")
    edit(kb, "    case RAJA_CUDA :\n      return run_params.getCudaDataSpace();\n", "    case RAJA_CUDA :\n    case Base_B200 :\n      return run_params.getCudaDataSpace();\n")
    edit(kb, "    case RAJA_CUDA :\n      return run_params.getCudaMPIDataSpace();\n",
         "    case RAJA_CUDA :\n      return run_params.getCudaMPIDataSpace();\n\n    case Base_B200 :\n      return DataSpace::CudaDevice;      // the kernels pack into device buffers\n")
    edit(kb, "    case RAJA_CUDA :\n      return run_params.getCudaReductionDataSpace();\n", "    case RAJA_CUDA :\n    case Base_B200 :\n      return run_params.getCudaReductionDataSpace();\n")
    edit(kb, "    case Base_CUDA :\n    case Lambda_CUDA :\n    case RAJA_CUDA :\n    {\n#if defined(RAJA_ENABLE_CUDA)\n      runCudaVariant(vid, tune_idx);\n#endif\n      break;\n    }\n",
         "    case Base_CUDA :\n    case Lambda_CUDA :\n    case RAJA_CUDA :\n    {\n#if defined(RAJA_ENABLE_CUDA)\n      runCudaVariant(vid, tune_idx);\n#endif\n      break;\n    }\n\n"
         "    case Base_B200 :\n    {\n#if defined(RAJA_ENABLE_CUDA)\n      runB200Variant(vid, tune_idx);\n#endif\n      break;\n    }\n")
    shutil.copy(os.path.join(HERE, "src", "common", "B200Utils.hpp"), os.path.join(common, "B200Utils.hpp"))

    # 5. the kernel classes: define the variant, declare the method, add the stub, list it in CMake
    for group, names in KERNELS.items():
        cm = os.path.join(src, group, "CMakeLists.txt")
        for k in names:
            edit(os.path.join(src, group, f"{k}.cpp"), "  setVariantDefined( RAJA_CUDA );\n", "  setVariantDefined( RAJA_CUDA );\n  setVariantDefined( Base_B200 );\n")
            edit(os.path.join(src, group, f"{k}.hpp"), "  void runCudaVariant(VariantID vid, size_t tune_idx);\n",
                 "  void runCudaVariant(VariantID vid, size_t tune_idx);\n  void runB200Variant(VariantID vid, size_t tune_idx);\n")
            shutil.copy(os.path.join(HERE, "src", group, f"{k}-B200.cpp"), os.path.join(src, group, f"{k}-B200.cpp"))
            s = open(cm).read()
            line = [l for l in s.splitlines(keepends=True) if l.strip() == f"{k}-Cuda.cpp"]
            if len(line) != 1:
                sys.exit(f"{cm}: {k}-Cuda.cpp listed {len(line)} times")
            open(cm, "w").write(s.replace(line[0], line[0] + line[0].replace(f"{k}-Cuda.cpp", f"{k}-B200.cpp")))
    print(f"Base_B200 applied to {a.dst} ({sum(len(v) for v in KERNELS.values())} kernels)")


if __name__ == "__main__":
    main()
