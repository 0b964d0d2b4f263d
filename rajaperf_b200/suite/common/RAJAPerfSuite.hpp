// RAJAPerfSuite.hpp -- enums, names and the kernel factory of the Base_B200 build of the suite.
//
// Mirrors common/RAJAPerfSuite.hpp:47-344 of the reference for the hot-path subset: the same GroupID /
// KernelID relative order, and a VariantID enum in which Base_Seq stays 0 (reports take the first
// variant in enum order as the checksum reference, Executor.cpp:1359-1373) and the new Base_B200
// follows the CPU variants.  Names must stay 1:1 with the arrays in RAJAPerfSuite.cpp.
#pragma once
#include <iosfwd>
#include <string>

#include "RPTypes.hpp"

namespace rajaperf {

class KernelBase;
class RunParams;

enum GroupID { Basic = 0, Polybench, Stream, Apps, Algorithm, Comm, NumGroups };   // reference order (RAJAPerfSuite.hpp:47-63)

enum KernelID {
  Basic_INDEXLIST = 0, Basic_INDEXLIST_3LOOP,          // widened rows
  Polybench_GEMM,
  Stream_ADD, Stream_COPY, Stream_DOT, Stream_MUL, Stream_TRIAD,
  Apps_CONVECTION3DPA, Apps_DIFFUSION3DPA, Apps_LTIMES, Apps_MASS3DPA,
  Algorithm_SCAN, Algorithm_SORT, Algorithm_SORTPAIRS, Algorithm_REDUCE_SUM, Algorithm_MEMSET, Algorithm_MEMCPY,
  Comm_HALO_PACKING, Comm_HALO_PACKING_FUSED, Comm_HALO_SENDRECV, Comm_HALO_EXCHANGE, Comm_HALO_EXCHANGE_FUSED,
  NumKernels
};

// Base_Seq / Base_OpenMP keep their slots and names so that -v, report headers and the
// "first variant listed is the reference" rule read like the reference's; in this build only
// Base_B200 is available (the CPU variants live in the reference binary -- there is no CPU path here).
enum VariantID { Base_Seq = 0, Base_OpenMP, Base_B200, NumVariants };

const std::string& getGroupName(GroupID gid);
std::string getKernelName(KernelID kid);             // "TRIAD"
const std::string& getFullKernelName(KernelID kid);  // "Stream_TRIAD"
GroupID getKernelGroup(KernelID kid);
const std::string& getVariantName(VariantID vid);
bool isVariantAvailable(VariantID vid);
bool isVariantGPU(VariantID vid);
KernelBase* getKernelObject(KernelID kid, const RunParams& run_params);
std::ostream& getCout();

}  // namespace rajaperf
