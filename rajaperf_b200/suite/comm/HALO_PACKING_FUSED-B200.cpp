// HALO_PACKING_FUSED-B200.cpp -- Base_B200 variant (the analogue of comm/HALO_PACKING_FUSED-Cuda.cpp:95-197):
// per rep one fused pack launch and one fused unpack launch over device-resident tuples; no
// cudaStreamSynchronize between or after them (HALO_PACKING_FUSED-Cuda.cpp:149, 187 have one each).
#include "Comm.hpp"

namespace rajaperf {
namespace comm {

void HALO_PACKING_FUSED::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_halo_plan_pack(m_plan, s), "rpb200_halo_plan_pack");
  checkAbi(rpb200_halo_plan_unpack(m_plan, s), "rpb200_halo_plan_unpack");
}

void HALO_PACKING_FUSED::runB200Variant(VariantID, size_t) { runRepLoop(); }

// Work-list walks of csrc/halo.cu as suite tunings (profiles/r01_halo_variants.md).  The unfused HALO_PACKING shares this
// constructor but has no entry of its own in the library's tuning table: it keeps the default only.
void HALO_PACKING_FUSED::setB200TuningDefinitions(VariantID vid)
{
  addB200Tuning(vid, getDefaultTuningName());               // contiguous chunk ranges, packs walk the list backwards
  if (getKernelID() != rajaperf::Comm_HALO_PACKING_FUSED) return;
  addB200Tuning(vid, "forward", 256, 4, 1);                 // packs walk forward too
  addB200Tuning(vid, "round_robin", 128, 4, 1);             // chunks dealt round-robin to the CTAs
}

}  // namespace comm
}  // namespace rajaperf
