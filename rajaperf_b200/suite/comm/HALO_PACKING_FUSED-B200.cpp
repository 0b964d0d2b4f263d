// HALO_PACKING_FUSED-B200.cpp -- Base_B200 variant (the analogue of comm/HALO_PACKING_FUSED-Cuda.cpp:95-197):
// per rep ONE launch over device-resident tuples (pack and unpack items interleaved: they touch disjoint memory,
// HALO_PACKING_FUSED-Seq.cpp:43-97); no cudaStreamSynchronize (HALO_PACKING_FUSED-Cuda.cpp:149, 187 have one each).
#include "Comm.hpp"

namespace rajaperf {
namespace comm {

void HALO_PACKING_FUSED::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_halo_plan_pack_unpack(m_plan, s), "rpb200_halo_plan_pack_unpack");
}

void HALO_PACKING_FUSED::runB200Variant(VariantID, size_t) { runRepLoop(); }

// Work-list walks of csrc/halo.cu as suite tunings (profiles/r01_halo_variants.md).  The unfused HALO_PACKING shares this
// constructor but has no entry of its own in the library's tuning table: it keeps the default only.
void HALO_PACKING_FUSED::setB200TuningDefinitions(VariantID vid)
{
  addB200Tuning(vid, getDefaultTuningName());               // one launch, x-face items mixed in with the streaming items
  if (getKernelID() != rajaperf::Comm_HALO_PACKING_FUSED) return;
  addB200Tuning(vid, "x_first", 192, 2, 3);                 // one launch, x-face units first
  addB200Tuning(vid, "two_phases", 192, 2, 5);              // one launch: every pack unit, then every unpack unit
  addB200Tuning(vid, "two_launches", 192, 4, 2);            // pack launch + unpack launch, contiguous chunk ranges, packs walk backwards
  addB200Tuning(vid, "two_launches_forward", 256, 4, 2);    // packs walk forward too
  addB200Tuning(vid, "two_launches_round_robin", 128, 4, 2);   // chunks dealt round-robin to the CTAs
}

}  // namespace comm
}  // namespace rajaperf
