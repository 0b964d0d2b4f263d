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

}  // namespace comm
}  // namespace rajaperf
