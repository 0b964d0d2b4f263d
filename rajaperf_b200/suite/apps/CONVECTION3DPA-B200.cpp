// CONVECTION3DPA-B200.cpp -- Base_B200 variant (the analogue of apps/CONVECTION3DPA-Cuda.cpp:24-127).
#include "Apps.hpp"

namespace rajaperf {
namespace apps {

void CONVECTION3DPA::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_convection3dpa(ctx(), m_B, m_Bt, m_G, m_D, m_X, m_Y, m_NE, s), "rpb200_convection3dpa");
}

void CONVECTION3DPA::runB200Variant(VariantID, size_t) { runRepLoop(); }

}  // namespace apps
}  // namespace rajaperf
