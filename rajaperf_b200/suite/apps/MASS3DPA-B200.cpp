// MASS3DPA-B200.cpp -- Base_B200 variant (the analogue of apps/MASS3DPA-Cuda.cpp:25-112).
#include "Apps.hpp"

namespace rajaperf {
namespace apps {

void MASS3DPA::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_mass3dpa(ctx(), m_B, m_Bt, m_D, m_X, m_Y, m_NE, s), "rpb200_mass3dpa");
}

void MASS3DPA::runB200Variant(VariantID, size_t) { runRepLoop(); }

}  // namespace apps
}  // namespace rajaperf
