// CONVECTION3DPA-B200.cpp -- Base_B200 variant (the analogue of apps/CONVECTION3DPA-Cuda.cpp:24-127).
#include "Apps.hpp"

namespace rajaperf {
namespace apps {

void CONVECTION3DPA::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_convection3dpa(ctx(), m_B, m_Bt, m_G, m_D, m_X, m_Y, m_NE, s), "rpb200_convection3dpa");
}

void CONVECTION3DPA::runB200Variant(VariantID, size_t) { runRepLoop(); }

// launch shapes of csrc/pa.cu as suite tunings: elements per CTA / threads / ring stages / CTAs per SM
void CONVECTION3DPA::setB200TuningDefinitions(VariantID vid)
{
  addB200Tuning(vid, getDefaultTuningName());               // 8 / 128 / 2 / 5
  addB200Tuning(vid, "elems8_ctas4", 0, -1, 11);            // 8 / 128 / 2 / 4
  addB200Tuning(vid, "elems16_block256", 0, -1, 10);        // 16 / 256 / 2 / 2
}

}  // namespace apps
}  // namespace rajaperf
