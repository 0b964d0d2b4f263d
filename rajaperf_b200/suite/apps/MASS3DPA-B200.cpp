// MASS3DPA-B200.cpp -- Base_B200 variant (the analogue of apps/MASS3DPA-Cuda.cpp:25-112).
#include "Apps.hpp"

namespace rajaperf {
namespace apps {

void MASS3DPA::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_mass3dpa(ctx(), m_B, m_Bt, m_D, m_X, m_Y, m_NE, s), "rpb200_mass3dpa");
}

void MASS3DPA::runB200Variant(VariantID, size_t) { runRepLoop(); }

// launch shapes of csrc/pa.cu as suite tunings: elements per CTA / threads / CTAs per SM / D ring stages
void MASS3DPA::setB200TuningDefinitions(VariantID vid)
{
  addB200Tuning(vid, getDefaultTuningName());               // 8 / 32 / 12 / 1
  addB200Tuning(vid, "elems8_ctas8_ring2", 0, -1, 30);      // 8 / 32 / 8 / 2
  addB200Tuning(vid, "elems16_block64", 0, -1, 10);         // 16 / 64 / 4 / 2
}

}  // namespace apps
}  // namespace rajaperf
