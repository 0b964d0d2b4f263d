// DIFFUSION3DPA-B200.cpp -- Base_B200 variant (the analogue of apps/DIFFUSION3DPA-Cuda.cpp:25-130).
// symmetric = true as in the reference's kernel body (DIFFUSION3DPA.hpp:231).
#include "Apps.hpp"

namespace rajaperf {
namespace apps {

void DIFFUSION3DPA::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_diffusion3dpa(ctx(), m_B, m_G, m_D, m_X, m_Y, m_NE, 1, s), "rpb200_diffusion3dpa");
}

void DIFFUSION3DPA::runB200Variant(VariantID, size_t) { runRepLoop(); }

}  // namespace apps
}  // namespace rajaperf
