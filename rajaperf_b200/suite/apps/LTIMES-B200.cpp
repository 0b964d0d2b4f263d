// LTIMES-B200.cpp -- Base_B200 variant (the analogue of apps/LTIMES-Cuda.cpp:44-102).
#include "Apps.hpp"

namespace rajaperf {
namespace apps {

void LTIMES::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_ltimes(ctx(), m_phidat, m_elldat, m_psidat, m_num_d, m_num_g, m_num_m, m_num_z, s), "rpb200_ltimes");
}

void LTIMES::runB200Variant(VariantID, size_t) { runRepLoop(); }

}  // namespace apps
}  // namespace rajaperf
