#include "Polybench.hpp"

namespace rajaperf {
namespace polybench {

void POLYBENCH_GEMM::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_polybench_gemm(ctx(), m_A, m_B, m_C, m_ni, m_nj, m_nk, m_alpha, m_beta, s), "rpb200_polybench_gemm");
}

void POLYBENCH_GEMM::runB200Variant(VariantID, size_t) { runRepLoop(); }

}  // namespace polybench
}  // namespace rajaperf
