#include "Polybench.hpp"

namespace rajaperf {
namespace polybench {

void POLYBENCH_GEMM::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_polybench_gemm(ctx(), m_A, m_B, m_C, m_ni, m_nj, m_nk, m_alpha, m_beta, s), "rpb200_polybench_gemm");
}

void POLYBENCH_GEMM::runB200Variant(VariantID, size_t) { runRepLoop(); }

// CTA tilings of csrc/gemm.cu as suite tunings (default: chosen from the problem shape)
void POLYBENCH_GEMM::setB200TuningDefinitions(VariantID vid)
{
  addB200Tuning(vid, getDefaultTuningName());
  addB200Tuning(vid, "tile_64", 64, -1, 0);
  addB200Tuning(vid, "tile_128", 128, -1, 0);
}

}  // namespace polybench
}  // namespace rajaperf
