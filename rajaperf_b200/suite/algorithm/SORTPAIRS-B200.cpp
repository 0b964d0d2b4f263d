// SORTPAIRS-B200.cpp -- Base_B200 variant (the analogue of algorithm/SORTPAIRS-Cuda.cpp:35-43).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

void SORTPAIRS::enqueueRep(rpb200_stream_t s)
{
  const Index_type n = getActualProblemSize();
  checkAbi(rpb200_sort_pairs_f64(ctx(), m_x + n * m_rep, m_i + n * m_rep, n, m_scratch, m_scratch_bytes, s),
           "rpb200_sort_pairs_f64");
  ++m_rep;
}

void SORTPAIRS::runB200Variant(VariantID, size_t)
{
  m_scratch_bytes = rpb200_sort_scratch_bytes(getActualProblemSize(), 1);
  checkAbi(rpb200_malloc(&m_scratch, m_scratch_bytes), "rpb200_malloc");
  m_rep = 0;
  runRepLoop();
  checkAbi(rpb200_free(m_scratch), "rpb200_free");
  m_scratch = nullptr;
}

}  // namespace algorithm
}  // namespace rajaperf
