// REDUCE_SUM-B200.cpp -- Base_B200 variant (the analogue of algorithm/REDUCE_SUM-Cuda.cpp:31-171).
// One launch per rep, result left on the device; one copy-back per rep batch inside the timer
// (the reference syncs and copies back every rep, REDUCE_SUM-Cuda.cpp:160-162).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

void REDUCE_SUM::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_reduce_sum(ctx(), m_x, getActualProblemSize(), m_sum_init, m_d_sum, s), "rpb200_reduce_sum");
}

void REDUCE_SUM::finishReps()
{
  if (getRunReps() > 0) copyToHost(&m_sum, m_d_sum, sizeof(Real_type));
}

void REDUCE_SUM::runB200Variant(VariantID, size_t)
{
  allocData(m_d_sum, 1);
  runRepLoop();
  deallocData(m_d_sum);
}

}  // namespace algorithm
}  // namespace rajaperf
