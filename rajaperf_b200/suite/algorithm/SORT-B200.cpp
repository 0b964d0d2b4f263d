// SORT-B200.cpp -- Base_B200 variant (the analogue of algorithm/SORT-Cuda.cpp:35-43, which calls
// RAJA::sort -> cub::DeviceRadixSort).  Scratch comes from one allocation outside the timer instead
// of RAJA's per-call pool malloc/free (tpl/RAJA/include/RAJA/policy/cuda/sort.hpp:103-141).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

void SORT::enqueueRep(rpb200_stream_t s)
{
  const Index_type n = getActualProblemSize();
  checkAbi(rpb200_sort_keys_f64(ctx(), m_x + n * m_rep, n, m_scratch, m_scratch_bytes, s), "rpb200_sort_keys_f64");
  ++m_rep;
}

void SORT::runB200Variant(VariantID, size_t)
{
  m_scratch_bytes = rpb200_sort_scratch_bytes(getActualProblemSize(), 0);
  checkAbi(rpb200_malloc(&m_scratch, m_scratch_bytes), "rpb200_malloc");
  m_rep = 0;
  runRepLoop();
  checkAbi(rpb200_free(m_scratch), "rpb200_free");
  m_scratch = nullptr;
}

}  // namespace algorithm
}  // namespace rajaperf
