#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

void MEMCPY::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_stream_copy(ctx(), m_y, m_x, getActualProblemSize(), s), "rpb200_stream_copy");
}
void MEMCPY::runB200Variant(VariantID, size_t) { runRepLoop(); }

void MEMSET::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_memset_f64(ctx(), m_x, m_val, getActualProblemSize(), s), "rpb200_memset_f64");
}
void MEMSET::runB200Variant(VariantID, size_t) { runRepLoop(); }

}  // namespace algorithm
}  // namespace rajaperf
