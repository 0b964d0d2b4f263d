#include "Basic.hpp"

namespace rajaperf {
namespace basic {

void INDEXLIST::enqueueRep(rpb200_stream_t s)
{
  static_assert(sizeof(Index_type) == sizeof(int64_t), "Index_type");
  checkAbi(rpb200_indexlist(ctx(), m_x, m_list, getActualProblemSize(), reinterpret_cast<int64_t*>(m_d_len), s), "rpb200_indexlist");
}

void INDEXLIST::finishReps()         // m_len = count of the last rep (INDEXLIST-Seq.cpp:50): one 8-byte D2H after the batch
{
  copyToHost(&m_len, m_d_len, sizeof(Index_type));
}

void INDEXLIST::runB200Variant(VariantID, size_t)
{
  checkAbi(rpb200_indexlist_reserve(ctx(), getActualProblemSize()), "rpb200_indexlist_reserve");   // scratch outside the timer
  runRepLoop();
}

}  // namespace basic
}  // namespace rajaperf
