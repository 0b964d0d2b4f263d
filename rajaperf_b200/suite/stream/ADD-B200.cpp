// ADD-B200.cpp -- Base_B200 variant of Stream_ADD (the analogue of stream/ADD-Cuda.cpp:27-100).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

void ADD::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_stream_add(ctx(), m_c, m_a, m_b, getActualProblemSize(), s), "rpb200_stream_add");
}

void ADD::runB200Variant(VariantID, size_t) { runRepLoop(); }

void ADD::setB200TuningDefinitions(VariantID vid) { defineElementwiseTunings(*this, vid); }

}  // namespace stream
}  // namespace rajaperf
