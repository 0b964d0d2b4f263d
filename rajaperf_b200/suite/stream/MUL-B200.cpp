// MUL-B200.cpp -- Base_B200 variant of Stream_MUL (the analogue of stream/MUL-Cuda.cpp:26-98).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

void MUL::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_stream_mul(ctx(), m_b, m_c, m_alpha, getActualProblemSize(), s), "rpb200_stream_mul");
}

void MUL::runB200Variant(VariantID, size_t) { runRepLoop(); }

void MUL::setB200TuningDefinitions(VariantID vid) { defineElementwiseTunings(*this, vid); }

}  // namespace stream
}  // namespace rajaperf
