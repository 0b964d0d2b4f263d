// TRIAD-B200.cpp -- Base_B200 variant of Stream_TRIAD (the analogue of stream/TRIAD-Cuda.cpp:24-100).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

void TRIAD::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_stream_triad(ctx(), m_a, m_b, m_c, m_alpha, getActualProblemSize(), s), "rpb200_stream_triad");
}

void TRIAD::runB200Variant(VariantID, size_t) { runRepLoop(); }

void TRIAD::setB200TuningDefinitions(VariantID vid) { defineElementwiseTunings(*this, vid); }

}  // namespace stream
}  // namespace rajaperf
