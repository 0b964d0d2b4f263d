// COPY-B200.cpp -- Base_B200 variant of Stream_COPY (the analogue of stream/COPY-Cuda.cpp:26-98).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

void COPY::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_stream_copy(ctx(), m_c, m_a, getActualProblemSize(), s), "rpb200_stream_copy");
}

void COPY::runB200Variant(VariantID, size_t) { runRepLoop(); }

void COPY::setB200TuningDefinitions(VariantID vid) { defineElementwiseTunings(*this, vid); }

}  // namespace stream
}  // namespace rajaperf
