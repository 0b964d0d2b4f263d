// SCAN-B200.cpp -- Base_B200 variant (the analogue of algorithm/SCAN-Cuda.cpp:34-188): one single-pass
// launch per rep, no per-rep memset of look-back flags, no scratch to allocate here.
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

void SCAN::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_scan_exclusive(ctx(), m_x, m_y, getActualProblemSize(), s), "rpb200_scan_exclusive");
}

void SCAN::runB200Variant(VariantID, size_t)
{
  checkAbi(rpb200_scan_reserve(ctx(), getActualProblemSize()), "rpb200_scan_reserve");   // scratch outside the timer
  runRepLoop();
}

}  // namespace algorithm
}  // namespace rajaperf
