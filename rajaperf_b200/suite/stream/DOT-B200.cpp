// DOT-B200.cpp -- Base_B200 variant of Stream_DOT (the analogue of stream/DOT-Cuda.cpp:28-99).
//
// The reference copies the partial result back and synchronises the stream inside every rep
// (DOT-Cuda.cpp:88-89, GPUUtils.hpp:313-319) so that the host can do `m_dot += dot` (DOT-Seq.cpp:45).
// Here the running m_dot lives on the device (`accumulate`): each rep adds m_dot_init + sum a[i]*b[i]
// to it in the same double arithmetic, and ONE copy-back, still inside the timed region, ends the batch.
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

void DOT::enqueueRep(rpb200_stream_t s)
{
  checkAbi(rpb200_stream_dot(ctx(), m_a, m_b, getActualProblemSize(), m_dot_init, m_d_dot, 1, s), "rpb200_stream_dot");
}

void DOT::finishReps()
{
  Real_type batch = 0.0;
  copyToHost(&batch, m_d_dot, sizeof(Real_type));          // the one copy-back + sync, inside the timer
  m_dot += batch;
}

void DOT::runB200Variant(VariantID, size_t)
{
  allocData(m_d_dot, 1);                                   // scratch outside the timer, like DOT-Cuda.cpp:62-70
  checkAbi(rpb200_memset(m_d_dot, 0, sizeof(Real_type), stream()), "rpb200_memset");
  runRepLoop();
  deallocData(m_d_dot);
}

void DOT::setB200TuningDefinitions(VariantID vid)
{
  addB200Tuning(vid, getDefaultTuningName());        // 256 threads, 8 persistent CTAs per SM, 2 vectors per thread per input
  addB200Tuning(vid, "block_512", 512, 4, 2);
}

}  // namespace stream
}  // namespace rajaperf
