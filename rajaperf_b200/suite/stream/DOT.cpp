// DOT.cpp -- Stream_DOT: sizes, synthetic inputs, checksum (reference: stream/DOT.cpp).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

DOT::DOT(const RunParams& params) : KernelBase(rajaperf::Stream_DOT, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(2000);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) + 2 * sizeof(Real_type) * getActualProblemSize());   // DOT.cpp:31-33: + the running dot
  setBytesWrittenPerRep(1 * sizeof(Real_type));
  setFLOPsPerRep(2 * getActualProblemSize());

  setVariantDefined(Base_B200);
}

void DOT::setUp(VariantID, size_t)      // DOT.cpp:64-69: a @0 -> 0.2, b @1 -> 0.1
{
  allocAndInitData(m_a, getActualProblemSize());
  allocAndInitData(m_b, getActualProblemSize());
  m_dot = 0.0;
  m_dot_init = 0.0;
}

void DOT::updateChecksum(VariantID vid, size_t tune_idx) { checksum[vid][tune_idx] += m_dot; }   // DOT.cpp:80

void DOT::tearDown(VariantID, size_t) { deallocData(m_a); deallocData(m_b); }

}  // namespace stream
}  // namespace rajaperf
