// ADD.cpp -- Stream_ADD: sizes, synthetic inputs, checksum (reference: stream/ADD.cpp).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

ADD::ADD(const RunParams& params) : KernelBase(rajaperf::Stream_ADD, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(1000);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(2 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(1 * getActualProblemSize());

  setVariantDefined(Base_B200);
}

void ADD::setUp(VariantID, size_t)      // ADD.cpp:71-73: a @0 -> 0.2, b @1 -> 0.1, c = 0
{
  allocAndInitData(m_a, getActualProblemSize());
  allocAndInitData(m_b, getActualProblemSize());
  allocAndInitDataConst(m_c, getActualProblemSize(), 0.0);
}

void ADD::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_c, getActualProblemSize(), checksum_scale_factor);
}

void ADD::tearDown(VariantID, size_t) { deallocData(m_a); deallocData(m_b); deallocData(m_c); }

}  // namespace stream
}  // namespace rajaperf
