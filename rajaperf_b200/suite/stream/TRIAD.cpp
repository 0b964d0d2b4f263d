// TRIAD.cpp -- Stream_TRIAD: sizes, synthetic inputs, checksum (reference: stream/TRIAD.cpp).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

TRIAD::TRIAD(const RunParams& params) : KernelBase(rajaperf::Stream_TRIAD, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(1000);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(2 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(2 * getActualProblemSize());
  // TRIAD.cpp:36-38: formed in Checksum_type, narrowed to Real_type where it is used
  checksum_scale_factor = 0.001 * (static_cast<Checksum_type>(getDefaultProblemSize()) / getActualProblemSize());
  setVariantDefined(Base_B200);
}

void TRIAD::setUp(VariantID, size_t)    // TRIAD.cpp:75-78: a = 0 (@0), b @1 -> 0.1, c @2 -> 0.2, alpha @3 -> 0.1*1.1/1.12345
{
  allocAndInitDataConst(m_a, getActualProblemSize(), 0.0);
  allocAndInitData(m_b, getActualProblemSize());
  allocAndInitData(m_c, getActualProblemSize());
  initData(m_alpha);
}

void TRIAD::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_a, getActualProblemSize(), static_cast<Real_type>(checksum_scale_factor));
}

void TRIAD::tearDown(VariantID, size_t) { deallocData(m_a); deallocData(m_b); deallocData(m_c); }

}  // namespace stream
}  // namespace rajaperf
