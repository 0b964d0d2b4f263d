// COPY.cpp -- Stream_COPY: sizes, synthetic inputs, checksum (reference: stream/COPY.cpp).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

COPY::COPY(const RunParams& params) : KernelBase(rajaperf::Stream_COPY, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(1800);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(0 * getActualProblemSize());

  setVariantDefined(Base_B200);
}

void COPY::setUp(VariantID, size_t)     // COPY.cpp:71-72: a @0 -> 0.2, c = 0
{
  allocAndInitData(m_a, getActualProblemSize());
  allocAndInitDataConst(m_c, getActualProblemSize(), 0.0);
}

void COPY::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_c, getActualProblemSize(), checksum_scale_factor);
}

void COPY::tearDown(VariantID, size_t) { deallocData(m_a); deallocData(m_c); }

}  // namespace stream
}  // namespace rajaperf
