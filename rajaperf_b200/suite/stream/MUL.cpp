// MUL.cpp -- Stream_MUL: sizes, synthetic inputs, checksum (reference: stream/MUL.cpp).
#include "Stream.hpp"

namespace rajaperf {
namespace stream {

MUL::MUL(const RunParams& params) : KernelBase(rajaperf::Stream_MUL, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(1800);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(1 * getActualProblemSize());

  setVariantDefined(Base_B200);
}

void MUL::setUp(VariantID, size_t)      // MUL.cpp:71-73: b = 0 (@0), c @1 -> 0.1, alpha @2 -> 0.2*1.1/1.12345
{
  allocAndInitDataConst(m_b, getActualProblemSize(), 0.0);
  allocAndInitData(m_c, getActualProblemSize());
  initData(m_alpha);
}

void MUL::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_b, getActualProblemSize(), checksum_scale_factor);
}

void MUL::tearDown(VariantID, size_t) { deallocData(m_b); deallocData(m_c); }

}  // namespace stream
}  // namespace rajaperf
