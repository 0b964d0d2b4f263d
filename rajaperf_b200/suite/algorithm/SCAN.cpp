// SCAN.cpp -- Algorithm_SCAN (reference: algorithm/SCAN.cpp:21-93).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

SCAN::SCAN(const RunParams& params) : KernelBase(rajaperf::Algorithm_SCAN, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(100);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(1 * getActualProblemSize());
  // SCAN.cpp:36-39: 1e-2 * (default/actual) / actual, formed in Checksum_type
  checksum_scale_factor = 1e-2 * (static_cast<Checksum_type>(getDefaultProblemSize()) / getActualProblemSize()) /
                          getActualProblemSize();
  setVariantDefined(Base_B200);
}

void SCAN::setUp(VariantID, size_t)       // SCAN.cpp:68-71
{
  allocAndInitDataRandValue(m_x, getActualProblemSize());
  allocAndInitDataConst(m_y, getActualProblemSize(), 0.0);
}

void SCAN::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_y, getActualProblemSize(), static_cast<Real_type>(checksum_scale_factor));
}

void SCAN::tearDown(VariantID, size_t) { deallocData(m_x); deallocData(m_y); }

}  // namespace algorithm
}  // namespace rajaperf
