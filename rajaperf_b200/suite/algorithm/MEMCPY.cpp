// MEMCPY / MEMSET: the suite's own bandwidth calibration kernels (algorithm/MEMCPY.cpp:21-90, MEMSET.cpp:21-90).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

MEMCPY::MEMCPY(const RunParams& params) : KernelBase(rajaperf::Algorithm_MEMCPY, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(100);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(0);
  setVariantDefined(Base_B200);
}

void MEMCPY::setUp(VariantID, size_t)        // MEMCPY.cpp:59-63: the sentinel must be overwritten everywhere
{
  allocAndInitDataConst(m_x, getActualProblemSize(), 0.0);
  allocAndInitDataConst(m_y, getActualProblemSize(), -1.234567e89);
}
void MEMCPY::updateChecksum(VariantID vid, size_t tune_idx) { checksum[vid][tune_idx] += calcChecksum(m_y, getActualProblemSize()); }
void MEMCPY::tearDown(VariantID, size_t) { deallocData(m_x); deallocData(m_y); }

MEMSET::MEMSET(const RunParams& params) : KernelBase(rajaperf::Algorithm_MEMSET, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(100);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(0);
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(0);
  setVariantDefined(Base_B200);
}

void MEMSET::setUp(VariantID, size_t)        // MEMSET.cpp:59-63
{
  allocAndInitDataConst(m_x, getActualProblemSize(), -1.234567e89);
  m_val = 0.0;
}
void MEMSET::updateChecksum(VariantID vid, size_t tune_idx) { checksum[vid][tune_idx] += calcChecksum(m_x, getActualProblemSize()); }
void MEMSET::tearDown(VariantID, size_t) { deallocData(m_x); }

}  // namespace algorithm
}  // namespace rajaperf
