// SORTPAIRS.cpp -- Algorithm_SORTPAIRS (reference: algorithm/SORTPAIRS.cpp:21-80).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

SORTPAIRS::SORTPAIRS(const RunParams& params) : KernelBase(rajaperf::Algorithm_SORTPAIRS, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(20);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(2 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(2 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(0);
  setVariantDefined(Base_B200);
}

void SORTPAIRS::setUp(VariantID, size_t)  // SORTPAIRS.cpp:56-57: both arrays re-seed => identical contents
{
  allocAndInitDataRandValue(m_x, getActualProblemSize() * getRunReps());
  allocAndInitDataRandValue(m_i, getActualProblemSize() * getRunReps());
}

void SORTPAIRS::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_x, getActualProblemSize() * getRunReps());
  checksum[vid][tune_idx] += calcChecksum(m_i, getActualProblemSize() * getRunReps());
}

void SORTPAIRS::tearDown(VariantID, size_t) { deallocData(m_x); deallocData(m_i); }

}  // namespace algorithm
}  // namespace rajaperf
