// SORT.cpp -- Algorithm_SORT (reference: algorithm/SORT.cpp:21-75).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

SORT::SORT(const RunParams& params) : KernelBase(rajaperf::Algorithm_SORT, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(20);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) * getActualProblemSize());     // SORT.cpp:31-32: nominal
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getActualProblemSize());
  setFLOPsPerRep(0);
  setVariantDefined(Base_B200);
}

void SORT::setUp(VariantID, size_t)       // SORT.cpp:56: ONE rand() stream over all reps' segments
{
  allocAndInitDataRandValue(m_x, getActualProblemSize() * getRunReps());
}

void SORT::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_x, getActualProblemSize() * getRunReps());
}

void SORT::tearDown(VariantID, size_t) { deallocData(m_x); }

}  // namespace algorithm
}  // namespace rajaperf
