// REDUCE_SUM.cpp -- Algorithm_REDUCE_SUM (reference: algorithm/REDUCE_SUM.cpp:21-84).
#include "Algorithm.hpp"

namespace rajaperf {
namespace algorithm {

REDUCE_SUM::REDUCE_SUM(const RunParams& params) : KernelBase(rajaperf::Algorithm_REDUCE_SUM, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(50);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) * (1 + getActualProblemSize()));                      // REDUCE_SUM.cpp:31-32
  setBytesWrittenPerRep(1 * sizeof(Real_type));
  setFLOPsPerRep(getActualProblemSize());
  setVariantDefined(Base_B200);
}

void REDUCE_SUM::setUp(VariantID, size_t)      // x @0 -> factor 0.2
{
  allocAndInitData(m_x, getActualProblemSize());
  m_sum_init = 0.0;
  m_sum = 0.0;
}

void REDUCE_SUM::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksumHost(&m_sum, 1);      // REDUCE_SUM.cpp:75
}

void REDUCE_SUM::tearDown(VariantID, size_t) { deallocData(m_x); }

}  // namespace algorithm
}  // namespace rajaperf
