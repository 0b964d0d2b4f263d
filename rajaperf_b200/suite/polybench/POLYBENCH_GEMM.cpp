#include <cmath>

#include "Polybench.hpp"

namespace rajaperf {
namespace polybench {

POLYBENCH_GEMM::POLYBENCH_GEMM(const RunParams& params) : KernelBase(rajaperf::Polybench_GEMM, params)
{
  const Index_type ni_default = 1000, nj_default = 1000, nk_default = 1200;
  setDefaultProblemSize(ni_default * nj_default);
  setDefaultReps(4);
  // POLYBENCH_GEMM.cpp:30-32: square C, nk = 1.2 ni, both truncated
  m_ni = std::sqrt(getTargetProblemSize()) + std::sqrt(2) - 1;
  m_nj = m_ni;
  m_nk = Index_type(double(nk_default) / ni_default * m_ni);
  m_alpha = 0.62;
  m_beta = 1.002;
  setActualProblemSize(m_ni * m_nj);
  setItsPerRep(m_ni * m_nj);
  setKernelsPerRep(1);
  setBytesReadPerRep(1 * sizeof(Real_type) * m_ni * m_nk + 1 * sizeof(Real_type) * m_nj * m_nk);
  setBytesWrittenPerRep(1 * sizeof(Real_type) * m_ni * m_nj);
  setFLOPsPerRep((1 + 3 * m_nk) * m_ni * m_nj);            // POLYBENCH_GEMM.cpp:45-46 (alpha*A*B + add counted as 3)
  checksum_scale_factor = 0.001 * (static_cast<Checksum_type>(getDefaultProblemSize()) / getActualProblemSize());
  setVariantDefined(Base_B200);
}

void POLYBENCH_GEMM::setUp(VariantID, size_t)               // POLYBENCH_GEMM.cpp:85-89
{
  allocAndInitData(m_A, m_ni * m_nk);
  allocAndInitData(m_B, m_nk * m_nj);
  allocAndInitDataConst(m_C, m_ni * m_nj, 0.0);
}

void POLYBENCH_GEMM::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_C, m_ni * m_nj, static_cast<Real_type>(checksum_scale_factor));
}

void POLYBENCH_GEMM::tearDown(VariantID, size_t) { deallocData(m_A); deallocData(m_B); deallocData(m_C); }

}  // namespace polybench
}  // namespace rajaperf
