// DIFFUSION3DPA.cpp -- Apps_DIFFUSION3DPA (reference: apps/DIFFUSION3DPA.cpp:23-104).
#include <cmath>

#include "Apps.hpp"

namespace rajaperf {
namespace apps {

DIFFUSION3DPA::DIFFUSION3DPA(const RunParams& params) : KernelBase(rajaperf::Apps_DIFFUSION3DPA, params)
{
  setDefaultProblemSize(m_NE_default * Q1D * Q1D * Q1D);
  setDefaultReps(50);
  m_NE = std::max((getTargetProblemSize() + (Q1D * Q1D * Q1D) / 2) / (Q1D * Q1D * Q1D), Index_type(1));
  setActualProblemSize(m_NE * Q1D * Q1D * Q1D);
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  // DIFFUSION3DPA.cpp:38-53: Basis, dBasis, D (SYM slabs), X read; Y read and written
  setBytesReadPerRep(2 * Q1D * D1D * sizeof(Real_type) + Q1D * Q1D * Q1D * SYM * m_NE * sizeof(Real_type) +
                     2 * D1D * D1D * D1D * m_NE * sizeof(Real_type));
  setBytesWrittenPerRep(D1D * D1D * D1D * m_NE * sizeof(Real_type));
  // 7065 flop per element: the two basis fills, the d->q contractions (x, y, z), the 15-flop
  // quadrature-point operator, the q->d contractions (z, y, x) and the 3 adds per output dof
  const Index_type d3 = D1D * D1D * D1D, q3 = Q1D * Q1D * Q1D, d2q2 = D1D * D1D * Q1D * Q1D;
  setFLOPsPerRep(m_NE * (2 * Q1D * D1D + 5 * d3 * Q1D + 7 * d2q2 + 7 * D1D * q3 + 15 * q3 +
                         7 * q3 * D1D + 7 * d2q2 + 7 * Q1D * d3 + 3 * d3));
  setVariantDefined(Base_B200);
}

void DIFFUSION3DPA::setUp(VariantID, size_t)   // DIFFUSION3DPA.cpp:83-88: everything 1.0, Y = 0
{
  allocAndInitDataConst(m_B, Q1D * D1D, 1.0);
  allocAndInitDataConst(m_G, Q1D * D1D, 1.0);
  allocAndInitDataConst(m_D, Q1D * Q1D * Q1D * SYM * m_NE, 1.0);
  allocAndInitDataConst(m_X, D1D * D1D * D1D * m_NE, 1.0);
  allocAndInitDataConst(m_Y, D1D * D1D * D1D * m_NE, 0.0);
}

void DIFFUSION3DPA::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_Y, D1D * D1D * D1D * m_NE, checksum_scale_factor);
}

void DIFFUSION3DPA::tearDown(VariantID, size_t)
{
  deallocData(m_B); deallocData(m_G); deallocData(m_D); deallocData(m_X); deallocData(m_Y);
}

}  // namespace apps
}  // namespace rajaperf
