// CONVECTION3DPA.cpp -- Apps_CONVECTION3DPA (reference: apps/CONVECTION3DPA.cpp:23-105).
#include <cmath>

#include "Apps.hpp"

namespace rajaperf {
namespace apps {

CONVECTION3DPA::CONVECTION3DPA(const RunParams& params) : KernelBase(rajaperf::Apps_CONVECTION3DPA, params)
{
  setDefaultProblemSize(m_NE_default * Q1D * Q1D * Q1D);
  setDefaultReps(50);
  m_NE = std::max((getTargetProblemSize() + (Q1D * Q1D * Q1D) / 2) / (Q1D * Q1D * Q1D), Index_type(1));
  setActualProblemSize(m_NE * Q1D * Q1D * Q1D);
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  // Basis, tBasis, dBasis, D (VDIM slabs), X read; Y read and written.  The reference's own written
  // count has a '+' for a '*' (CONVECTION3DPA.cpp:41); the intended 8*27*NE is used here.
  setBytesReadPerRep(3 * Q1D * D1D * sizeof(Real_type) + Q1D * Q1D * Q1D * VDIM * m_NE * sizeof(Real_type) +
                     2 * D1D * D1D * D1D * m_NE * sizeof(Real_type));
  setBytesWrittenPerRep(D1D * D1D * D1D * m_NE * sizeof(Real_type));
  setFLOPsPerRep(m_NE * 3683);
  setVariantDefined(Base_B200);
}

void CONVECTION3DPA::setUp(VariantID, size_t)   // CONVECTION3DPA.cpp:83-89: everything 1.0, Y = 0
{
  allocAndInitDataConst(m_B, Q1D * D1D, 1.0);
  allocAndInitDataConst(m_Bt, Q1D * D1D, 1.0);
  allocAndInitDataConst(m_G, Q1D * D1D, 1.0);
  allocAndInitDataConst(m_D, Q1D * Q1D * Q1D * VDIM * m_NE, 1.0);
  allocAndInitDataConst(m_X, D1D * D1D * D1D * m_NE, 1.0);
  allocAndInitDataConst(m_Y, D1D * D1D * D1D * m_NE, 0.0);
}

void CONVECTION3DPA::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_Y, D1D * D1D * D1D * m_NE, checksum_scale_factor);
}

void CONVECTION3DPA::tearDown(VariantID, size_t)
{
  deallocData(m_B); deallocData(m_Bt); deallocData(m_G); deallocData(m_D); deallocData(m_X); deallocData(m_Y);
}

}  // namespace apps
}  // namespace rajaperf
