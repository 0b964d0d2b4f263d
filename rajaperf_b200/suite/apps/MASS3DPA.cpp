// MASS3DPA.cpp -- Apps_MASS3DPA: sizes, inputs, checksum (reference: apps/MASS3DPA.cpp:23-100).
#include <cmath>

#include "Apps.hpp"

namespace rajaperf {
namespace apps {

MASS3DPA::MASS3DPA(const RunParams& params) : KernelBase(rajaperf::Apps_MASS3DPA, params)
{
  setDefaultProblemSize(m_NE_default * Q1D * Q1D * Q1D);
  setDefaultReps(50);
  // MASS3DPA.cpp:31: NE = max(round(target / Q1D^3), 1)
  m_NE = std::max((getTargetProblemSize() + (Q1D * Q1D * Q1D) / 2) / (Q1D * Q1D * Q1D), Index_type(1));
  setActualProblemSize(m_NE * Q1D * Q1D * Q1D);
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  // MASS3DPA.cpp:38-49: B, Bt, D, X read; Y read and written
  setBytesReadPerRep(2 * Q1D * D1D * sizeof(Real_type) + Q1D * Q1D * Q1D * m_NE * sizeof(Real_type) +
                     2 * D1D * D1D * D1D * m_NE * sizeof(Real_type));
  setBytesWrittenPerRep(D1D * D1D * D1D * m_NE * sizeof(Real_type));
  setFLOPsPerRep(m_NE * (2 * D1D * D1D * D1D * Q1D + 2 * D1D * D1D * Q1D * Q1D + 2 * D1D * Q1D * Q1D * Q1D + Q1D * Q1D * Q1D +
                         2 * Q1D * Q1D * Q1D * D1D + 2 * Q1D * Q1D * D1D * D1D + 2 * Q1D * D1D * D1D * D1D + D1D * D1D * D1D));
  setVariantDefined(Base_B200);
}

void MASS3DPA::setUp(VariantID, size_t)    // MASS3DPA.cpp:79-83: everything 1.0, Y = 0
{
  allocAndInitDataConst(m_B, Q1D * D1D, 1.0);
  allocAndInitDataConst(m_Bt, Q1D * D1D, 1.0);
  allocAndInitDataConst(m_D, Q1D * Q1D * Q1D * m_NE, 1.0);
  allocAndInitDataConst(m_X, D1D * D1D * D1D * m_NE, 1.0);
  allocAndInitDataConst(m_Y, D1D * D1D * D1D * m_NE, 0.0);
}

void MASS3DPA::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_Y, D1D * D1D * D1D * m_NE, checksum_scale_factor);
}

void MASS3DPA::tearDown(VariantID, size_t)
{
  deallocData(m_B); deallocData(m_Bt); deallocData(m_D); deallocData(m_X); deallocData(m_Y);
}

}  // namespace apps
}  // namespace rajaperf
