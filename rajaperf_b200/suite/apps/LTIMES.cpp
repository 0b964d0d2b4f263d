// LTIMES.cpp -- Apps_LTIMES (reference: apps/LTIMES.cpp:23-107).
#include <cmath>

#include "Apps.hpp"

namespace rajaperf {
namespace apps {

LTIMES::LTIMES(const RunParams& params) : KernelBase(rajaperf::Apps_LTIMES, params)
{
  m_num_d = params.getLtimesNumD();      // 64
  m_num_g = params.getLtimesNumG();      // 32
  m_num_m = params.getLtimesNumM();      // 25
  const Index_type dg = m_num_d * m_num_g;
  // LTIMES.cpp:23-41: default num_z = round(1e6 / (d*g)); the problem size is psi's length
  m_num_z_default = std::max((Index_type(1000000) + dg / 2) / dg, Index_type(1));
  setDefaultProblemSize(dg * m_num_z_default);
  setDefaultReps(50);
  m_num_z = std::max((getTargetProblemSize() + dg / 2) / dg, Index_type(1));
  m_philen = m_num_m * m_num_g * m_num_z;
  m_elllen = m_num_d * m_num_m;
  m_psilen = dg * m_num_z;
  setActualProblemSize(m_psilen);
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  setBytesReadPerRep((m_philen + m_elllen + m_psilen) * sizeof(Real_type));
  setBytesWrittenPerRep(m_philen * sizeof(Real_type));
  setFLOPsPerRep(2 * m_num_z * m_num_g * m_num_m * m_num_d);
  // LTIMES.cpp:52-54
  checksum_scale_factor = 0.001 * (static_cast<Checksum_type>(getDefaultProblemSize()) / getActualProblemSize());
  setVariantDefined(Base_B200);
}

void LTIMES::setUp(VariantID, size_t)      // LTIMES.cpp:90-92: phi = 0 (@0), ell @1 -> 0.1, psi @2 -> 0.2
{
  allocAndInitDataConst(m_phidat, m_philen, 0.0);
  allocAndInitData(m_elldat, m_elllen);
  allocAndInitData(m_psidat, m_psilen);
}

void LTIMES::updateChecksum(VariantID vid, size_t tune_idx)
{
  checksum[vid][tune_idx] += calcChecksum(m_phidat, m_philen, static_cast<Real_type>(checksum_scale_factor));
}

void LTIMES::tearDown(VariantID, size_t) { deallocData(m_phidat); deallocData(m_elldat); deallocData(m_psidat); }

}  // namespace apps
}  // namespace rajaperf
