#include "Basic.hpp"

namespace rajaperf {
namespace basic {

INDEXLIST::INDEXLIST(KernelID kid, const RunParams& params) : KernelBase(kid, params)
{
  setDefaultProblemSize(1000000);
  setDefaultReps(100);
  setActualProblemSize(getTargetProblemSize());
  setItsPerRep(getActualProblemSize());
  setKernelsPerRep(1);
  // INDEXLIST.cpp:30-34: the scalar count + x in, the count + "about 50 % output" of Int_type out
  setBytesReadPerRep(1 * sizeof(Index_type) + 1 * sizeof(Real_type) * getActualProblemSize());
  setBytesWrittenPerRep(1 * sizeof(Index_type) + 1 * sizeof(Int_type) * getActualProblemSize() / 2);
  setFLOPsPerRep(0);
  setVariantDefined(Base_B200);
}

INDEXLIST_3LOOP::INDEXLIST_3LOOP(const RunParams& params) : INDEXLIST(rajaperf::Basic_INDEXLIST_3LOOP, params)
{
  // INDEXLIST_3LOOP.cpp:28-40 counts the `counts` temporary's traffic (3 loops); the fused variant moves
  // only x and the list, so the rate below is quoted on the bytes this variant actually needs
  setItsPerRep(3 * getActualProblemSize() + 1);
  setKernelsPerRep(1);
}

void INDEXLIST::setUp(VariantID, size_t)            // INDEXLIST.cpp:62-67
{
  allocAndInitDataRandSign(m_x, getActualProblemSize());
  allocAndInitData(m_list, getActualProblemSize());
  m_len = -1;
  void* p = nullptr;
  checkAbi(rpb200_malloc(&p, sizeof(Index_type)), "rpb200_malloc");
  m_d_len = static_cast<Index_type*>(p);
}

void INDEXLIST::updateChecksum(VariantID vid, size_t tune_idx)     // INDEXLIST.cpp:69-73
{
  checksum[vid][tune_idx] += calcChecksum(m_list, getActualProblemSize());
  checksum[vid][tune_idx] += Checksum_type(m_len);
}

void INDEXLIST::tearDown(VariantID, size_t)
{
  deallocData(m_x);
  deallocData(m_list);
  if (m_d_len) checkAbi(rpb200_free(m_d_len), "rpb200_free");
  m_d_len = nullptr;
}

}  // namespace basic
}  // namespace rajaperf
