// HALO_PACKING_FUSED.cpp -- Comm_HALO_PACKING_FUSED (reference: comm/HALO_PACKING_FUSED.cpp:17-140).
#include <vector>

#include "Comm.hpp"

namespace rajaperf {
namespace comm {

HALO_PACKING_FUSED::HALO_PACKING_FUSED(KernelID kid, const RunParams& params) : HALO_base(kid, params)
{
  setDefaultReps(200);
  setItsPerRep(m_num_vars * m_halo_elems);                  // HALO_PACKING_FUSED.cpp:26: num_vars x (var_size - owned cells)
  setKernelsPerRep(2);
  // HALO_PACKING_FUSED.cpp:28-35: per packed element an Int_type index + a Real_type read + a Real_type
  // write, once for pack and once for unpack
  setBytesReadPerRep(2 * m_num_vars * m_halo_elems * (sizeof(Int_type) + sizeof(Real_type)));
  setBytesWrittenPerRep(2 * m_num_vars * m_halo_elems * sizeof(Real_type));
  setFLOPsPerRep(0);
  setVariantDefined(Base_B200);
}

void HALO_PACKING_FUSED::setUp(VariantID, size_t)
{
  const int one[3] = {1, 1, 1};
  m_plan = setUp_base(ctx(), 0, one);                      // 52 list allocations: init counter -> 52
  // HALO_PACKING_FUSED.cpp:72-81: vars are allocAndInit'ed (counter bump), then overwritten with i + v
  m_vars.assign(m_num_vars, nullptr);
  std::vector<Real_type> h((size_t)m_var_size);
  for (Index_type v = 0; v < m_num_vars; ++v) {
    allocData(m_vars[v], m_var_size);
    detail::incDataInitCount();
    for (Index_type i = 0; i < m_var_size; ++i) h[i] = i + v;
    copyToDevice(m_vars[v], h.data(), sizeof(Real_type) * (size_t)m_var_size);
  }
  // HALO_PACKING_FUSED.cpp:85-109: 26 pack buffers, then 26 unpack buffers, each num_vars segments,
  // filled by initData (so their factor follows the counter: SURVEY appendix A.1)
  m_pack_buffers.assign(s_num_neighbors, nullptr);
  m_unpack_buffers.assign(s_num_neighbors, nullptr);
  m_pack_lens.assign(s_num_neighbors, 0);
  m_unpack_lens.assign(s_num_neighbors, 0);
  for (int l = 0; l < s_num_neighbors; ++l) {
    int64_t pl = 0, ul = 0;
    checkAbi(rpb200_halo_plan_neighbor(m_plan, l, nullptr, nullptr, nullptr, &pl, &ul, nullptr, nullptr), "rpb200_halo_plan_neighbor");
    m_pack_lens[l] = pl; m_unpack_lens[l] = ul;
  }
  for (int l = 0; l < s_num_neighbors; ++l) allocAndInitData(m_pack_buffers[l], m_num_vars * m_pack_lens[l]);
  for (int l = 0; l < s_num_neighbors; ++l) allocAndInitData(m_unpack_buffers[l], m_num_vars * m_unpack_lens[l]);
  checkAbi(rpb200_halo_plan_bind(m_plan, m_vars.data(), m_pack_buffers.data(), m_unpack_buffers.data()), "rpb200_halo_plan_bind");
}

void HALO_PACKING_FUSED::updateChecksum(VariantID vid, size_t tune_idx)    // HALO_PACKING_FUSED.cpp:112-128
{
  for (Real_ptr var : m_vars) checksum[vid][tune_idx] += calcChecksum(var, m_var_size);
  for (int l = 0; l < s_num_neighbors; ++l)
    checksum[vid][tune_idx] += calcChecksum(m_pack_buffers[l], m_num_vars * m_pack_lens[l]);
}

void HALO_PACKING_FUSED::tearDown(VariantID, size_t)
{
  for (Real_ptr& p : m_pack_buffers) deallocData(p);
  for (Real_ptr& p : m_unpack_buffers) deallocData(p);
  for (Real_ptr& p : m_vars) deallocData(p);
  rpb200_halo_plan_destroy(m_plan);
  m_plan = nullptr;
}

}  // namespace comm
}  // namespace rajaperf
