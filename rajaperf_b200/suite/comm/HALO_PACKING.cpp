// HALO_PACKING.cpp -- the unfused pack / unpack kernel (comm/HALO_PACKING.cpp:21-141, HALO_PACKING-Seq.cpp:34-77).
#include "Comm.hpp"

namespace rajaperf {
namespace comm {

HALO_PACKING::HALO_PACKING(const RunParams& params) : HALO_PACKING_FUSED(rajaperf::Comm_HALO_PACKING, params)
{
  setKernelsPerRep(2 * s_num_neighbors * m_num_vars);       // HALO_PACKING.cpp:27
}

void HALO_PACKING::setUp(VariantID vid, size_t tune_idx)
{
  HALO_PACKING_FUSED::setUp(vid, tune_idx);                 // identical data (HALO_PACKING.cpp:62-110)
  m_pack_wl.assign((size_t)s_num_neighbors * m_num_vars, nullptr);
  m_unpack_wl.assign((size_t)s_num_neighbors * m_num_vars, nullptr);
  for (int l = 0; l < s_num_neighbors; ++l) {
    const int* pl = nullptr; const int* ul = nullptr;
    checkAbi(rpb200_halo_plan_neighbor(m_plan, l, nullptr, nullptr, nullptr, nullptr, nullptr, &pl, &ul), "rpb200_halo_plan_neighbor");
    for (Index_type v = 0; v < m_num_vars; ++v) {
      rpb200_halo_seg seg;
      seg.var = m_vars[v]; seg.msg = l; seg.flags = 0;
      seg.buffer = m_pack_buffers[l] + v * m_pack_lens[l]; seg.list = pl; seg.len = m_pack_lens[l];
      checkAbi(rpb200_halo_worklist_create(ctx(), &seg, 1, &m_pack_wl[(size_t)l * m_num_vars + v]), "rpb200_halo_worklist_create");
      seg.buffer = m_unpack_buffers[l] + v * m_unpack_lens[l]; seg.list = ul; seg.len = m_unpack_lens[l];
      checkAbi(rpb200_halo_worklist_create(ctx(), &seg, 1, &m_unpack_wl[(size_t)l * m_num_vars + v]), "rpb200_halo_worklist_create");
    }
  }
}

void HALO_PACKING::tearDown(VariantID vid, size_t tune_idx)
{
  for (rpb200_halo_worklist* w : m_pack_wl) rpb200_halo_worklist_destroy(w);
  for (rpb200_halo_worklist* w : m_unpack_wl) rpb200_halo_worklist_destroy(w);
  m_pack_wl.clear(); m_unpack_wl.clear();
  HALO_PACKING_FUSED::tearDown(vid, tune_idx);
}

void HALO_PACKING::enqueueRep(rpb200_stream_t s)            // neighbour-major, variable-minor, packs then unpacks
{
  for (rpb200_halo_worklist* w : m_pack_wl) checkAbi(rpb200_halo_pack(ctx(), w, s), "rpb200_halo_pack");
  for (rpb200_halo_worklist* w : m_unpack_wl) checkAbi(rpb200_halo_unpack(ctx(), w, s), "rpb200_halo_unpack");
}

}  // namespace comm
}  // namespace rajaperf
