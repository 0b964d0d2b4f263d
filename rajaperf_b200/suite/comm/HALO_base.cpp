// HALO_base.cpp -- grid dimensions and index-list set-up shared by the Comm kernels
// (reference: comm/HALO_base.cpp:24-67; the lists themselves are built by rpb200_halo_plan_create,
// which restates HALO_base.cpp:118-291 on the device).
#include <cmath>

#include "Comm.hpp"

namespace rajaperf {
namespace comm {

HALO_base::HALO_base(KernelID kid, const RunParams& params) : KernelBase(kid, params)
{
  setDefaultProblemSize(100 * 100 * 100);
  int64_t dims[3];
  rpb200_halo_grid_dims(getTargetProblemSize(), dims);     // HALO_base.cpp:31-35
  for (int d = 0; d < 3; ++d) m_grid_dims[d] = dims[d];
  m_halo_width = params.getHaloWidth();
  m_num_vars = params.getHaloNumVars();
  m_var_size = (m_grid_dims[0] + 2 * m_halo_width) * (m_grid_dims[1] + 2 * m_halo_width) * (m_grid_dims[2] + 2 * m_halo_width);
  setActualProblemSize(m_grid_dims[0] * m_grid_dims[1] * m_grid_dims[1]);   // sic: HALO_base.cpp:46 uses dims[1] twice
  // halo elements per variable: the full box minus the interior that is not sent
  const Index_type w = m_halo_width;
  m_halo_elems = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        if (!dx && !dy && !dz) continue;
        m_halo_elems += (dx ? w : m_grid_dims[0]) * (dy ? w : m_grid_dims[1]) * (dz ? w : m_grid_dims[2]);
      }
}

rpb200_halo_plan* HALO_base::setUp_base(rpb200_ctx* c, int rank, const int* rank_dims)
{
  rpb200_halo_plan* plan = nullptr;
  const int64_t dims[3] = {m_grid_dims[0], m_grid_dims[1], m_grid_dims[2]};
  checkAbi(rpb200_halo_plan_create(c, dims, m_halo_width, (int)m_num_vars, rank, rank_dims, &plan), "rpb200_halo_plan_create");
  for (int l = 0; l < 2 * s_num_neighbors; ++l) detail::incDataInitCount();
  return plan;
}

}  // namespace comm
}  // namespace rajaperf
