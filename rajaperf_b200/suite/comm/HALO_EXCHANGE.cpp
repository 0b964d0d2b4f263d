// HALO_EXCHANGE.cpp -- the unfused exchange (comm/HALO_EXCHANGE.cpp:21-150, HALO_EXCHANGE-Seq.cpp:34-116).
#include <cuda_runtime_api.h>

#include "Comm.hpp"

namespace rajaperf {
namespace comm {

HALO_EXCHANGE::HALO_EXCHANGE(const RunParams& params) : HALO_EXCHANGE_FUSED(rajaperf::Comm_HALO_EXCHANGE, params)
{
  setKernelsPerRep(2 * s_num_neighbors * m_num_vars);       // HALO_EXCHANGE.cpp:33
}

void HALO_EXCHANGE::enqueueRep(rpb200_stream_t s)
{
  // all packs of all ranks first (see HALO_EXCHANGE_FUSED::enqueueRep), then the unpacks; neighbour-major,
  // variable-minor like the reference loops; the last unpack launch of a rank commits its epoch
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    for (int l = 0; l < s_num_neighbors; ++l)
      for (Index_type v = 0; v < m_num_vars; ++v)
        checkAbi(rpb200_halo_exchange_pack_seg(rk.plan, l, (int)v, s), "rpb200_halo_exchange_pack_seg");
  }
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    for (int l = 0; l < s_num_neighbors; ++l)
      for (Index_type v = 0; v < m_num_vars; ++v)
        checkAbi(rpb200_halo_exchange_unpack_seg(rk.plan, l, (int)v, (l == s_num_neighbors - 1 && v == m_num_vars - 1) ? 1 : 0, s),
                 "rpb200_halo_exchange_unpack_seg");
  }
  cudaSetDevice(m_first_device);
}

}  // namespace comm
}  // namespace rajaperf
