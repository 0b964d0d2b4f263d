// HALO_EXCHANGE_FUSED-B200.cpp -- Base_B200 variant (the analogue of comm/HALO_EXCHANGE_FUSED-Cuda.cpp:95-206).
//
//   reference rep:  MPI_Irecv x26 -> pack kernel -> cudaStreamSynchronize -> MPI_Isend x26 ->
//                   MPI_Waitall(recv) -> unpack kernel -> cudaStreamSynchronize -> MPI_Waitall(send)
//   here:           pack kernel that stores straight into the neighbour rank's receive window and
//                   releases a per-message flag  ->  unpack kernel that acquires the flags.
// No host synchronisation inside the rep loop; with several GPUs every rank's launches go to its own
// device's stream and proceed concurrently.
#include <cuda_runtime_api.h>

#include "Comm.hpp"

namespace rajaperf {
namespace comm {

void HALO_EXCHANGE_FUSED::enqueueRep(rpb200_stream_t s)
{
  // every rank packs (and signals), then every rank waits and unpacks: when ranks share a GPU the
  // kernels of one stream run in order, so all packs must be queued before the first unpack spins
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    checkAbi(rpb200_halo_exchange_pack(rk.plan, s), "rpb200_halo_exchange_pack");
  }
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    checkAbi(rpb200_halo_exchange_unpack(rk.plan, s), "rpb200_halo_exchange_unpack");
  }
  cudaSetDevice(m_first_device);
}

void HALO_EXCHANGE_FUSED::runB200Variant(VariantID, size_t)
{
  m_graph_ok = (m_num_devices == 1);     // one capture stream cannot span devices
  runRepLoop();
}

}  // namespace comm
}  // namespace rajaperf
