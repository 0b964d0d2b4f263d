//
// HALO_EXCHANGE_FUSED-B200.cpp -- the Base_B200 variant of Comm_HALO_EXCHANGE_FUSED: the analogue of
// HALO_EXCHANGE_FUSED-Cuda.cpp, added to the reference tree by rajaperf_b200/integration/apply_base_b200.py.
//
// MPI_Irecv x26 / pack / sync / MPI_Isend x26 / MPI_Waitall / unpack / sync / MPI_Waitall through host-pinned buffers
// (HALO_EXCHANGE_FUSED-Cuda.cpp:109-196) become two launches per rep and no host synchronisation: the pack kernel stores
// every message straight into the receive window of its destination rank over NVLink and releases a flag, the unpack kernel
// acquires the flags (include/rpb200.h, section Comm (3)).  MPI remains for what it is good at: the rendezvous (one
// MPI_Allgather of the 64-byte CUDA IPC handles of the windows) and the timer's barriers.  One rank per process and GPU.
//
#include "HALO_EXCHANGE_FUSED.hpp"

#include "RAJA/RAJA.hpp"

#if defined(RAJA_PERFSUITE_ENABLE_MPI) && defined(RAJA_ENABLE_CUDA)

#include "common/B200Utils.hpp"

#include <iostream>
#include <vector>

namespace rajaperf
{
namespace comm
{

void HALO_EXCHANGE_FUSED::runB200Variant(VariantID vid, size_t RAJAPERF_UNUSED_ARG(tune_idx))
{
  const Index_type run_reps = getRunReps();
  auto res{getCudaResource()};
  rpb200_stream_t stream = res.get_stream();
  rpb200_ctx* ctx = getB200Context();

  if ( vid != Base_B200 ) {
    getCout() << "\n  HALO_EXCHANGE_FUSED : Unknown B200 variant id = " << vid << std::endl;
    return;
  }

  // the plan rebuilds the 26-neighbour decomposition of HALO_base::create_lists for this rank on the device
  const int64_t grid_dims[3] = { m_grid_dims[0], m_grid_dims[1], m_grid_dims[2] };
  const int rank_dims[3] = { m_mpi_dims[0], m_mpi_dims[1], m_mpi_dims[2] };
  rpb200_halo_plan* plan = nullptr;
  checkB200( rpb200_halo_plan_create(ctx, grid_dims, m_halo_width, static_cast<int>(m_num_vars), m_my_mpi_rank, rank_dims, &plan),
             "rpb200_halo_plan_create" );

  std::vector<double*> vars(m_vars.begin(), m_vars.end());
  void* window = nullptr;
  size_t window_bytes = 0;
  unsigned char handle[RPB200_IPC_HANDLE_BYTES];
  checkB200( rpb200_halo_exchange_window(plan, vars.data(), &window, &window_bytes, handle), "rpb200_halo_exchange_window" );
  if (m_mpi_size > 1) {
    std::vector<unsigned char> handles(static_cast<size_t>(RPB200_IPC_HANDLE_BYTES) * m_mpi_size);
    MPI_Allgather(handle, RPB200_IPC_HANDLE_BYTES, MPI_BYTE, handles.data(), RPB200_IPC_HANDLE_BYTES, MPI_BYTE, MPI_COMM_WORLD);
    checkB200( rpb200_halo_exchange_connect(plan, m_mpi_size, handles.data()), "rpb200_halo_exchange_connect" );
    MPI_Barrier(MPI_COMM_WORLD);
  } else {
    void* windows[1] = { window };
    checkB200( rpb200_halo_exchange_connect_ptrs(plan, 1, windows), "rpb200_halo_exchange_connect_ptrs" );
  }

  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) {
    checkB200( rpb200_halo_exchange(plan, stream), "rpb200_halo_exchange" );
  }
  stopTimer();

  checkB200( rpb200_halo_exchange_status(plan), "rpb200_halo_exchange_status (a flag wait timed out)" );
  MPI_Barrier(MPI_COMM_WORLD);                               // nobody unmaps a window a neighbour may still be writing
  rpb200_halo_plan_destroy(plan);
}

} // end namespace comm
} // end namespace rajaperf

#endif  // RAJA_PERFSUITE_ENABLE_MPI && RAJA_ENABLE_CUDA
