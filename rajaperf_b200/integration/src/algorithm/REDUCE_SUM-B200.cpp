//
// REDUCE_SUM-B200.cpp -- the Base_B200 variant of Algorithm_REDUCE_SUM: the analogue of REDUCE_SUM-Cuda.cpp, added to the reference tree by
// rajaperf_b200/integration/apply_base_b200.py.  The kernel is one call into librpb200.so (include/rpb200.h) per rep,
// enqueued on the suite's CUDA stream between startTimer() and stopTimer(); data are the arrays setUp() allocated.
//
#include "REDUCE_SUM.hpp"

#include "RAJA/RAJA.hpp"

#if defined(RAJA_ENABLE_CUDA)

#include "common/B200Utils.hpp"

#include <iostream>

namespace rajaperf
{
namespace algorithm
{

void REDUCE_SUM::runB200Variant(VariantID vid, size_t RAJAPERF_UNUSED_ARG(tune_idx))
{
  const Index_type run_reps = getRunReps();
  auto res{getCudaResource()};
  rpb200_stream_t stream = res.get_stream();
  rpb200_ctx* ctx = getB200Context();

  if ( vid != Base_B200 ) {
    getCout() << "\n  REDUCE_SUM : Unknown B200 variant id = " << vid << std::endl;
    return;
  }

  Real_ptr d_sum = nullptr;
  cudaErrchk( cudaMalloc(reinterpret_cast<void**>(&d_sum), sizeof(Real_type)) );

  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) {
    checkB200( rpb200_reduce_sum(ctx, m_x, getActualProblemSize(), m_sum_init, d_sum, stream), "rpb200_reduce_sum" );
  }
  Real_type sum = m_sum_init;
  cudaErrchk( cudaMemcpyAsync(&sum, d_sum, sizeof(Real_type), cudaMemcpyDeviceToHost, res.get_stream()) );
  cudaErrchk( cudaStreamSynchronize(res.get_stream()) );
  m_sum = sum;                                             // REDUCE_SUM-Seq.cpp:45: the last rep wins
  stopTimer();

  cudaErrchk( cudaFree(d_sum) );
}

} // end namespace algorithm
} // end namespace rajaperf

#endif  // RAJA_ENABLE_CUDA
