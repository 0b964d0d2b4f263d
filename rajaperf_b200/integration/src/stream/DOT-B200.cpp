//
// DOT-B200.cpp -- the Base_B200 variant of Stream_DOT: the analogue of DOT-Cuda.cpp, added to the reference tree by
// rajaperf_b200/integration/apply_base_b200.py.  The kernel is one call into librpb200.so (include/rpb200.h) per rep,
// enqueued on the suite's CUDA stream between startTimer() and stopTimer(); data are the arrays setUp() allocated.
//
#include "DOT.hpp"

#include "RAJA/RAJA.hpp"

#if defined(RAJA_ENABLE_CUDA)

#include "common/B200Utils.hpp"

#include <iostream>

namespace rajaperf
{
namespace stream
{

void DOT::runB200Variant(VariantID vid, size_t RAJAPERF_UNUSED_ARG(tune_idx))
{
  const Index_type run_reps = getRunReps();
  auto res{getCudaResource()};
  rpb200_stream_t stream = res.get_stream();
  rpb200_ctx* ctx = getB200Context();

  if ( vid != Base_B200 ) {
    getCout() << "\n  DOT : Unknown B200 variant id = " << vid << std::endl;
    return;
  }

  // the running m_dot of DOT-Seq.cpp:45 lives on the device for the rep batch (same additions, same order): ONE copy-back,
  // inside the timer, replaces the per-rep copy-back + stream synchronisation of DOT-Cuda.cpp:88-89
  Real_ptr d_dot = nullptr;
  cudaErrchk( cudaMalloc(reinterpret_cast<void**>(&d_dot), sizeof(Real_type)) );
  cudaErrchk( cudaMemsetAsync(d_dot, 0, sizeof(Real_type), res.get_stream()) );

  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) {
    checkB200( rpb200_stream_dot(ctx, m_a, m_b, getActualProblemSize(), m_dot_init, d_dot, /*accumulate=*/1, stream),
               "rpb200_stream_dot" );
  }
  Real_type batch = 0.0;
  cudaErrchk( cudaMemcpyAsync(&batch, d_dot, sizeof(Real_type), cudaMemcpyDeviceToHost, res.get_stream()) );
  cudaErrchk( cudaStreamSynchronize(res.get_stream()) );
  m_dot += batch;
  stopTimer();

  cudaErrchk( cudaFree(d_dot) );
}

} // end namespace stream
} // end namespace rajaperf

#endif  // RAJA_ENABLE_CUDA
