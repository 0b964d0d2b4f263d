//
// SORT-B200.cpp -- the Base_B200 variant of Algorithm_SORT: the analogue of SORT-Cuda.cpp, added to the reference tree by
// rajaperf_b200/integration/apply_base_b200.py.  The kernel is one call into librpb200.so (include/rpb200.h) per rep,
// enqueued on the suite's CUDA stream between startTimer() and stopTimer(); data are the arrays setUp() allocated.
//
#include "SORT.hpp"

#include "RAJA/RAJA.hpp"

#if defined(RAJA_ENABLE_CUDA)

#include "common/B200Utils.hpp"

#include <iostream>

namespace rajaperf
{
namespace algorithm
{

void SORT::runB200Variant(VariantID vid, size_t RAJAPERF_UNUSED_ARG(tune_idx))
{
  const Index_type run_reps = getRunReps();
  auto res{getCudaResource()};
  rpb200_stream_t stream = res.get_stream();
  rpb200_ctx* ctx = getB200Context();

  if ( vid != Base_B200 ) {
    getCout() << "\n  SORT : Unknown B200 variant id = " << vid << std::endl;
    return;
  }

  const Index_type iend = getActualProblemSize();
  const size_t scratch_bytes = rpb200_sort_scratch_bytes(iend, /*pairs=*/0);
  void* scratch = nullptr;
  cudaErrchk( cudaMalloc(&scratch, scratch_bytes) );

  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) {                 // rep irep sorts its own segment (SORT.hpp:21-25)
    checkB200( rpb200_sort_keys_f64(ctx, m_x + iend*irep, iend, scratch, scratch_bytes, stream), "rpb200_sort_keys_f64" );
  }
  stopTimer();

  cudaErrchk( cudaFree(scratch) );
}

} // end namespace algorithm
} // end namespace rajaperf

#endif  // RAJA_ENABLE_CUDA
