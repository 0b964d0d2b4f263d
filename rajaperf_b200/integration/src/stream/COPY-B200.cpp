//
// COPY-B200.cpp -- the Base_B200 variant of Stream_COPY: the analogue of COPY-Cuda.cpp, added to the reference tree by
// rajaperf_b200/integration/apply_base_b200.py.  The kernel is one call into librpb200.so (include/rpb200.h) per rep,
// enqueued on the suite's CUDA stream between startTimer() and stopTimer(); data are the arrays setUp() allocated.
//
#include "COPY.hpp"

#include "RAJA/RAJA.hpp"

#if defined(RAJA_ENABLE_CUDA)

#include "common/B200Utils.hpp"

#include <iostream>

namespace rajaperf
{
namespace stream
{

void COPY::runB200Variant(VariantID vid, size_t RAJAPERF_UNUSED_ARG(tune_idx))
{
  const Index_type run_reps = getRunReps();
  auto res{getCudaResource()};
  rpb200_stream_t stream = res.get_stream();
  rpb200_ctx* ctx = getB200Context();

  if ( vid != Base_B200 ) {
    getCout() << "\n  COPY : Unknown B200 variant id = " << vid << std::endl;
    return;
  }

  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) {
    checkB200( rpb200_stream_copy(ctx, m_c, m_a, getActualProblemSize(), stream), "rpb200_stream_copy" );
  }
  stopTimer();
}

} // end namespace stream
} // end namespace rajaperf

#endif  // RAJA_ENABLE_CUDA
