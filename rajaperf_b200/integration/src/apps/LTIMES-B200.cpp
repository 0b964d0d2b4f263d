//
// LTIMES-B200.cpp -- the Base_B200 variant of Apps_LTIMES: the analogue of LTIMES-Cuda.cpp, added to the reference tree by
// rajaperf_b200/integration/apply_base_b200.py.  The kernel is one call into librpb200.so (include/rpb200.h) per rep,
// enqueued on the suite's CUDA stream between startTimer() and stopTimer(); data are the arrays setUp() allocated.
//
#include "LTIMES.hpp"

#include "RAJA/RAJA.hpp"

#if defined(RAJA_ENABLE_CUDA)

#include "common/B200Utils.hpp"

#include <iostream>

namespace rajaperf
{
namespace apps
{

void LTIMES::runB200Variant(VariantID vid, size_t RAJAPERF_UNUSED_ARG(tune_idx))
{
  const Index_type run_reps = getRunReps();
  auto res{getCudaResource()};
  rpb200_stream_t stream = res.get_stream();
  rpb200_ctx* ctx = getB200Context();

  if ( vid != Base_B200 ) {
    getCout() << "\n  LTIMES : Unknown B200 variant id = " << vid << std::endl;
    return;
  }

  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) {
    checkB200( rpb200_ltimes(ctx, m_phidat, m_elldat, m_psidat, m_num_d, m_num_g, m_num_m, m_num_z, stream), "rpb200_ltimes" );
  }
  stopTimer();
}

} // end namespace apps
} // end namespace rajaperf

#endif  // RAJA_ENABLE_CUDA
