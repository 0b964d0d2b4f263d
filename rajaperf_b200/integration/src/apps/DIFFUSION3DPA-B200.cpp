//
// DIFFUSION3DPA-B200.cpp -- the Base_B200 variant of Apps_DIFFUSION3DPA: the analogue of DIFFUSION3DPA-Cuda.cpp, added to the reference tree by
// rajaperf_b200/integration/apply_base_b200.py.  The kernel is one call into librpb200.so (include/rpb200.h) per rep,
// enqueued on the suite's CUDA stream between startTimer() and stopTimer(); data are the arrays setUp() allocated.
//
#include "DIFFUSION3DPA.hpp"

#include "RAJA/RAJA.hpp"

#if defined(RAJA_ENABLE_CUDA)

#include "common/B200Utils.hpp"

#include <iostream>

namespace rajaperf
{
namespace apps
{

void DIFFUSION3DPA::runB200Variant(VariantID vid, size_t RAJAPERF_UNUSED_ARG(tune_idx))
{
  const Index_type run_reps = getRunReps();
  auto res{getCudaResource()};
  rpb200_stream_t stream = res.get_stream();
  rpb200_ctx* ctx = getB200Context();

  if ( vid != Base_B200 ) {
    getCout() << "\n  DIFFUSION3DPA : Unknown B200 variant id = " << vid << std::endl;
    return;
  }

  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) {                 // Basis = m_B, dBasis = m_G (DIFFUSION3DPA.hpp:216-223)
    checkB200( rpb200_diffusion3dpa(ctx, m_B, m_G, m_D, m_X, m_Y, m_NE, /*symmetric=*/1, stream), "rpb200_diffusion3dpa" );
  }
  stopTimer();
}

} // end namespace apps
} // end namespace rajaperf

#endif  // RAJA_ENABLE_CUDA
