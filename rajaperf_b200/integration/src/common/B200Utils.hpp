//
// B200Utils.hpp -- glue between the RAJA Performance Suite and librpb200.so (include/rpb200.h), added to the reference tree by
// rajaperf_b200/integration/apply_base_b200.py.  One process-wide context replaces the per-kernel camp::resources::Cuda
// scratch of the reference's GPU variants; errors abort like cudaErrchk (common/CudaDataUtils.hpp:67): nothing falls back.
//
#ifndef RAJAPerf_B200Utils_HPP
#define RAJAPerf_B200Utils_HPP

#include "rajaperf_config.hpp"

#if defined(RAJA_ENABLE_CUDA)

#include "rpb200.h"

#include "common/RAJAPerfSuite.hpp"
#include "common/CudaDataUtils.hpp"

#include <cstdlib>
#include <iostream>

namespace rajaperf
{

inline void checkB200(int err, const char* what)
{
  if (err != 0) {
    std::cerr << "\nBase_B200: " << what << " failed: " << rpb200_error_string(err)
              << " (there is no fallback path)" << std::endl;
    std::abort();
  }
}

inline rpb200_ctx* getB200Context()
{
  static rpb200_ctx* ctx = nullptr;
  if (!ctx) {
    int dev = 0;
    cudaErrchk( cudaGetDevice(&dev) );
    checkB200( rpb200_create(dev, &ctx), "rpb200_create (an sm_100 device is required)" );
  }
  return ctx;
}

}  // closing brace for rajaperf namespace

#endif  // RAJA_ENABLE_CUDA

#endif  // closing endif for header file include guard
