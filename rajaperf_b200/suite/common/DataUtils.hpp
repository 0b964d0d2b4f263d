// DataUtils.hpp -- the suite's synthetic inputs and checksum, for device-resident data.
//
// Restates common/DataUtils.{hpp,cpp} of the reference for DataSpace::CudaDevice: every array is
// initialised on the host with the reference's formulas (they define the benchmark's inputs, so they
// are reproduced bit for bit, including the global init-call counter that selects the 0.1 / 0.2
// factor, DataUtils.cpp:131-147), copied to cudaMalloc memory, and checksummed with the reference's
// long-double Kahan sum after a copy back (DataUtils.hpp:386-409, DataUtils.cpp:600-621).
#pragma once
#include "RPTypes.hpp"

namespace rajaperf {
namespace detail {
void resetDataInitCount();     // DataUtils.cpp:131-139 (called by KernelBase::execute)
void incDataInitCount();       // DataUtils.cpp:144-147
int getDataInitCount();
}  // namespace detail

// device allocation + host-side initialisation + H2D copy; each bumps the init counter once
void allocData(Real_ptr& d_ptr, Index_type len);                              // uninitialised
void allocAndInitData(Real_ptr& d_ptr, Index_type len);                       // initData:       DataUtils.cpp:504-513
void allocAndInitDataConst(Real_ptr& d_ptr, Index_type len, Real_type val);   // initDataConst:  DataUtils.cpp:518-525
void allocAndInitDataRandValue(Real_ptr& d_ptr, Index_type len);              // initDataRandValue: :560-569
void allocAndInitDataRandSign(Real_ptr& d_ptr, Index_type len);               // initDataRandSign:  :542-555
void allocAndInitData(Int_ptr& d_ptr, Index_type len);                        // initData(Int_ptr): :477-497
void initData(Real_type& d);                                                  // scalar: :589-595
void deallocData(Real_ptr& d_ptr);
void deallocData(Int_ptr& d_ptr);

void copyToDevice(void* d_dst, const void* h_src, Size_type bytes);
void copyToHost(void* h_dst, const void* d_src, Size_type bytes);

// checksum of a DEVICE array (copied back first) / of a host array
Checksum_type calcChecksum(const Real_type* d_ptr, Index_type len, Real_type scale_factor = 1.0);
Checksum_type calcChecksum(const Int_type* d_ptr, Index_type len, Real_type scale_factor = 1.0);   // :623-629
Checksum_type calcChecksumHost(const Real_type* h_ptr, Index_type len, Real_type scale_factor = 1.0);

void checkAbi(int err, const char* what);   // aborts with the rpb200 error string (the cudaErrchk analogue)

}  // namespace rajaperf
