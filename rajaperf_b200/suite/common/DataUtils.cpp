#include "DataUtils.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../../include/rpb200.h"

namespace rajaperf {

void checkAbi(int err, const char* what)
{
  if (err != 0) {
    std::fprintf(stderr, "\nrpb200 error %d (%s) in %s -- aborting (there is no CPU fallback)\n", err,
                 rpb200_error_string(err), what);
    std::abort();
  }
}

namespace detail {
static int data_init_count = 0;
void resetDataInitCount() { data_init_count = 0; }
void incDataInitCount() { data_init_count++; }
int getDataInitCount() { return data_init_count; }
}  // namespace detail

static Real_type parityFactor() { return (detail::getDataInitCount() % 2) ? 0.1 : 0.2; }

static Real_ptr deviceAlloc(Index_type len)
{
  void* p = nullptr;
  checkAbi(rpb200_malloc(&p, sizeof(Real_type) * (Size_type)(len > 0 ? len : 1)), "rpb200_malloc");
  return static_cast<Real_ptr>(p);
}

void copyToDevice(void* d, const void* h, Size_type bytes)
{
  if (bytes == 0) return;
  checkAbi(rpb200_memcpy_h2d(d, h, bytes, nullptr), "rpb200_memcpy_h2d");
  checkAbi(rpb200_stream_synchronize(nullptr), "rpb200_stream_synchronize");
}

void copyToHost(void* h, const void* d, Size_type bytes)
{
  if (bytes == 0) return;
  checkAbi(rpb200_memcpy_d2h(h, d, bytes, nullptr), "rpb200_memcpy_d2h");
  checkAbi(rpb200_stream_synchronize(nullptr), "rpb200_stream_synchronize");
}

void allocData(Real_ptr& d_ptr, Index_type len) { d_ptr = deviceAlloc(len); }

template <typename F>
static void allocInit(Real_ptr& d_ptr, Index_type len, F fill)
{
  d_ptr = deviceAlloc(len);
  std::vector<Real_type> h((Size_type)len);
  fill(h.data());
  copyToDevice(d_ptr, h.data(), sizeof(Real_type) * (Size_type)len);
  detail::incDataInitCount();
}

void allocAndInitData(Real_ptr& d_ptr, Index_type len)
{
  const Real_type factor = parityFactor();
  allocInit(d_ptr, len, [&](Real_type* h) {
    for (Index_type i = 0; i < len; ++i) h[i] = factor * (i + 1.1) / (i + 1.12345);
  });
}

void allocAndInitDataConst(Real_ptr& d_ptr, Index_type len, Real_type val)
{
  allocInit(d_ptr, len, [&](Real_type* h) { for (Index_type i = 0; i < len; ++i) h[i] = val; });
}

void allocAndInitDataRandValue(Real_ptr& d_ptr, Index_type len)
{
  allocInit(d_ptr, len, [&](Real_type* h) {
    std::srand(4793);                                                    // re-seeded on every call
    for (Index_type i = 0; i < len; ++i) h[i] = Real_type(std::rand()) / RAND_MAX;
  });
}

void allocAndInitDataRandSign(Real_ptr& d_ptr, Index_type len)
{
  const Real_type factor = parityFactor();
  allocInit(d_ptr, len, [&](Real_type* h) {
    std::srand(4793);
    for (Index_type i = 0; i < len; ++i) {
      Real_type signfact = Real_type(std::rand()) / RAND_MAX;
      signfact = (signfact < 0.5) ? -1.0 : 1.0;
      h[i] = signfact * factor * (i + 1.1) / (i + 1.12345);
    }
  });
}

void allocAndInitData(Int_ptr& d_ptr, Index_type len)        // DataUtils.cpp:477-497
{
  void* p = nullptr;
  checkAbi(rpb200_malloc(&p, sizeof(Int_type) * (Size_type)(len > 0 ? len : 1)), "rpb200_malloc");
  d_ptr = static_cast<Int_ptr>(p);
  std::vector<Int_type> h((Size_type)len);
  std::srand(4793);
  Real_type signfact = 0.0;
  for (Index_type i = 0; i < len; ++i) {
    signfact = Real_type(std::rand()) / RAND_MAX;
    h[i] = (signfact < 0.5 ? -1 : 1);
  }
  if (len > 0) {
    signfact = Real_type(std::rand()) / RAND_MAX;
    h[(Size_type)(len * signfact)] = -58;
    signfact = Real_type(std::rand()) / RAND_MAX;
    h[(Size_type)(len * signfact)] = 19;
  }
  copyToDevice(d_ptr, h.data(), sizeof(Int_type) * (Size_type)len);
  detail::incDataInitCount();
}

void initData(Real_type& d)
{
  const Real_type factor = parityFactor();
  d = factor * 1.1 / 1.12345;
  detail::incDataInitCount();
}

void deallocData(Real_ptr& d_ptr) { if (d_ptr) checkAbi(rpb200_free(d_ptr), "rpb200_free"); d_ptr = nullptr; }
void deallocData(Int_ptr& d_ptr) { if (d_ptr) checkAbi(rpb200_free(d_ptr), "rpb200_free"); d_ptr = nullptr; }

Checksum_type calcChecksumHost(const Real_type* ptr, Index_type len, Real_type scale_factor)
{
  Checksum_type tchk = 0.0, ckahan = 0.0;
  for (Index_type j = 0; j < len; ++j) {
    // the weight is formed in double, the product in long double (DataUtils.cpp:606, :631)
    const Checksum_type x = (std::abs(std::sin(j + 1.0)) + 0.5) * static_cast<Checksum_type>(ptr[j]);
    const Checksum_type y = x - ckahan;
    volatile Checksum_type t = tchk + y;
    volatile Checksum_type z = t - tchk;
    ckahan = z - y;
    tchk = t;
  }
  tchk *= scale_factor;
  return tchk;
}

Checksum_type calcChecksum(const Int_type* d_ptr, Index_type len, Real_type scale_factor)
{
  std::vector<Int_type> h((Size_type)len);
  copyToHost(h.data(), d_ptr, sizeof(Int_type) * (Size_type)len);
  Checksum_type tchk = 0.0, ckahan = 0.0;
  for (Index_type j = 0; j < len; ++j) {
    const Checksum_type x = (std::abs(std::sin(j + 1.0)) + 0.5) * static_cast<Checksum_type>(h[j]);
    const Checksum_type y = x - ckahan;
    volatile Checksum_type t = tchk + y;
    volatile Checksum_type z = t - tchk;
    ckahan = z - y;
    tchk = t;
  }
  tchk *= scale_factor;
  return tchk;
}

Checksum_type calcChecksum(const Real_type* d_ptr, Index_type len, Real_type scale_factor)
{
  std::vector<Real_type> h((Size_type)len);
  copyToHost(h.data(), d_ptr, sizeof(Real_type) * (Size_type)len);
  return calcChecksumHost(h.data(), len, scale_factor);
}

}  // namespace rajaperf
