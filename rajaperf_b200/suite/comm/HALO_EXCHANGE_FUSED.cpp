// HALO_EXCHANGE_FUSED.cpp -- Comm_HALO_EXCHANGE_FUSED (reference: comm/HALO_EXCHANGE_FUSED.cpp:17-140).
#include <cstdio>
#include <vector>

#include <cuda_runtime_api.h>

#include "Comm.hpp"

namespace rajaperf {
namespace comm {

HALO_EXCHANGE_FUSED::HALO_EXCHANGE_FUSED(KernelID kid, const RunParams& params) : HALO_base(kid, params)
{
  setDefaultReps(200);
  setItsPerRep(m_num_vars * m_halo_elems);                  // HALO_EXCHANGE_FUSED.cpp:32
  setKernelsPerRep(2);
  // HALO_EXCHANGE_FUSED.cpp:34-45: pack + unpack as in HALO_PACKING_FUSED plus the message itself
  // (one Real_type read by the sender and written at the receiver)
  setBytesReadPerRep(2 * m_num_vars * m_halo_elems * (sizeof(Int_type) + sizeof(Real_type)) + m_num_vars * m_halo_elems * sizeof(Real_type));
  setBytesWrittenPerRep(2 * m_num_vars * m_halo_elems * sizeof(Real_type) + m_num_vars * m_halo_elems * sizeof(Real_type));
  setFLOPsPerRep(0);
  setVariantDefined(Base_B200);
}

HALO_EXCHANGE_FUSED::~HALO_EXCHANGE_FUSED()
{
  for (rpb200_ctx* c : m_dev_ctx) rpb200_destroy(c);
}

void HALO_EXCHANGE_FUSED::setUp(VariantID, size_t)
{
  const auto& div = run_params.getMPI3DDivision();
  const int rank_dims[3] = {div[0], div[1], div[2]};
  const int P = run_params.getNumRanks();
  m_first_device = run_params.getDevice();
  int ndev = 1;
  cudaGetDeviceCount(&ndev);
  m_num_devices = std::max(1, std::min(ndev - m_first_device, P));
  if (m_dev_ctx.empty()) {
    m_dev_ctx.assign(m_num_devices, nullptr);
    for (int d = 0; d < m_num_devices; ++d) checkAbi(rpb200_create(m_first_device + d, &m_dev_ctx[d]), "rpb200_create");
    for (int a = 0; a < m_num_devices; ++a)
      for (int b = 0; b < m_num_devices; ++b)
        if (a != b) checkAbi(rpb200_enable_peer_access(m_first_device + a, m_first_device + b), "rpb200_enable_peer_access");
  }
  m_ranks.assign(P, Rank());
  std::vector<Real_type> h((size_t)m_var_size);
  std::vector<void*> windows(P, nullptr);
  for (int r = 0; r < P; ++r) {
    Rank& rk = m_ranks[r];
    rk.device = m_first_device + r % m_num_devices;
    rk.c = m_dev_ctx[r % m_num_devices];
    cudaSetDevice(rk.device);
    detail::resetDataInitCount();                       // every MPI rank runs its own setUp
    rk.plan = setUp_base(rk.c, r, rank_dims);
    rk.vars.assign(m_num_vars, nullptr);
    for (Index_type v = 0; v < m_num_vars; ++v) {       // HALO_EXCHANGE_FUSED.cpp:83-92: var[i] = i + v
      allocData(rk.vars[v], m_var_size);
      detail::incDataInitCount();
      for (Index_type i = 0; i < m_var_size; ++i) h[i] = i + v;
      copyToDevice(rk.vars[v], h.data(), sizeof(Real_type) * (size_t)m_var_size);
    }
    checkAbi(rpb200_halo_exchange_window(rk.plan, rk.vars.data(), &windows[r], nullptr, nullptr), "rpb200_halo_exchange_window");
  }
  for (int r = 0; r < P; ++r) {
    cudaSetDevice(m_ranks[r].device);
    checkAbi(rpb200_halo_exchange_connect_ptrs(m_ranks[r].plan, P, windows.data()), "rpb200_halo_exchange_connect_ptrs");
  }
  cudaSetDevice(m_first_device);
}

// HALO_EXCHANGE_FUSED.cpp:123-128 per rank; the report averages the ranks' checksums (Executor.cpp:1392-1467)
void HALO_EXCHANGE_FUSED::updateChecksum(VariantID vid, size_t tune_idx)
{
  Checksum_type sum = 0.0;
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    int err = rpb200_halo_exchange_status(rk.plan);
    if (err != 0) std::fprintf(stderr, "\nHALO_EXCHANGE_FUSED: a rank timed out waiting for a message (%d)\n", err);
    Checksum_type ck = 0.0;
    for (Real_ptr var : rk.vars) ck += calcChecksum(var, m_var_size);
    sum += ck;
  }
  cudaSetDevice(m_first_device);
  checksum[vid][tune_idx] += sum / static_cast<Checksum_type>(m_ranks.size());
}

void HALO_EXCHANGE_FUSED::tearDown(VariantID, size_t)
{
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    for (Real_ptr& p : rk.vars) deallocData(p);
    rpb200_halo_plan_destroy(rk.plan);
  }
  m_ranks.clear();
  cudaSetDevice(m_first_device);
}

}  // namespace comm
}  // namespace rajaperf
