// HALO_SENDRECV.cpp -- transport only (comm/HALO_SENDRECV.cpp:21-123, HALO_SENDRECV-Seq.cpp:34-52).
#include <algorithm>
#include <cstdio>
#include <cuda_runtime_api.h>

#include "Comm.hpp"

namespace rajaperf {
namespace comm {

HALO_SENDRECV::HALO_SENDRECV(const RunParams& params) : HALO_EXCHANGE_FUSED(rajaperf::Comm_HALO_SENDRECV, params)
{
  setItsPerRep(m_num_vars * m_halo_elems);
  setKernelsPerRep(2);                                    // put + wait (the reference counts 0: MPI calls only)
  setBytesReadPerRep(1 * sizeof(Real_type) * getItsPerRep());      // HALO_SENDRECV.cpp:33-34
  setBytesWrittenPerRep(1 * sizeof(Real_type) * getItsPerRep());
  setFLOPsPerRep(0);
}

void HALO_SENDRECV::setUp(VariantID, size_t)
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
  m_send.assign(P, std::vector<Real_ptr>(s_num_neighbors, nullptr));
  std::vector<void*> windows(P, nullptr);
  for (int r = 0; r < P; ++r) {
    Rank& rk = m_ranks[r];
    rk.device = m_first_device + r % m_num_devices;
    rk.c = m_dev_ctx[r % m_num_devices];
    cudaSetDevice(rk.device);
    detail::resetDataInitCount();
    rk.plan = setUp_base(rk.c, r, rank_dims);               // 52 list allocations: init counter -> 52
    for (int l = 0; l < s_num_neighbors; ++l) {             // HALO_SENDRECV.cpp:66-76: send buffers, initData
      int64_t pl = 0;
      checkAbi(rpb200_halo_plan_neighbor(rk.plan, l, nullptr, nullptr, nullptr, &pl, nullptr, nullptr, nullptr), "rpb200_halo_plan_neighbor");
      allocAndInitData(m_send[r][l], m_num_vars * pl);
    }
    // :78-88: the receive buffers are allocAndInitData'ed too; here they ARE the window, so only the counter moves
    for (int l = 0; l < s_num_neighbors; ++l) detail::incDataInitCount();
    checkAbi(rpb200_halo_exchange_window(rk.plan, nullptr, &windows[r], nullptr, nullptr), "rpb200_halo_exchange_window");
  }
  for (int r = 0; r < P; ++r) {
    cudaSetDevice(m_ranks[r].device);
    checkAbi(rpb200_halo_exchange_connect_ptrs(m_ranks[r].plan, P, windows.data()), "rpb200_halo_exchange_connect_ptrs");
    checkAbi(rpb200_halo_sendrecv_bind(m_ranks[r].plan, m_send[r].data()), "rpb200_halo_sendrecv_bind");
  }
  cudaSetDevice(m_first_device);
}

void HALO_SENDRECV::enqueueRep(rpb200_stream_t s)
{
  // ranks may share a GPU and its stream: every put is queued before the first wait spins
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    checkAbi(rpb200_halo_sendrecv_put(rk.plan, s), "rpb200_halo_sendrecv_put");
  }
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    checkAbi(rpb200_halo_sendrecv_wait(rk.plan, s), "rpb200_halo_sendrecv_wait");
  }
  cudaSetDevice(m_first_device);
}

void HALO_SENDRECV::updateChecksum(VariantID vid, size_t tune_idx)     // HALO_SENDRECV.cpp:91-105, rank-averaged
{
  Checksum_type sum = 0.0;
  for (Rank& rk : m_ranks) {
    cudaSetDevice(rk.device);
    int err = rpb200_halo_exchange_status(rk.plan);
    if (err != 0) std::fprintf(stderr, "\nHALO_SENDRECV: a rank timed out waiting for a message (%d)\n", err);
    for (int l = 0; l < s_num_neighbors; ++l) {
      const double* p = nullptr; int64_t len = 0;
      checkAbi(rpb200_halo_recv_buffer(rk.plan, l, &p, &len), "rpb200_halo_recv_buffer");
      sum += calcChecksum(p, len);
    }
  }
  cudaSetDevice(m_first_device);
  checksum[vid][tune_idx] += sum / static_cast<Checksum_type>(m_ranks.size());
}

void HALO_SENDRECV::tearDown(VariantID, size_t)
{
  for (size_t r = 0; r < m_ranks.size(); ++r) {
    cudaSetDevice(m_ranks[r].device);
    for (Real_ptr& p : m_send[r]) deallocData(p);
    rpb200_halo_plan_destroy(m_ranks[r].plan);
  }
  m_ranks.clear(); m_send.clear();
  cudaSetDevice(m_first_device);
}

}  // namespace comm
}  // namespace rajaperf
