// Comm.hpp -- HALO_base, HALO_PACKING_FUSED, HALO_EXCHANGE_FUSED behind KernelBase
// (reference: comm/HALO_base.{hpp,cpp}, HALO_PACKING_FUSED.{hpp,cpp}, HALO_EXCHANGE_FUSED.{hpp,cpp}).
#pragma once
#include <vector>

#include "../common/KernelBase.hpp"

namespace rajaperf {
namespace comm {

// Grid dimensions and the per-rank halo plan (26 neighbours, 52 index lists on the device).
class HALO_base : public KernelBase {
public:
  HALO_base(KernelID kid, const RunParams& params);
  static constexpr int s_num_neighbors = RPB200_HALO_NEIGHBORS;
protected:
  Index_type m_grid_dims[3], m_halo_width, m_num_vars, m_var_size = 0;
  Index_type m_halo_elems = 0;      // sum over neighbours of the pack-list length
  // allocates the plan of `rank` on `c` and bumps the init counter once per list, as
  // HALO_base::create_lists does through allocAndInitData (HALO_base.cpp:228-232, 262-266)
  rpb200_halo_plan* setUp_base(rpb200_ctx* c, int rank, const int* rank_dims);
};

class HALO_PACKING_FUSED : public HALO_base {
public:
  explicit HALO_PACKING_FUSED(const RunParams& params) : HALO_PACKING_FUSED(rajaperf::Comm_HALO_PACKING_FUSED, params) {}
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
  void setB200TuningDefinitions(VariantID vid) override;
protected:
  HALO_PACKING_FUSED(KernelID kid, const RunParams& params);
  rpb200_halo_plan* m_plan = nullptr;
  std::vector<Real_ptr> m_vars, m_pack_buffers, m_unpack_buffers;
  std::vector<Index_type> m_pack_lens, m_unpack_lens;
};

// The unfused kernel (widened row; reference comm/HALO_PACKING.{hpp,cpp}, -Cuda.cpp:26-47): same
// setUp, same result, but one launch per (neighbour, variable) -- 78 packs + 78 unpacks per rep -- so the
// report shows what the fusion buys.  Each launch is a one-tuple work list; the rep batch replays from a
// CUDA graph like every other kernel.
class HALO_PACKING : public HALO_PACKING_FUSED {
public:
  explicit HALO_PACKING(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
private:
  std::vector<rpb200_halo_worklist*> m_pack_wl, m_unpack_wl;     // [neighbour * num_vars + var]
};

// All px*py*pz ranks live in this process; rank r runs on CUDA device first + (r mod ndev).  Every rank
// owns its vars and its receive window; windows are connected through peer pointers (NVLink P2P when
// the ranks sit on different GPUs).
class HALO_EXCHANGE_FUSED : public HALO_base {
public:
  explicit HALO_EXCHANGE_FUSED(const RunParams& params) : HALO_EXCHANGE_FUSED(rajaperf::Comm_HALO_EXCHANGE_FUSED, params) {}
  ~HALO_EXCHANGE_FUSED() override;
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
protected:
  HALO_EXCHANGE_FUSED(KernelID kid, const RunParams& params);
  struct Rank {
    int device = 0;
    rpb200_ctx* c = nullptr;
    rpb200_halo_plan* plan = nullptr;
    std::vector<Real_ptr> vars;
  };
  std::vector<Rank> m_ranks;
  std::vector<rpb200_ctx*> m_dev_ctx;     // one context per device used
  int m_first_device = 0, m_num_devices = 1;
};

// HALO_SENDRECV (widened row; reference comm/HALO_SENDRECV.{hpp,cpp}): transport only.  Per rank 26 send buffers
// (initData) are put into the neighbours' receive windows; the checksum is over what arrived.
class HALO_SENDRECV : public HALO_EXCHANGE_FUSED {
public:
  explicit HALO_SENDRECV(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
private:
  std::vector<std::vector<Real_ptr>> m_send;      // [rank][neighbour]
};

// The unfused exchange (widened row; reference comm/HALO_EXCHANGE.{hpp,cpp}, -Cuda.cpp:26-123): same data,
// same result, one pack launch and one unpack launch per (neighbour, variable) on every rank.
class HALO_EXCHANGE : public HALO_EXCHANGE_FUSED {
public:
  explicit HALO_EXCHANGE(const RunParams& params);
  void enqueueRep(rpb200_stream_t s) override;
};

}  // namespace comm
}  // namespace rajaperf
