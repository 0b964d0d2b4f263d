// Basic.hpp -- Basic_INDEXLIST / Basic_INDEXLIST_3LOOP behind KernelBase (widened rows, SURVEY 8f;
// reference: basic/INDEXLIST.{hpp,cpp}, basic/INDEXLIST_3LOOP.{hpp,cpp}).
#pragma once
#include "../common/KernelBase.hpp"

namespace rajaperf {
namespace basic {

class INDEXLIST : public KernelBase {      // list[count++] = i where x[i] < 0.0; m_len = count
public:
  explicit INDEXLIST(const RunParams& params) : INDEXLIST(rajaperf::Basic_INDEXLIST, params) {}
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
  void finishReps() override;
protected:
  INDEXLIST(KernelID kid, const RunParams& params);
  Real_ptr m_x = nullptr;
  Int_ptr m_list = nullptr;
  Index_type* m_d_len = nullptr;           // device Index_type written by the kernel
  Index_type m_len = -1;
};

// Same result through three loops + an (N+1)-entry `counts` temporary in the reference
// (INDEXLIST_3LOOP-Seq.cpp:43-65); Base_B200 serves it with the same fused single pass.
class INDEXLIST_3LOOP : public INDEXLIST {
public:
  explicit INDEXLIST_3LOOP(const RunParams& params);
};

}  // namespace basic
}  // namespace rajaperf
