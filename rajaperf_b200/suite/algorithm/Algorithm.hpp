// Algorithm.hpp -- REDUCE_SUM, SCAN, SORT, SORTPAIRS behind KernelBase
// (reference: algorithm/{REDUCE_SUM,SCAN,SORT,SORTPAIRS}.hpp).
#pragma once
#include "../common/KernelBase.hpp"

namespace rajaperf {
namespace algorithm {

class REDUCE_SUM : public KernelBase {     // sum = init + sum_i x[i]; m_sum = last rep's sum
public:
  explicit REDUCE_SUM(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
  void finishReps() override;
private:
  Real_ptr m_x = nullptr, m_d_sum = nullptr;
  Real_type m_sum_init = 0.0, m_sum = 0.0;
};

class SCAN : public KernelBase {           // y[i] = sum_{j<i} x[j]
public:
  explicit SCAN(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
private:
  Real_ptr m_x = nullptr, m_y = nullptr;
};

class SORT : public KernelBase {           // rep irep sorts segment irep of x (SORT.hpp:21-25)
public:
  explicit SORT(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
protected:
  Real_ptr m_x = nullptr;
  void* m_scratch = nullptr;
  size_t m_scratch_bytes = 0;
  Index_type m_rep = 0;
};

class SORTPAIRS : public KernelBase {      // same, (key x, value i) pairs
public:
  explicit SORTPAIRS(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
private:
  Real_ptr m_x = nullptr, m_i = nullptr;
  void* m_scratch = nullptr;
  size_t m_scratch_bytes = 0;
  Index_type m_rep = 0;
};

// Calibration streams (widened rows; reference algorithm/MEMCPY.{hpp,cpp}, MEMSET.{hpp,cpp}).
class MEMCPY : public KernelBase {         // y[i] = x[i]
public:
  explicit MEMCPY(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
private:
  Real_ptr m_x = nullptr, m_y = nullptr;
};

class MEMSET : public KernelBase {         // x[i] = val
public:
  explicit MEMSET(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
private:
  Real_ptr m_x = nullptr;
  Real_type m_val = 0.0;
};

}  // namespace algorithm
}  // namespace rajaperf
