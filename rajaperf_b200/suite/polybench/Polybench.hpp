// Polybench.hpp -- Polybench_GEMM behind KernelBase (widened row; reference: polybench/POLYBENCH_GEMM.{hpp,cpp}).
#pragma once
#include "../common/KernelBase.hpp"

namespace rajaperf {
namespace polybench {

class POLYBENCH_GEMM : public KernelBase {     // C[i][j] = sum_k alpha * A[i][k] * B[k][j]
public:
  explicit POLYBENCH_GEMM(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
  void setB200TuningDefinitions(VariantID vid) override;
private:
  Index_type m_ni, m_nj, m_nk;
  Real_type m_alpha, m_beta;
  Real_ptr m_A = nullptr, m_B = nullptr, m_C = nullptr;
};

}  // namespace polybench
}  // namespace rajaperf
