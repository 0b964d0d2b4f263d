// Stream.hpp -- the Stream group behind KernelBase (reference: stream/{ADD,COPY,DOT,MUL,TRIAD}.hpp).
//   ADD   c[i] = a[i] + b[i]          COPY  c[i] = a[i]          MUL  b[i] = alpha*c[i]
//   TRIAD a[i] = b[i] + alpha*c[i]    DOT   dot += a[i]*b[i]
#pragma once
#include "../common/KernelBase.hpp"

namespace rajaperf {
namespace stream {

// Launch shapes of the element-wise kernels as suite tunings (the reference's block_<N> tunings, stream/TRIAD-Cuda.cpp:
// 60-100): {threads per CTA, persistent CTAs per SM (0 = one tile per CTA), 256-bit vectors per thread}.
inline void defineElementwiseTunings(KernelBase& k, VariantID vid)
{
  k.addB200Tuning(vid, KernelBase::getDefaultTuningName());      // 512 threads, one tile per CTA, 2 vectors per thread
  k.addB200Tuning(vid, "block_256", 256, 0, 2);
  k.addB200Tuning(vid, "persistent_8", 512, 8, 4);               // grid-stride, 8 CTAs per SM, 4 vectors per thread
}

#define RPB_STREAM_KERNEL(NAME, MEMBERS)                                   \
  class NAME : public KernelBase {                                         \
  public:                                                                  \
    explicit NAME(const RunParams& params);                                \
    void setUp(VariantID vid, size_t tune_idx) override;                   \
    void updateChecksum(VariantID vid, size_t tune_idx) override;          \
    void tearDown(VariantID vid, size_t tune_idx) override;                \
    void runB200Variant(VariantID vid, size_t tune_idx) override;          \
    void enqueueRep(rpb200_stream_t s) override;                           \
    void setB200TuningDefinitions(VariantID vid) override;                 \
  private:                                                                 \
    MEMBERS                                                                \
  };

RPB_STREAM_KERNEL(ADD, Real_ptr m_a = nullptr; Real_ptr m_b = nullptr; Real_ptr m_c = nullptr;)
RPB_STREAM_KERNEL(COPY, Real_ptr m_a = nullptr; Real_ptr m_c = nullptr;)
RPB_STREAM_KERNEL(MUL, Real_ptr m_b = nullptr; Real_ptr m_c = nullptr; Real_type m_alpha = 0.0;)
RPB_STREAM_KERNEL(TRIAD, Real_ptr m_a = nullptr; Real_ptr m_b = nullptr; Real_ptr m_c = nullptr; Real_type m_alpha = 0.0;)
RPB_STREAM_KERNEL(DOT, void finishReps() override; Real_ptr m_a = nullptr; Real_ptr m_b = nullptr; Real_type m_dot = 0.0;
                  Real_type m_dot_init = 0.0; Real_ptr m_d_dot = nullptr;)
#undef RPB_STREAM_KERNEL

}  // namespace stream
}  // namespace rajaperf
