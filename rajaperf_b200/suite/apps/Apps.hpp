// Apps.hpp -- MASS3DPA, DIFFUSION3DPA, CONVECTION3DPA, LTIMES behind KernelBase
// (reference: apps/{MASS3DPA,DIFFUSION3DPA,CONVECTION3DPA,LTIMES}.hpp).
#pragma once
#include "../common/KernelBase.hpp"

namespace rajaperf {
namespace apps {

// partial-assembly operators: Y_e += Op(B.., D_e) X_e for NE elements
class MASS3DPA : public KernelBase {        // D1D = 4, Q1D = 5 (MASS3DPA.hpp:172-173)
public:
  explicit MASS3DPA(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
  void setB200TuningDefinitions(VariantID vid) override;
  static constexpr Index_type D1D = 4, Q1D = 5;
private:
  Real_ptr m_B = nullptr, m_Bt = nullptr, m_D = nullptr, m_X = nullptr, m_Y = nullptr;
  Index_type m_NE = 0, m_NE_default = 8000;
};

class DIFFUSION3DPA : public KernelBase {   // D1D = 3, Q1D = 4, SYM = 6 (DIFFUSION3DPA.hpp:231-233)
public:
  explicit DIFFUSION3DPA(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
  static constexpr Index_type D1D = 3, Q1D = 4, SYM = 6;
private:
  Real_ptr m_B = nullptr, m_G = nullptr, m_D = nullptr, m_X = nullptr, m_Y = nullptr;
  Index_type m_NE = 0, m_NE_default = 15625;
};

class CONVECTION3DPA : public KernelBase {  // D1D = 3, Q1D = 4, VDIM = 3 (CONVECTION3DPA.hpp)
public:
  explicit CONVECTION3DPA(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
  void setB200TuningDefinitions(VariantID vid) override;
  static constexpr Index_type D1D = 3, Q1D = 4, VDIM = 3;
private:
  Real_ptr m_B = nullptr, m_Bt = nullptr, m_G = nullptr, m_D = nullptr, m_X = nullptr, m_Y = nullptr;
  Index_type m_NE = 0, m_NE_default = 15625;
};

class LTIMES : public KernelBase {          // phi[z][g][m] += sum_d ell[m][d] * psi[z][g][d]
public:
  explicit LTIMES(const RunParams& params);
  void setUp(VariantID vid, size_t tune_idx) override;
  void updateChecksum(VariantID vid, size_t tune_idx) override;
  void tearDown(VariantID vid, size_t tune_idx) override;
  void runB200Variant(VariantID vid, size_t tune_idx) override;
  void enqueueRep(rpb200_stream_t s) override;
private:
  Real_ptr m_phidat = nullptr, m_elldat = nullptr, m_psidat = nullptr;
  Index_type m_num_d, m_num_z, m_num_g, m_num_m, m_num_z_default;
  Index_type m_philen = 0, m_elllen = 0, m_psilen = 0;
};

}  // namespace apps
}  // namespace rajaperf
