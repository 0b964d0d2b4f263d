// KernelBase.hpp -- the kernel contract (reference: common/KernelBase.{hpp,cpp}).
//
// Same life cycle and bookkeeping as the reference: execute() = reset init counter -> setUp ->
// runKernel -> updateChecksum -> tearDown (KernelBase.cpp:359-377); runKernel switches on the
// VariantID (KernelBase.cpp:391-489) -- here to the one new arm, runB200Variant(); startTimer /
// stopTimer bracket the rep loop with a device synchronisation and a host steady_clock exactly like
// KernelBase.hpp:423-440, and ADDITIONALLY record a cudaEvent pair on the kernel's stream, so every
// (variant, tuning) carries both the reference's wall time and the device time of the rep batch.
#pragma once
#include <chrono>
#include <limits>
#include <string>
#include <vector>

#include "../../../include/rpb200.h"
#include "DataUtils.hpp"
#include "RAJAPerfSuite.hpp"
#include "RPTypes.hpp"
#include "RunParams.hpp"

namespace rajaperf {

class KernelBase {
public:
  KernelBase(KernelID kid, const RunParams& params);
  virtual ~KernelBase();

  KernelID getKernelID() const { return kernel_id; }
  const std::string& getName() const { return name; }
  // the library's tuning-table entry of this kernel (rpb200_set_tuning's `kernel`): its own full name
  virtual const std::string& tuningKernelName() const { return name; }

  // properties set by kernel constructors (KernelBase.hpp:100-116)
  void setDefaultProblemSize(Index_type size) { default_prob_size = size; }
  void setActualProblemSize(Index_type size) { actual_prob_size = size; }
  void setDefaultReps(Index_type reps) { default_reps = reps; }
  void setItsPerRep(Index_type its) { its_per_rep = its; }
  void setKernelsPerRep(Index_type n) { kernels_per_rep = n; }
  void setBytesReadPerRep(Index_type b) { bytes_read_per_rep = b; }
  void setBytesWrittenPerRep(Index_type b) { bytes_written_per_rep = b; }
  void setFLOPsPerRep(Index_type f) { FLOPs_per_rep = f; }
  void setVariantDefined(VariantID vid);
  void addVariantTuningName(VariantID vid, std::string n) { variant_tuning_names[vid].emplace_back(std::move(n)); }
  static const std::string& getDefaultTuningName() { static const std::string n("default"); return n; }   // KernelBase.hpp:76
  // A Base_B200 tuning = a named launch shape: the rpb200_set_tuning arguments execute() applies around the run
  // (the reference's block_128/block_256/... tunings, GPUUtils.hpp:345-373).  {0,-1,0} keeps the library default.
  void addB200Tuning(VariantID vid, std::string n, int block_size = 0, int ctas_per_sm = -1, int unroll = 0)
  { addVariantTuningName(vid, std::move(n)); b200_tunings.push_back({block_size, ctas_per_sm, unroll}); }
  virtual void setB200TuningDefinitions(VariantID vid) { addB200Tuning(vid, getDefaultTuningName()); }

  Index_type getDefaultProblemSize() const { return default_prob_size; }
  Index_type getActualProblemSize() const { return actual_prob_size; }
  Index_type getDefaultReps() const { return default_reps; }
  Index_type getItsPerRep() const { return its_per_rep; }
  Index_type getKernelsPerRep() const { return kernels_per_rep; }
  Index_type getBytesPerRep() const { return bytes_read_per_rep + bytes_written_per_rep; }
  Index_type getBytesReadPerRep() const { return bytes_read_per_rep; }
  Index_type getBytesWrittenPerRep() const { return bytes_written_per_rep; }
  Index_type getFLOPsPerRep() const { return FLOPs_per_rep; }
  Index_type getTargetProblemSize() const;    // KernelBase.cpp:115-125
  Index_type getRunReps() const;              // KernelBase.cpp:127-136

  bool hasVariantDefined(VariantID vid) const { return !variant_tuning_names[vid].empty(); }
  size_t getNumVariantTunings(VariantID vid) const { return variant_tuning_names[vid].size(); }
  const std::string& getVariantTuningName(VariantID vid, size_t t) const { return variant_tuning_names[vid][t]; }
  const std::vector<std::string>& getVariantTuningNames(VariantID vid) const { return variant_tuning_names[vid]; }
  // KernelBase.hpp:196-204: npos when this kernel does not define that tuning
  size_t getVariantTuningIndex(VariantID vid, const std::string& tuning_name) const;
  bool hasVariantTuningDefined(VariantID vid, const std::string& tuning_name) const
  { return getVariantTuningIndex(vid, tuning_name) != std::string::npos; }
  bool wasVariantTuningRun(VariantID vid, size_t t) const { return num_exec[vid][t] > 0; }

  // results (KernelBase.hpp:206-226)
  double getMinTime(VariantID vid, size_t t) const { return min_time[vid][t]; }
  double getMaxTime(VariantID vid, size_t t) const { return max_time[vid][t]; }
  double getTotTime(VariantID vid, size_t t) const { return tot_time[vid][t]; }
  double getMinDeviceTime(VariantID vid, size_t t) const { return min_dev_time[vid][t]; }
  Checksum_type getChecksum(VariantID vid, size_t t) const { return checksum[vid][t]; }

  void execute(VariantID vid, size_t tune_idx);    // KernelBase.cpp:359-377
  void runKernel(VariantID vid, size_t tune_idx);  // KernelBase.cpp:391-489

  void startTimer();     // KernelBase.hpp:423-431
  void stopTimer();      // KernelBase.hpp:433-440
  void synchronize();    // KernelBase.hpp:271-294: a GPU variant must drain the device around the timer

  // per-kernel pieces (pure virtuals of KernelBase.hpp:453-473)
  virtual void setUp(VariantID vid, size_t tune_idx) = 0;
  virtual void updateChecksum(VariantID vid, size_t tune_idx) = 0;
  virtual void tearDown(VariantID vid, size_t tune_idx) = 0;
  virtual void runB200Variant(VariantID vid, size_t tune_idx) = 0;
  // the rep body alone: enqueue ONE rep on the stream.  runB200Variant's default rep loop calls it
  // run_reps times, or captures that loop into one CUDA graph when --graph is given.
  virtual void enqueueRep(rpb200_stream_t) {}
  // runs once after the last rep, still inside the timed region (e.g. DOT's copy-back of the result)
  virtual void finishReps() {}

protected:
  const RunParams& run_params;
  std::vector<Checksum_type> checksum[NumVariants];
  Checksum_type checksum_scale_factor = 1.0;
  rpb200_ctx* ctx();                 // the process-wide Base_B200 context of device run_params.getDevice()
  rpb200_stream_t stream() const { return nullptr; }   // the legacy default stream (--gpu_stream_0 of the reference)
  void runRepLoop();                 // startTimer; reps x enqueueRep (optionally as one graph); stopTimer
  bool m_graph_ok = true;            // kernels that launch on several devices opt out of --graph

private:
  KernelID kernel_id;
  std::string name;
  Index_type default_prob_size = 0, actual_prob_size = 0, default_reps = 0;
  Index_type its_per_rep = 0, kernels_per_rep = 0, bytes_read_per_rep = 0, bytes_written_per_rep = 0, FLOPs_per_rep = 0;
  std::vector<std::string> variant_tuning_names[NumVariants];
  struct LaunchShape { int block_size, ctas_per_sm, unroll; };
  std::vector<LaunchShape> b200_tunings;      // parallel to variant_tuning_names[Base_B200]
  VariantID running_variant = NumVariants;
  size_t running_tuning = std::numeric_limits<size_t>::max();
  std::vector<int> num_exec[NumVariants];
  std::vector<double> min_time[NumVariants], max_time[NumVariants], tot_time[NumVariants], min_dev_time[NumVariants];
  std::chrono::steady_clock::time_point t_start;
  rpb200_timer* dev_timer = nullptr;
  void recordExecTime(double host_s, double dev_s);
};

}  // namespace rajaperf
