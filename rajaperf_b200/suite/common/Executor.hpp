// Executor.hpp -- suite orchestration and reports (reference: common/Executor.{hpp,cpp}), hot-path subset:
// setupSuite -> reportRunSummary -> runSuite (warm-up, npasses x kernels x variants x tunings) ->
// outputRunData (timing-{Average,Minimum,Maximum}.csv, checksum.txt, kernels.csv) plus a new
// bandwidth.csv with GB/s, GFLOP/s and the fraction of the B200 roofline per kernel.
#pragma once
#include <iosfwd>
#include <string>
#include <vector>

#include "RAJAPerfSuite.hpp"
#include "RunParams.hpp"

namespace rajaperf {

class KernelBase;

class Executor {
public:
  Executor(int argc, char** argv);
  ~Executor();
  void setupSuite();          // Executor.cpp:246
  void reportRunSummary(std::ostream& str) const;   // Executor.cpp:360
  void runSuite();            // Executor.cpp:645
  void outputRunData();       // Executor.cpp:825
  const std::vector<KernelBase*>& getKernels() const { return kernels; }
  const RunParams& getRunParams() const { return run_params; }

private:
  enum CSVRepMode { Timing = 0, Speedup };
  void runKernel(KernelBase* kern, bool print_kernel_name);
  void runWarmupKernels();
  void writeKernelInfoSummary(std::ostream& str) const;
  void writeTimingCSV(const std::string& filename, int combiner);   // 0 avg, 1 min, 2 max
  void writeChecksumReport(const std::string& filename);
  void writeKernelsCSV(const std::string& filename);
  void writeBandwidthCSV(const std::string& filename);

  const std::vector<std::string>& tuningNames(VariantID vid) const { return tuning_names[vid]; }

  RunParams run_params;
  std::vector<KernelBase*> kernels;
  std::vector<VariantID> variant_ids;
  std::vector<std::string> tuning_names[NumVariants];     // per variant: the ordered tunings to run and report
};

}  // namespace rajaperf
