// RunParams.hpp -- the driver's command line (reference: common/RunParams.{hpp,cpp}), hot-path subset:
//   -k/--kernels, -ek/--exclude-kernels, -v/--variants, -ev/--exclude-variants, -t/--tunings, -et/--exclude-tunings,
//   --size | --sizefact, --npasses, --repfact, --checkrun N, --dryrun,
//   -od/--outdir, -of/--outfile, --disable-warmup, -sp/--show-progress, -pk/--print-kernels,
//   -pv/--print-variants, --halo_width, --halo_num_vars, --ltimes_num_{d,g,m}, --mpi_3d_division,
//   plus --device N (first CUDA device of this process) and --graph (capture the rep loop in a CUDA graph).
#pragma once
#include <array>
#include <iosfwd>
#include <set>
#include <string>
#include <vector>

#include "RAJAPerfSuite.hpp"

namespace rajaperf {

class RunParams {
public:
  enum InputOpt { InfoRequest, DryRun, CheckRun, PerfRun, BadInput, Undefined };   // RunParams.hpp:44-53
  enum class SizeMeaning { Unset, Factor, Direct };

  RunParams(int argc, char** argv);

  InputOpt getInputState() const { return input_state; }
  int getNumPasses() const { return npasses; }
  double getRepFactor() const { return rep_fact; }
  int getCheckRunReps() const { return checkrun_reps; }
  SizeMeaning getSizeMeaning() const { return size_meaning; }
  double getSize() const { return size; }
  double getSizeFactor() const { return size_factor; }
  Index_type getHaloWidth() const { return halo_width; }
  Index_type getHaloNumVars() const { return halo_num_vars; }
  Index_type getLtimesNumD() const { return ltimes_num_d; }
  Index_type getLtimesNumG() const { return ltimes_num_g; }
  Index_type getLtimesNumM() const { return ltimes_num_m; }
  const std::array<int, 3>& getMPI3DDivision() const { return mpi_3d_division; }
  int getNumRanks() const { return mpi_3d_division[0] * mpi_3d_division[1] * mpi_3d_division[2]; }
  bool getDisableWarmup() const { return disable_warmup; }
  bool showProgress() const { return show_progress; }
  bool useGraph() const { return use_graph; }
  int getDevice() const { return device; }
  const std::string& getOutputDirName() const { return outdir; }
  const std::string& getOutputFilePrefix() const { return outfile_prefix; }
  const std::set<KernelID>& getKernelIDsToRun() const { return run_kernels; }
  const std::set<VariantID>& getVariantIDsToRun() const { return run_variants; }
  // tuning names are validated by the Executor once the kernel objects exist (Executor.cpp:290-358)
  const std::vector<std::string>& getTuningInput() const { return tuning_input; }
  const std::vector<std::string>& getExcludeTuningInput() const { return exclude_tuning_input; }
  void setInvalidTuningInput(const std::vector<std::string>& v) { invalid_tuning_input = v; }
  void setInputState(InputOpt s) { input_state = s; }

  void print(std::ostream& str) const;

private:
  void parseCommandLineOptions(int argc, char** argv);
  void processKernelInput();
  void processVariantInput();
  void printHelpMessage(std::ostream& str) const;
  void printKernelNames(std::ostream& str) const;
  void printVariantNames(std::ostream& str) const;

  InputOpt input_state = Undefined;
  int npasses = 1;
  double rep_fact = 1.0;
  int checkrun_reps = 1;
  SizeMeaning size_meaning = SizeMeaning::Factor;
  double size = 0.0, size_factor = 1.0;
  bool size_seen = false, sizefact_seen = false;
  Index_type halo_width = 1, halo_num_vars = 3;                     // RunParams.cpp:46-47
  Index_type ltimes_num_d = 64, ltimes_num_g = 32, ltimes_num_m = 25;   // RunParams.cpp:43-45
  std::array<int, 3> mpi_3d_division{{1, 1, 1}};
  bool disable_warmup = false, show_progress = false, use_graph = false;
  int device = 0;
  std::string outdir, outfile_prefix = "RAJAPerf";
  std::vector<std::string> kernel_input, variant_input, invalid_kernel_input, invalid_variant_input;
  std::vector<std::string> exclude_kernel_input, exclude_variant_input, tuning_input, exclude_tuning_input, invalid_tuning_input;
  std::set<KernelID> run_kernels;
  std::set<VariantID> run_variants;
};

}  // namespace rajaperf
