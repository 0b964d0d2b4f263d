#include "RunParams.hpp"

#include <cstdlib>
#include <iostream>

namespace rajaperf {

RunParams::RunParams(int argc, char** argv) { parseCommandLineOptions(argc, argv); }

static bool is_opt(const std::string& s) { return !s.empty() && s[0] == '-'; }

void RunParams::parseCommandLineOptions(int argc, char** argv)
{
  getCout() << "\n\nReading command line input..." << std::endl;
  auto need_value = [&](int& i, const std::string& opt, double lo, double& out) {
    if (i + 1 < argc && !is_opt(argv[i + 1])) {
      out = std::atof(argv[++i]);
      if (out < lo) { getCout() << "\nBad input: " << opt << " value too small" << std::endl; input_state = BadInput; }
    } else {
      getCout() << "\nBad input: must give " << opt << " a value" << std::endl;
      input_state = BadInput;
    }
  };
  for (int i = 1; i < argc; ++i) {
    const std::string opt(argv[i]);
    double v = 0.0;
    if (opt == "--help" || opt == "-h") { printHelpMessage(getCout()); input_state = InfoRequest; }
    else if (opt == "--print-kernels" || opt == "-pk") { printKernelNames(getCout()); input_state = InfoRequest; }
    else if (opt == "--print-variants" || opt == "-pv") { printVariantNames(getCout()); input_state = InfoRequest; }
    else if (opt == "--npasses") { need_value(i, opt, 1, v); npasses = (int)v; }
    else if (opt == "--repfact") { need_value(i, opt, 0, v); rep_fact = v; }
    else if (opt == "--sizefact") {              // RunParams.cpp:413-438: exclusive with --size
      if (size_seen) { getCout() << "\nBad input: may only set one of --size and --sizefact" << std::endl; input_state = BadInput; }
      need_value(i, opt, 0, v); size_factor = v; size_meaning = SizeMeaning::Factor; sizefact_seen = true;
      if (v <= 0.0) { getCout() << "\nBad input: --sizefact must be > 0" << std::endl; input_state = BadInput; }
    }
    else if (opt == "--size") {                  // RunParams.cpp:440-465
      if (sizefact_seen) { getCout() << "\nBad input: may only set one of --size and --sizefact" << std::endl; input_state = BadInput; }
      need_value(i, opt, 0, v); size = v; size_meaning = SizeMeaning::Direct; size_seen = true;
      if (v <= 0.0) { getCout() << "\nBad input: --size must be > 0" << std::endl; input_state = BadInput; }
    }
    else if (opt == "--kernels" || opt == "-k") {
      while (i + 1 < argc && !is_opt(argv[i + 1])) kernel_input.push_back(argv[++i]);
    }
    else if (opt == "--variants" || opt == "-v") {
      while (i + 1 < argc && !is_opt(argv[i + 1])) variant_input.push_back(argv[++i]);
    }
    else if (opt == "--exclude-kernels" || opt == "-ek") {       // RunParams.cpp:853-867
      while (i + 1 < argc && !is_opt(argv[i + 1])) exclude_kernel_input.push_back(argv[++i]);
    }
    else if (opt == "--exclude-variants" || opt == "-ev") {      // RunParams.cpp:885-899
      while (i + 1 < argc && !is_opt(argv[i + 1])) exclude_variant_input.push_back(argv[++i]);
    }
    else if (opt == "--tunings" || opt == "-t") {                // RunParams.cpp:1023-1037
      while (i + 1 < argc && !is_opt(argv[i + 1])) tuning_input.push_back(argv[++i]);
    }
    else if (opt == "--exclude-tunings" || opt == "-et") {       // RunParams.cpp:1039-1053
      while (i + 1 < argc && !is_opt(argv[i + 1])) exclude_tuning_input.push_back(argv[++i]);
    }
    else if (opt == "--outdir" || opt == "-od") { if (i + 1 < argc && !is_opt(argv[i + 1])) outdir = argv[++i]; }
    else if (opt == "--outfile" || opt == "-of") { if (i + 1 < argc && !is_opt(argv[i + 1])) outfile_prefix = argv[++i]; }
    else if (opt == "--halo_width") { need_value(i, opt, 1, v); halo_width = (Index_type)v; }
    else if (opt == "--halo_num_vars") { need_value(i, opt, 1, v); halo_num_vars = (Index_type)v; }
    else if (opt == "--ltimes_num_d") { need_value(i, opt, 1, v); ltimes_num_d = (Index_type)v; }
    else if (opt == "--ltimes_num_g") { need_value(i, opt, 1, v); ltimes_num_g = (Index_type)v; }
    else if (opt == "--ltimes_num_m") { need_value(i, opt, 1, v); ltimes_num_m = (Index_type)v; }
    else if (opt == "--mpi_3d_division") {       // RunParams.cpp:777-806
      for (int d = 0; d < 3; ++d) { need_value(i, opt, 1, v); mpi_3d_division[d] = (int)v; }
    }
    else if (opt == "--device") { need_value(i, opt, 0, v); device = (int)v; }
    else if (opt == "--graph") { use_graph = true; }
    else if (opt == "--dryrun") { if (input_state != BadInput) input_state = DryRun; }
    else if (opt == "--checkrun") {
      if (input_state != BadInput) input_state = CheckRun;
      if (i + 1 < argc && !is_opt(argv[i + 1])) checkrun_reps = std::atoi(argv[++i]);
    }
    else if (opt == "--disable-warmup") { disable_warmup = true; }
    else if (opt == "--show-progress" || opt == "-sp") { show_progress = true; }
    else {
      getCout() << "\nBad input: unknown option '" << opt << "'" << std::endl;
      input_state = BadInput;
    }
  }
  if (input_state == Undefined) input_state = PerfRun;
  processKernelInput();
  processVariantInput();
  if (input_state == BadInput) {
    if (!invalid_kernel_input.empty()) { getCout() << "\nInvalid kernel input:"; for (auto& s : invalid_kernel_input) getCout() << ' ' << s; getCout() << std::endl; }
    if (!invalid_variant_input.empty()) { getCout() << "\nInvalid variant input:"; for (auto& s : invalid_variant_input) getCout() << ' ' << s; getCout() << std::endl; }
  }
}

// group name | kernel name | full kernel name -> std::set<KernelID> (RunParams.cpp:1893-2102)
static bool matchKernels(const std::string& in, std::set<KernelID>& out)
{
  for (int g = 0; g < NumGroups; ++g)
    if (getGroupName((GroupID)g) == in) {
      for (int k = 0; k < NumKernels; ++k) if (getKernelGroup((KernelID)k) == (GroupID)g) out.insert((KernelID)k);
      return true;
    }
  for (int k = 0; k < NumKernels; ++k)
    if (getKernelName((KernelID)k) == in || getFullKernelName((KernelID)k) == in) { out.insert((KernelID)k); return true; }
  return false;
}

void RunParams::processKernelInput()
{
  std::set<KernelID> excluded;                                   // RunParams.cpp:1903-1972
  for (const std::string& in : exclude_kernel_input)
    if (!matchKernels(in, excluded)) { invalid_kernel_input.push_back(in); input_state = BadInput; }
  if (kernel_input.empty()) {
    for (int k = 0; k < NumKernels; ++k) run_kernels.insert((KernelID)k);
  } else {
    for (const std::string& in : kernel_input)
      if (!matchKernels(in, run_kernels)) { invalid_kernel_input.push_back(in); input_state = BadInput; }
  }
  for (KernelID k : excluded) run_kernels.erase(k);
}

// RunParams.cpp:2295-2440: requested ∩ available; unknown names are bad input, known-but-unavailable
// names are reported and dropped (as when the reference is built without that back-end).
void RunParams::processVariantInput()
{
  std::set<VariantID> excluded;                                  // RunParams.cpp:2305-2345
  for (const std::string& in : exclude_variant_input) {
    bool found = false;
    for (int v = 0; v < NumVariants; ++v) if (getVariantName((VariantID)v) == in) { excluded.insert((VariantID)v); found = true; }
    if (!found) { invalid_variant_input.push_back(in); input_state = BadInput; }
  }
  if (variant_input.empty()) {
    for (int v = 0; v < NumVariants; ++v)
      if (isVariantAvailable((VariantID)v) && !excluded.count((VariantID)v)) run_variants.insert((VariantID)v);
    return;
  }
  for (const std::string& in : variant_input) {
    bool found = false;
    for (int v = 0; v < NumVariants; ++v)
      if (getVariantName((VariantID)v) == in) {
        found = true;
        if (isVariantAvailable((VariantID)v)) { if (!excluded.count((VariantID)v)) run_variants.insert((VariantID)v); }
        else getCout() << "\nVariant " << in << " is not available in this build (CPU variants: run the reference binary)" << std::endl;
      }
    if (!found) { invalid_variant_input.push_back(in); input_state = BadInput; }
  }
}

void RunParams::print(std::ostream& str) const
{
  if (!invalid_tuning_input.empty()) { str << "\n Invalid tuning input:"; for (auto& s : invalid_tuning_input) str << ' ' << s; }
  str << "\n npasses = " << npasses << "\n rep_fact = " << rep_fact << "\n size_meaning = "
      << (size_meaning == SizeMeaning::Direct ? "Direct" : "Factor") << "\n size = " << size << "\n size_factor = " << size_factor
      << "\n checkrun_reps = " << checkrun_reps << "\n halo_width = " << halo_width << "\n halo_num_vars = " << halo_num_vars
      << "\n ltimes_num_d,g,m = " << ltimes_num_d << "," << ltimes_num_g << "," << ltimes_num_m
      << "\n mpi_3d_division = " << mpi_3d_division[0] << " " << mpi_3d_division[1] << " " << mpi_3d_division[2]
      << "\n outdir = " << outdir << "\n outfile_prefix = " << outfile_prefix << "\n graph = " << use_graph << std::endl;
}

void RunParams::printHelpMessage(std::ostream& str) const
{
  str << "\nUsage: ./raja-perf-b200.exe [options]\nValid options are:\n\n"
      << "\t --help, -h (print options with descriptions)\n"
      << "\t --print-kernels, -pk / --print-variants, -pv\n"
      << "\t --kernels, -k <space-separated strings> (group names, kernel names or full names; default: all)\n"
      << "\t --exclude-kernels, -ek <space-separated strings> (same name forms; removed from the selection)\n"
      << "\t --variants, -v <space-separated strings> (default: every available variant: Base_B200)\n"
      << "\t --exclude-variants, -ev <space-separated strings>\n"
      << "\t --tunings, -t <space-separated strings> (default: every tuning a kernel defines; 'default' = the measured best)\n"
      << "\t --exclude-tunings, -et <space-separated strings>\n"
      << "\t --npasses <int> (passes through the suite; default 1)\n"
      << "\t --repfact <double> (multiplies each kernel's default rep count)\n"
      << "\t --size <int> (problem size of every kernel run) | --sizefact <double> (multiplies each default size)\n"
      << "\t --checkrun <int> (run each kernel that many reps, default 1: a quick correctness run)\n"
      << "\t --dryrun (print the summary, run nothing)\n"
      << "\t --outdir, -od <dir>   --outfile, -of <prefix>\n"
      << "\t --halo_width <int> --halo_num_vars <int> (Comm kernels; defaults 1 and 3)\n"
      << "\t --ltimes_num_d <int> --ltimes_num_g <int> --ltimes_num_m <int> (defaults 64 32 25)\n"
      << "\t --mpi_3d_division <int> <int> <int> (rank grid of HALO_EXCHANGE_FUSED; ranks are dealt to the visible GPUs)\n"
      << "\t --device <int> (first CUDA device)   --graph (capture each rep loop in one CUDA graph)\n"
      << "\t --disable-warmup   --show-progress, -sp\n" << std::endl;
}

void RunParams::printKernelNames(std::ostream& str) const
{
  str << "\nAvailable kernels:\n------------------\n";
  for (int k = 0; k < NumKernels; ++k) str << getFullKernelName((KernelID)k) << std::endl;
}

void RunParams::printVariantNames(std::ostream& str) const
{
  str << "\nAvailable variants:\n-------------------\n";
  for (int v = 0; v < NumVariants; ++v) if (isVariantAvailable((VariantID)v)) str << getVariantName((VariantID)v) << std::endl;
}

}  // namespace rajaperf
