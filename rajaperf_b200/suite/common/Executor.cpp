#include "Executor.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>

#include "KernelBase.hpp"

namespace rajaperf {

// roofline denominators: nominal B200 HBM3e and the copy bandwidth measured on this pool's B200s
static constexpr double kNominalGBs = 8000.0;
static constexpr double kMeasuredCopyGBs = 6540.2;

Executor::Executor(int argc, char** argv) : run_params(argc, argv) {}

Executor::~Executor()
{
  for (KernelBase* k : kernels) delete k;
}

void Executor::setupSuite()
{
  const RunParams::InputOpt in_state = run_params.getInputState();
  if (in_state == RunParams::InfoRequest || in_state == RunParams::BadInput) return;
  getCout() << "\nSetting up suite based on input..." << std::endl;
  for (KernelID kid : run_params.getKernelIDsToRun()) kernels.push_back(getKernelObject(kid, run_params));
  for (VariantID vid : run_params.getVariantIDsToRun()) variant_ids.push_back(vid);   // std::set => enum order

  // Executor.cpp:290-358: per variant the ordered union of the tuning names the selected kernels define, filtered by
  // --tunings / --exclude-tunings, "default" first.  A name no selected kernel defines is bad input.
  const std::vector<std::string>& selected = run_params.getTuningInput();
  const std::vector<std::string>& excluded = run_params.getExcludeTuningInput();
  auto listed = [](const std::vector<std::string>& v, const std::string& s) { return std::find(v.begin(), v.end(), s) != v.end(); };
  std::vector<std::string> known;
  for (VariantID vid : variant_ids) {
    std::vector<std::string>& names = tuning_names[vid];
    for (KernelBase* k : kernels)
      for (const std::string& t : k->getVariantTuningNames(vid)) {
        if (listed(names, t)) continue;
        if ((selected.empty() || listed(selected, t)) && !listed(excluded, t)) names.push_back(t);
      }
    auto def = std::find(names.begin(), names.end(), KernelBase::getDefaultTuningName());
    if (def != names.end()) std::rotate(names.begin(), def, def + 1);
  }
  for (KernelBase* k : kernels)                    // a name is valid if any variant of a selected kernel defines it
    for (int v = 0; v < NumVariants; ++v)
      for (const std::string& t : k->getVariantTuningNames((VariantID)v)) if (!listed(known, t)) known.push_back(t);
  std::vector<std::string> invalid;
  for (const std::vector<std::string>* in : {&selected, &excluded})
    for (const std::string& t : *in) if (!listed(known, t) && !listed(invalid, t)) invalid.push_back(t);
  if (!invalid.empty()) {
    run_params.setInvalidTuningInput(invalid);
    run_params.setInputState(RunParams::BadInput);
  }
}

void Executor::reportRunSummary(std::ostream& str) const
{
  const RunParams::InputOpt in_state = run_params.getInputState();
  if (in_state == RunParams::BadInput) {
    str << "\nRunParams state:\n----------------";
    run_params.print(str);
    str << "\n\nSuite will not be run now due to bad input.\n  See run parameters or option messages above.\n" << std::endl;
    return;
  }
  if (in_state != RunParams::PerfRun && in_state != RunParams::DryRun && in_state != RunParams::CheckRun) return;
  if (in_state == RunParams::DryRun) {
    str << "\n\nRAJA performance suite dry run summary....\n--------------------------------------\n\nInput state:";
    run_params.print(str);
  }
  if (in_state == RunParams::PerfRun || in_state == RunParams::CheckRun) {
    str << "\n\nRAJA performance suite run summary....\n--------------------------------------" << std::endl;
    if (in_state == RunParams::CheckRun) str << "\nThe suite will run in a check mode: each kernel runs " << run_params.getCheckRunReps() << " rep(s)" << std::endl;
  }
  str << "\nHow suite will be run:\n\t # passes = " << run_params.getNumPasses() << "\n\t Kernel rep factor = " << run_params.getRepFactor()
      << "\n\t Output files will be named " << (run_params.getOutputDirName().empty() ? "." : run_params.getOutputDirName()) << "/"
      << run_params.getOutputFilePrefix() << "*" << std::endl;
  str << "\nThe following kernels and variants (when available for a kernel) will be run:" << std::endl;
  str << "\nVariants\n--------\n";
  for (VariantID v : variant_ids)                                  // Executor.cpp:488-496
    for (const std::string& t : tuningNames(v)) str << getVariantName(v) << "-" << t << std::endl;
  str << std::endl;
  writeKernelInfoSummary(str);
  str.flush();
}

void Executor::writeKernelInfoSummary(std::ostream& str) const    // Executor.cpp:508-643
{
  size_t w = 0;
  for (KernelBase* k : kernels) w = std::max(w, k->getName().size());
  w += 2;
  str << std::left << std::setw(w) << "Kernels" << std::right << std::setw(14) << "Problem size" << std::setw(10) << "Reps"
      << std::setw(16) << "Iterations/rep" << std::setw(13) << "Kernels/rep" << std::setw(16) << "Bytes/rep"
      << std::setw(16) << "FLOPS/rep" << std::endl;
  for (KernelBase* k : kernels)
    str << std::left << std::setw(w) << k->getName() << std::right << std::setw(14) << k->getActualProblemSize() << std::setw(10)
        << k->getRunReps() << std::setw(16) << k->getItsPerRep() << std::setw(13) << k->getKernelsPerRep() << std::setw(16)
        << k->getBytesPerRep() << std::setw(16) << k->getFLOPsPerRep() << std::endl;
}

void Executor::runSuite()
{
  const RunParams::InputOpt in_state = run_params.getInputState();
  if (in_state != RunParams::PerfRun && in_state != RunParams::CheckRun) return;
  runWarmupKernels();
  getCout() << "\n\nRunning specified kernels and variants...\n";
  const int npasses = run_params.getNumPasses();
  for (int ip = 0; ip < npasses; ++ip) {
    if (run_params.showProgress()) getCout() << "\nPass through suite # " << ip << "\n";
    for (KernelBase* kern : kernels) runKernel(kern, false);
  }
}

void Executor::runKernel(KernelBase* kern, bool print_kernel_name)    // Executor.cpp:679-726
{
  if (run_params.showProgress() || print_kernel_name) getCout() << "\nRun kernel -- " << kern->getName() << "\n";
  for (VariantID vid : variant_ids) {
    if (!kern->hasVariantDefined(vid)) {
      if (run_params.showProgress()) getCout() << "\tNo " << getVariantName(vid) << " variant" << std::endl;
      continue;
    }
    for (size_t t = 0; t < kern->getNumVariantTunings(vid); ++t) {          // Executor.cpp:697-722
      const std::string& tname = kern->getVariantTuningName(vid, t);
      const std::vector<std::string>& run = tuningNames(vid);
      if (std::find(run.begin(), run.end(), tname) == run.end()) {
        if (run_params.showProgress()) getCout() << "\t\tSkipping " << tname << " tuning" << std::endl;
        continue;
      }
      if (run_params.showProgress()) getCout() << "\tRunning " << getVariantName(vid) << "-" << tname << " variant" << std::endl;
      kern->execute(vid, t);
    }
  }
}

// The reference warms the device up with one kernel per feature in use (Executor.cpp:728-823).  Here
// every selected kernel object is executed once on a throw-away copy, so clocks, caches of lazily
// loaded cubins and the context's scratch are settled before the first timed pass.
void Executor::runWarmupKernels()
{
  if (run_params.getDisableWarmup()) return;
  getCout() << "\n\nRun warmup kernels...\n";
  for (KernelID kid : run_params.getKernelIDsToRun()) {
    KernelBase* w = getKernelObject(kid, run_params);
    runKernel(w, false);
    delete w;
  }
}

static void makeDir(const std::string& dir)
{
  if (dir.empty()) return;
  std::string cur;
  for (size_t i = 0; i <= dir.size(); ++i) {
    if (i == dir.size() || dir[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0755); }
    if (i < dir.size()) cur += dir[i];
  }
}

void Executor::outputRunData()
{
  const RunParams::InputOpt in_state = run_params.getInputState();
  if (in_state != RunParams::PerfRun && in_state != RunParams::CheckRun) return;
  getCout() << "\n\nGenerate run report files...\n";
  makeDir(run_params.getOutputDirName());
  const std::string base = (run_params.getOutputDirName().empty() ? std::string() : run_params.getOutputDirName() + "/") +
                           run_params.getOutputFilePrefix();
  writeTimingCSV(base + "-timing-Average.csv", 0);
  writeTimingCSV(base + "-timing-Minimum.csv", 1);
  writeTimingCSV(base + "-timing-Maximum.csv", 2);
  writeChecksumReport(base + "-checksum.txt");
  writeKernelsCSV(base + "-kernels.csv");
  writeBandwidthCSV(base + "-bandwidth.csv");
}

// Executor.cpp:901-998: one column per (variant, tuning); the entry is seconds per rep batch
void Executor::writeTimingCSV(const std::string& filename, int combiner)
{
  std::ofstream file(filename);
  if (!file) return;
  static const char* title[] = {"Mean Runtime Report (sec.) ", "Min Runtime Report (sec.) ", "Max Runtime Report (sec.) "};
  file << title[combiner] << std::endl;
  file << "Kernel";
  for (VariantID v : variant_ids) for (const std::string& tn : tuningNames(v)) file << ", " << getVariantName(v) << "-" << tn;
  file << std::endl;
  file << std::setprecision(9) << std::scientific;
  for (KernelBase* k : kernels) {
    file << k->getName();
    for (VariantID v : variant_ids)
      for (const std::string& tn : tuningNames(v)) {
        const size_t ti = k->getVariantTuningIndex(v, tn);
        if (ti != std::string::npos && k->wasVariantTuningRun(v, ti)) {
          const double t = combiner == 0 ? k->getTotTime(v, ti) / run_params.getNumPasses() : combiner == 1 ? k->getMinTime(v, ti) : k->getMaxTime(v, ti);
          file << ", " << t;
        } else file << ", Not run";
      }
    file << std::endl;
  }
}

// Executor.cpp:1281-1500: 20 significant digits; the first variant listed that ran is the reference
void Executor::writeChecksumReport(const std::string& filename)
{
  std::ofstream file(filename);
  if (!file) return;
  const std::string equal_line(99, '='), dash_line(88, '-'), dot_line(56, '.');
  const size_t prec = 20, checksum_width = prec + 8;
  size_t namecol_width = 0;
  for (KernelBase* k : kernels) namecol_width = std::max(namecol_width, k->getName().size());
  for (VariantID v : variant_ids) for (const std::string& tn : tuningNames(v)) namecol_width = std::max(namecol_width, getVariantName(v).size() + 1 + tn.size());
  namecol_width += 2;
  file << equal_line << std::endl;
  file << "Checksum Report " << std::endl;
  file << equal_line << std::endl;
  file << std::left << std::setw(namecol_width) << "Kernel  " << std::endl;
  file << dot_line << std::endl;
  file << std::left << std::setw(namecol_width) << "Variants  " << std::left << std::setw(checksum_width) << "Checksum  "
       << std::left << std::setw(checksum_width) << "Checksum Diff  " << std::endl;
  file << std::left << std::setw(namecol_width) << "  " << std::left << std::setw(checksum_width) << "  "
       << std::left << std::setw(checksum_width) << "(vs. first variant listed)  " << std::endl;
  file << dash_line << std::endl;
  for (KernelBase* k : kernels) {
    file << std::left << std::setw(namecol_width) << k->getName() << std::endl;
    file << dot_line << std::endl;
    Checksum_type ref = 0.0;                 // the first (variant, tuning) listed that ran (Executor.cpp:1359-1373)
    bool found = false;
    for (size_t i = 0; i < variant_ids.size() && !found; ++i)
      for (const std::string& tn : tuningNames(variant_ids[i])) {
        const size_t ti = k->getVariantTuningIndex(variant_ids[i], tn);
        if (ti != std::string::npos && k->wasVariantTuningRun(variant_ids[i], ti)) { ref = k->getChecksum(variant_ids[i], ti); found = true; break; }
      }
    for (VariantID v : variant_ids)
      for (const std::string& tn : tuningNames(v)) {
        const size_t ti = k->getVariantTuningIndex(v, tn);
        if (ti == std::string::npos) continue;               // the reference lists only tunings the kernel defines
        const std::string vname = getVariantName(v) + "-" + tn;
        if (k->wasVariantTuningRun(v, ti)) {
          const Checksum_type ck = k->getChecksum(v, ti);
          file << std::left << std::setw(namecol_width) << vname << std::showpoint << std::setprecision(prec) << std::left
               << std::setw(checksum_width) << ck << std::left << std::setw(checksum_width) << (ref - ck) << std::endl;
        } else {
          file << std::left << std::setw(namecol_width) << vname << std::left << std::setw(checksum_width) << "Not Run" << std::left
               << std::setw(checksum_width) << "Not Run" << std::endl;
        }
      }
    file << std::endl << dash_line << std::endl;
  }
}

void Executor::writeKernelsCSV(const std::string& filename)
{
  std::ofstream file(filename);
  if (!file) return;
  file << "Kernels, Problem size, Reps, Iterations/rep, Kernels/rep, Bytes/rep, FLOPS/rep" << std::endl;
  for (KernelBase* k : kernels)
    file << k->getName() << ", " << k->getActualProblemSize() << ", " << k->getRunReps() << ", " << k->getItsPerRep() << ", "
         << k->getKernelsPerRep() << ", " << k->getBytesPerRep() << ", " << k->getFLOPsPerRep() << std::endl;
}

// New with Base_B200 (the reference never turns Bytes/rep into a rate, Executor.cpp:626-638):
// GB/s = Bytes/rep x reps / time, from the best pass, for the host-clock time the reference reports and for
// the cudaEvent time of the rep batch; fractions against the nominal and the measured-copy roofline.
void Executor::writeBandwidthCSV(const std::string& filename)
{
  std::ofstream file(filename);
  std::ostream& out = getCout();
  const char* hdr = "Kernel, Variant-tuning, Problem size, Reps, Bytes/rep, FLOPs/rep, Host time (s), Device time (s), GB/s (device time), "
                    "GFLOP/s (device time), Fraction of 8000 GB/s, Fraction of measured 6540.2 GB/s copy";
  if (file) file << hdr << std::endl;
  out << "\nBandwidth report (best pass)\n" << std::left << std::setw(28) << "Kernel" << std::right << std::setw(14) << "ms/rep" << std::setw(12)
      << "GB/s" << std::setw(12) << "GFLOP/s" << std::setw(12) << "%8TB/s" << std::setw(12) << "%copy" << std::endl;
  for (KernelBase* k : kernels)
    for (VariantID v : variant_ids)
     for (const std::string& tn : tuningNames(v)) {
      const size_t ti = k->getVariantTuningIndex(v, tn);
      if (ti == std::string::npos || !k->wasVariantTuningRun(v, ti)) continue;
      const bool is_default = tn == KernelBase::getDefaultTuningName();
      const std::string label = is_default ? k->getName() : "  " + k->getName().substr(k->getName().find('_') + 1) + "-" + tn;
      const double reps = (double)k->getRunReps();
      const double th = k->getMinTime(v, ti), td = k->getMinDeviceTime(v, ti);
      const double gbs = td > 0 ? k->getBytesPerRep() * reps / td * 1e-9 : 0.0;
      const double gfs = td > 0 ? k->getFLOPsPerRep() * reps / td * 1e-9 : 0.0;
      if (file)
        file << k->getName() << ", " << getVariantName(v) << (is_default ? std::string() : "-" + tn) << ", " << k->getActualProblemSize() << ", " << k->getRunReps() << ", "
             << k->getBytesPerRep() << ", " << k->getFLOPsPerRep() << ", " << std::setprecision(9) << th << ", " << td << ", " << gbs << ", "
             << gfs << ", " << gbs / kNominalGBs << ", " << gbs / kMeasuredCopyGBs << std::endl;
      out << std::left << std::setw(28) << label << std::right << std::fixed << std::setprecision(4) << std::setw(14)
          << (reps > 0 ? td / reps * 1e3 : 0.0) << std::setprecision(1) << std::setw(12) << gbs << std::setw(12) << gfs << std::setw(12)
          << 100.0 * gbs / kNominalGBs << std::setw(12) << 100.0 * gbs / kMeasuredCopyGBs << std::endl;
      out.unsetf(std::ios::fixed);
    }
}

}  // namespace rajaperf
