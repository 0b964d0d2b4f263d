// RAJAPerfSuiteDriver.cpp -- main() of raja-perf-b200.exe: the five Executor steps of the reference
// driver (src/RAJAPerfSuiteDriver.cpp:36-49), without MPI / Kokkos initialisation.
#include <iostream>

#include "common/Executor.hpp"

int main(int argc, char** argv)
{
  rajaperf::Executor executor(argc, argv);   // STEP 1: parse the command line
  executor.setupSuite();                     // STEP 2: assemble kernels and variants
  executor.reportRunSummary(rajaperf::getCout());   // STEP 3
  executor.runSuite();                       // STEP 4
  executor.outputRunData();                  // STEP 5
  rajaperf::getCout() << "\n\nDONE!!!...." << std::endl;
  return executor.getRunParams().getInputState() == rajaperf::RunParams::BadInput ? 1 : 0;
}
