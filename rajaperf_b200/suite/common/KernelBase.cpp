#include "KernelBase.hpp"

#include <cstdio>
#include <cstdlib>
#include <iostream>

#include <cuda_runtime_api.h>

namespace rajaperf {

// one context per process and device: the analogue of the camp::resources::Cuda handle each kernel
// obtains through getCudaResource() (KernelBase.hpp:242-248)
static rpb200_ctx* g_ctx = nullptr;
static int g_ctx_device = -1;

rpb200_ctx* KernelBase::ctx()
{
  const int dev = run_params.getDevice();
  if (!g_ctx || g_ctx_device != dev) {
    if (g_ctx) rpb200_destroy(g_ctx);
    g_ctx = nullptr;
    checkAbi(rpb200_create(dev, &g_ctx), "rpb200_create (a B200 / sm_100 device is required: no fallback)");
    g_ctx_device = dev;
  }
  return g_ctx;
}

KernelBase::KernelBase(KernelID kid, const RunParams& params) : run_params(params), kernel_id(kid), name(getFullKernelName(kid)) {}

KernelBase::~KernelBase() { if (dev_timer) rpb200_timer_destroy(dev_timer); }

Index_type KernelBase::getTargetProblemSize() const
{
  Index_type target = 0;
  if (run_params.getSizeMeaning() == RunParams::SizeMeaning::Factor)
    target = static_cast<Index_type>(default_prob_size * run_params.getSizeFactor());
  else if (run_params.getSizeMeaning() == RunParams::SizeMeaning::Direct)
    target = static_cast<Index_type>(run_params.getSize());
  return target;
}

Index_type KernelBase::getRunReps() const
{
  if (run_params.getInputState() == RunParams::CheckRun) return static_cast<Index_type>(run_params.getCheckRunReps());
  return static_cast<Index_type>(default_reps * run_params.getRepFactor());
}

// KernelBase.cpp:138-232: a variant without a tuning list never runs, so defining it also sizes
// the per-tuning bookkeeping
void KernelBase::setVariantDefined(VariantID vid)
{
  if (!isVariantAvailable(vid)) return;
  if (vid == Base_B200) setB200TuningDefinitions(vid);
  const size_t n = variant_tuning_names[vid].size();
  checksum[vid].assign(n, 0.0);
  num_exec[vid].assign(n, 0);
  min_time[vid].assign(n, std::numeric_limits<double>::max());
  max_time[vid].assign(n, -std::numeric_limits<double>::max());
  tot_time[vid].assign(n, 0.0);
  min_dev_time[vid].assign(n, std::numeric_limits<double>::max());
}

size_t KernelBase::getVariantTuningIndex(VariantID vid, const std::string& tuning_name) const
{
  const std::vector<std::string>& names = variant_tuning_names[vid];
  for (size_t t = 0; t < names.size(); ++t) if (names[t] == tuning_name) return t;
  return std::string::npos;
}

void KernelBase::execute(VariantID vid, size_t tune_idx)
{
  running_variant = vid;
  running_tuning = tune_idx;
  detail::resetDataInitCount();
  // a non-default tuning is a launch shape of the library: set it for this run only
  const bool shaped = vid == Base_B200 && tune_idx < b200_tunings.size() &&
                      (b200_tunings[tune_idx].block_size > 0 || b200_tunings[tune_idx].ctas_per_sm >= 0 || b200_tunings[tune_idx].unroll > 0);
  if (shaped) {
    const LaunchShape& t = b200_tunings[tune_idx];
    checkAbi(rpb200_set_tuning(ctx(), tuningKernelName().c_str(), t.block_size, t.ctas_per_sm, t.unroll), "rpb200_set_tuning");
  }
  this->setUp(vid, tune_idx);
  this->runKernel(vid, tune_idx);
  this->updateChecksum(vid, tune_idx);
  this->tearDown(vid, tune_idx);
  if (shaped) checkAbi(rpb200_reset_tuning(ctx(), tuningKernelName().c_str()), "rpb200_reset_tuning");
  running_variant = NumVariants;
  running_tuning = std::numeric_limits<size_t>::max();
}

void KernelBase::runKernel(VariantID vid, size_t tune_idx)
{
  if (!hasVariantDefined(vid)) return;
  switch (vid) {
    case Base_B200:
      runB200Variant(vid, tune_idx);
      break;
    default:
      getCout() << "\n  " << getName() << " : Unknown variant id = " << vid << std::endl;
  }
}

void KernelBase::synchronize()
{
  if (isVariantGPU(running_variant)) {
    const int first = run_params.getDevice();
    int ndev = 1;
    cudaGetDeviceCount(&ndev);
    const int used = (kernel_id == Comm_HALO_EXCHANGE_FUSED || kernel_id == Comm_HALO_EXCHANGE || kernel_id == Comm_HALO_SENDRECV) ? std::min(ndev - first, run_params.getNumRanks()) : 1;
    for (int d = 0; d < used; ++d) {
      cudaSetDevice(first + d);
      checkAbi(rpb200_device_synchronize(), "rpb200_device_synchronize");
    }
    cudaSetDevice(first);
  }
}

void KernelBase::startTimer()
{
  synchronize();
  if (!dev_timer) checkAbi(rpb200_timer_create(&dev_timer), "rpb200_timer_create");
  t_start = std::chrono::steady_clock::now();
  checkAbi(rpb200_timer_start(dev_timer, stream()), "rpb200_timer_start");
}

void KernelBase::stopTimer()
{
  checkAbi(rpb200_timer_stop(dev_timer, stream()), "rpb200_timer_stop");
  synchronize();
  const auto t_stop = std::chrono::steady_clock::now();
  float ms = 0.0f;
  checkAbi(rpb200_timer_elapsed_ms(dev_timer, &ms), "rpb200_timer_elapsed_ms");
  recordExecTime(std::chrono::duration<double>(t_stop - t_start).count(), ms * 1e-3);
}

void KernelBase::recordExecTime(double host_s, double dev_s)
{
  const VariantID v = running_variant;
  const size_t t = running_tuning;
  num_exec[v][t]++;
  min_time[v][t] = std::min(min_time[v][t], host_s);
  max_time[v][t] = std::max(max_time[v][t], host_s);
  tot_time[v][t] += host_s;
  min_dev_time[v][t] = std::min(min_dev_time[v][t], dev_s);
}

void KernelBase::runRepLoop()
{
  const Index_type run_reps = getRunReps();
  if (run_params.useGraph() && m_graph_ok && run_reps > 0) {
    // "capture launch-bound inner loops in CUDA graphs": the whole rep batch becomes one launch
    (void)ctx();                      // the context (and its scratch) must exist before capture starts
    cudaStream_t cap;
    cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
    checkAbi(rpb200_stream_attach(ctx(), cap), "rpb200_stream_attach");      // its scratch set exists before the capture starts
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed);
    for (RepIndex_type irep = 0; irep < run_reps; ++irep) enqueueRep(cap);
    if (cudaStreamEndCapture(cap, &graph) != cudaSuccess || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      std::fprintf(stderr, "\n%s: CUDA graph capture failed\n", getName().c_str());
      std::abort();
    }
    startTimer();
    cudaGraphLaunch(exec, static_cast<cudaStream_t>(stream()));
    finishReps();
    stopTimer();
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    checkAbi(rpb200_stream_detach(ctx(), cap), "rpb200_stream_detach");
    cudaStreamDestroy(cap);
    return;
  }
  startTimer();
  for (RepIndex_type irep = 0; irep < run_reps; ++irep) enqueueRep(stream());
  finishReps();
  stopTimer();
}

}  // namespace rajaperf
