// incumbent.cu -- what the reference's CUDA variants would call on this box, timed beside Base_B200:
// CUB 2.8.2 (the toolkit's, which RAJA uses with CUDA >= 11: tpl/RAJA/cmake/SetupPackages.cmake:19-27)
// DeviceRadixSort::SortKeys / SortPairs on doubles (RAJA::sort / sort_pairs, policy/cuda/sort.hpp:86-146,
// 337-409), DeviceScan::ExclusiveSum (SCAN-Cuda.cpp:95-120), DeviceReduce::Sum (REDUCE_SUM-Cuda.cpp:80-106),
// and the reference-style one-element-per-thread TRIAD launch shape (TRIAD-Cuda.cpp:26-58, block 256).
// Measurement tool only (not part of the product):  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void triad_1pt(double* a, const double* b, const double* c, double alpha, long n)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) a[i] = b[i] + alpha * c[i];
}

template <typename F, typename S>
static double time_ms(F fn, S setup, int reps)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) { setup(); fn(); }
  CK(cudaDeviceSynchronize());
  double tot = 0;
  for (int r = 0; r < reps; ++r) {
    setup();
    CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); tot += ms;
  }
  return tot / reps;
}

int main(int argc, char** argv)
{
  const long n = argc > 1 ? atol(argv[1]) : (1L << 27);
  std::vector<double> h(n);
  srand(4793);
  for (long i = 0; i < n; ++i) h[i] = (double)rand() / RAND_MAX;
  double *src, *k0, *k1, *v0, *v1;
  CK(cudaMalloc(&src, n * 8)); CK(cudaMalloc(&k0, n * 8)); CK(cudaMalloc(&k1, n * 8)); CK(cudaMalloc(&v0, n * 8)); CK(cudaMalloc(&v1, n * 8));
  CK(cudaMemcpy(src, h.data(), n * 8, cudaMemcpyHostToDevice));
  void* tmp = nullptr; size_t tb = 0, need = 0;
  cub::DoubleBuffer<double> dk(k0, k1), dv(v0, v1);
  cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, 64); tb = need;
  cub::DeviceScan::ExclusiveSum(nullptr, need, src, k1, (int)n); if (need > tb) tb = need;
  cub::DeviceReduce::Sum(nullptr, need, src, v1, (int)n); if (need > tb) tb = need;
  CK(cudaMalloc(&tmp, tb));
  auto reset = [&]() { CK(cudaMemcpyAsync(k0, src, n * 8, cudaMemcpyDeviceToDevice)); CK(cudaMemcpyAsync(v0, src, n * 8, cudaMemcpyDeviceToDevice)); dk = cub::DoubleBuffer<double>(k0, k1); dv = cub::DoubleBuffer<double>(v0, v1); };
  auto none = []() {};
  double ms;
  ms = time_ms([&]() { size_t t = tb; cub::DeviceRadixSort::SortKeys(tmp, t, dk, (int)n, 0, 64); }, reset, 5);
  printf("{\"name\": \"cub::DeviceRadixSort::SortKeys<double>\", \"n\": %ld, \"ms\": %.4f, \"mkeys_per_s\": %.0f}\n", n, ms, n / ms / 1e3);
  ms = time_ms([&]() { size_t t = tb; cub::DeviceRadixSort::SortPairs(tmp, t, dk, dv, (int)n, 0, 64); }, reset, 5);
  printf("{\"name\": \"cub::DeviceRadixSort::SortPairs<double,double>\", \"n\": %ld, \"ms\": %.4f, \"mkeys_per_s\": %.0f}\n", n, ms, n / ms / 1e3);
  ms = time_ms([&]() { size_t t = tb; cub::DeviceScan::ExclusiveSum(tmp, t, src, k1, (int)n); }, none, 20);
  printf("{\"name\": \"cub::DeviceScan::ExclusiveSum<double>\", \"n\": %ld, \"ms\": %.4f, \"gbs\": %.1f}\n", n, ms, 16.0 * n / ms / 1e6);
  ms = time_ms([&]() { size_t t = tb; cub::DeviceReduce::Sum(tmp, t, src, v1, (int)n); }, none, 20);
  printf("{\"name\": \"cub::DeviceReduce::Sum<double>\", \"n\": %ld, \"ms\": %.4f, \"gbs\": %.1f}\n", n, ms, 8.0 * n / ms / 1e6);
  ms = time_ms([&]() { triad_1pt<<<(unsigned)((n + 255) / 256), 256>>>(k1, src, v0, 0.3, n); }, none, 20);
  printf("{\"name\": \"reference-shape TRIAD (1 element/thread, block 256)\", \"n\": %ld, \"ms\": %.4f, \"gbs\": %.1f}\n", n, ms, 24.0 * n / ms / 1e6);
  ms = time_ms([&]() { CK(cudaMemcpyAsync(k1, src, n * 8, cudaMemcpyDeviceToDevice)); }, none, 20);
  printf("{\"name\": \"cudaMemcpy D2D\", \"n\": %ld, \"ms\": %.4f, \"gbs\": %.1f}\n", n, ms, 16.0 * n / ms / 1e6);
  return 0;
}
