#include <cstdint>
__global__ void k(const double* __restrict__ a, double* __restrict__ c, long n){
  long i = (blockIdx.x*(long)blockDim.x+threadIdx.x)*4;
  double x,y,z,w;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x),"=d"(y),"=d"(z),"=d"(w) : "l"(a+i));
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(c+i),"d"(x),"d"(y),"d"(z),"d"(w));
}
