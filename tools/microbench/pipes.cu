// Microbenchmark: sustained throughput of IMAD.WIDE.U32 (carry chains as in field.cuh) vs DFMA on
// one B200, to decide whether a floating-point (52-bit limb) Montgomery multiplication could beat
// the integer one.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipes.cu -o pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int ILP>
__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t lo[ILP], hi[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { lo[i] = threadIdx.x + i; hi[i] = blockIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a + i), "r"(b));
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += lo[i] ^ hi[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) acc[i] = __fma_rz(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// both pipes at once: even warps integer, odd warps double
__global__ void k_mixed(uint32_t* out, double* outd, uint32_t a, uint32_t b, double da, double db, int iters) {
  constexpr int ILP = 8;
  if ((threadIdx.x >> 5) & 1) {
    double acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = __fma_rz(acc[i], da, db);
    double s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    outd[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    uint32_t lo[ILP], hi[ILP];
    for (int i = 0; i < ILP; i++) { lo[i] = threadIdx.x + i; hi[i] = blockIdx.x + i; }
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int i = 0; i < ILP; i++)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a + i), "r"(b));
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) s += lo[i] ^ hi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}
int main() {
  const int blocks = 148 * 4, threads = 256, iters = 4096;
  uint32_t* o; double* od;
  cudaMalloc(&o, blocks * threads * 4); cudaMalloc(&od, blocks * threads * 8);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  float ms;
  const double ops = (double)blocks * threads * iters * 8 * 8;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(s); k_imad<8><<<blocks, threads>>>(o, 12345u, 6789u, iters); cudaEventRecord(e); cudaEventSynchronize(e);
    cudaEventElapsedTime(&ms, s, e);
    if (rep) printf("IMAD.WIDE  : %.2f T/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaEventRecord(s); k_dfma<8><<<blocks, threads>>>(od, 1.0000001, 0.5, iters); cudaEventRecord(e); cudaEventSynchronize(e);
    cudaEventElapsedTime(&ms, s, e);
    if (rep) printf("DFMA       : %.2f T/s  (%.1f lanes/clk/SM)\n", ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaEventRecord(s); k_mixed<<<blocks, threads>>>(o, od, 12345u, 6789u, 1.0000001, 0.5, iters); cudaEventRecord(e); cudaEventSynchronize(e);
    cudaEventElapsedTime(&ms, s, e);
    if (rep) printf("mixed (half the warps each): %.3f ms for %.2f G IMAD.WIDE + %.2f G DFMA\n", ms, ops / 2 / 1e9, ops / 2 / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
