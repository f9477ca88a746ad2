// Field-multiplication throughput: Fp::mul (word-serial CIOS) vs Fp::mul_sos (Karatsuba product + separated
// reduction) for the 381-bit base field and the 255-bit scalar field.  Every thread runs two independent
// multiplication chains (the ILP a point-addition formula offers); enough blocks to fill the machine.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fieldmul fieldmul.cu && ./fieldmul
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ckb_zkp_b200/csrc/field.cuh"
using namespace zkb;

template <class F, int WHICH, int CHAINS>
__global__ void __launch_bounds__(128) k_chain(const F* in, F* out, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  F a[CHAINS], b = in[t];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) a[c] = in[t + c + 1];
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = WHICH ? F::mul_sos(a[c], b) : F::mul(a[c], b);
  }
  F r = a[0];
#pragma unroll
  for (int c = 1; c < CHAINS; c++) r = F::add(r, a[c]);
  out[t] = r;
}

template <class F, int WHICH, int CHAINS>
static double run(const char* name, int blocks_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = sms * blocks_per_sm * 128, iters = 2000;
  F *in, *out;
  cudaMalloc(&in, sizeof(F) * (threads + CHAINS + 1));
  cudaMalloc(&out, sizeof(F) * threads);
  cudaMemset(in, 0x5a, sizeof(F) * (threads + CHAINS + 1));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<F, WHICH, CHAINS><<<threads / 128, 128>>>(in, out, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_chain<F, WHICH, CHAINS><<<threads / 128, 128>>>(in, out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)threads * iters * CHAINS;
  printf("%-34s chains=%d blocks/SM=%d  %8.3f ms  %7.2f G mul/s  (%s)\n", name, CHAINS, blocks_per_sm, ms, muls / ms / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(in); cudaFree(out);
  return muls / ms / 1e6;
}

int main() {
  for (int bps : {2, 4}) {
    run<Fp<BlsFq>, 0, 2>("BlsFq mul (CIOS)", bps);
    run<Fp<BlsFq>, 1, 2>("BlsFq mul_sos (Karatsuba)", bps);
    run<Fp<BlsFq>, 0, 1>("BlsFq mul (CIOS)", bps);
    run<Fp<BlsFq>, 1, 1>("BlsFq mul_sos (Karatsuba)", bps);
    run<Fp<BlsFr>, 0, 2>("BlsFr mul (CIOS)", bps);
    run<Fp<BlsFr>, 1, 2>("BlsFr mul_sos (Karatsuba)", bps);
  }
  return 0;
}
