// How much of the IMAD.WIDE pipe does a kernel keep busy as a function of resident warps per SM and of the number of
// independent multiplication chains per thread?  Decides the shape of the batched-affine bucket kernels:
//   mul1 / mul2  : 1 / 2 independent chains of dependent 381-bit Montgomery multiplications per thread;
//   affine       : the dependency graph of one batched-affine addition per iteration (5 mul + 1 sqr, operands in
//                  registers): inv_d = I * pre; I = I * d; lam = num * inv_d; x3 = lam^2 - x1 - x2; y3 = lam (x1 - x3) - y1;
//   xyzz         : one XYZZ mixed addition (madd-2008-s, 8 mul + 2 sqr) per iteration.
// Reported: G field multiplications per second and the fraction of the 9.3 T IMAD.WIDE/s pipe (300 per 381-bit product).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o chains chains.cu && ./chains
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ckb_zkp_b200/csrc/curve.cuh"
using namespace zkb;
using F = Fp<BlsFq>;

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) k_chain(const F* in, F* out, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  F a = in[t], b = in[t + 1], c = in[t + 2], d = in[t + 3], e = in[t + 4];
  if (MODE == 0) {
    for (int i = 0; i < iters; i++) a = F::mul(a, b);
  } else if (MODE == 1) {
    for (int i = 0; i < iters; i++) { a = F::mul(a, b); c = F::mul(c, b); }
  } else if (MODE == 2) {
    // a = running inverse, b = prefix product / denominator stand-ins, (c, d) = point 1, e = x2
    for (int i = 0; i < iters; i++) {
      F inv_d = F::mul(a, b);
      a = F::mul(a, e);
      F lam = F::mul(F::sub(d, c), inv_d);
      F x3 = F::sub(F::sub(F::sqr(lam), c), e);
      d = F::sub(F::mul(lam, F::sub(c, x3)), d);
      c = x3;
    }
  } else {
    XYZZ<F> acc;
    acc.X = a; acc.Y = b; acc.ZZ = c; acc.ZZZ = d;
    F px = e, py = F::add(e, a);
    for (int i = 0; i < iters; i++) { acc.madd_xy(px, py, false); px = F::add(px, acc.ZZ); }
    a = acc.X; c = acc.Y; d = F::add(acc.ZZ, acc.ZZZ);
  }
  out[t] = F::add(F::add(a, c), d);
}

template <int MODE, int THREADS>
static void run(const char* name, int blocks_per_sm, double muls_per_iter) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = sms * blocks_per_sm * THREADS, iters = MODE == 3 ? 400 : 1000;
  F *in, *out;
  cudaMalloc(&in, sizeof(F) * (threads + 8));
  cudaMalloc(&out, sizeof(F) * threads);
  cudaMemset(in, 0x11, sizeof(F) * (threads + 8));
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_chain<MODE, THREADS>, THREADS, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k_chain<MODE, THREADS>);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<MODE, THREADS><<<threads / THREADS, THREADS>>>(in, out, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_chain<MODE, THREADS><<<threads / THREADS, THREADS>>>(in, out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)threads * iters * muls_per_iter;
  double gmul = muls / ms / 1e6;
  printf("{\"kernel\": \"%s\", \"threads_per_block\": %d, \"blocks_per_sm\": %d, \"warps_per_sm\": %d, \"regs\": %d, \"max_blocks_per_sm\": %d, "
         "\"ms\": %.3f, \"gmul_per_s\": %.2f, \"imad_pipe_frac\": %.3f, \"err\": \"%s\"}\n",
         name, THREADS, blocks_per_sm, blocks_per_sm * THREADS / 32, fa.numRegs, occ, ms, gmul, gmul * 300.0 / 9300.0,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(in); cudaFree(out);
}

int main() {
  for (int bps : {1, 2, 3, 4, 6, 8}) {
    run<0, 128>("mul1", bps, 1);
    run<1, 128>("mul2", bps, 2);
    run<2, 128>("affine", bps, 6);
    run<3, 128>("xyzz", bps, 10);
  }
  for (int bps : {1, 2, 4}) {          // the same warps per SM in fewer, larger blocks does not matter; one warp per block does
    run<2, 32>("affine", bps * 4, 6);
  }
  return 0;
}
