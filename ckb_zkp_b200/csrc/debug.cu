// Diagnostic entry points: run single field / group operations of the device code on arrays of
// operands, so the GPU arithmetic can be checked against the CPU oracle in isolation
// (tests/test_gpu_field.py).  Not used by the prove path.
#include "common.cuh"
#include "curve.cuh"
#include "devutil.cuh"

namespace zkb {

template <class F>
__global__ void k_dbg_fp(int op, const F* a, const F* b, F* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = a[i], y = b[i], r;
  switch (op) {
    case 0: r = F::mul(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::inv(x); break;
    case 4: r = F::to_mont(x); break;
    case 5: r = F::from_mont(x); break;
    case 6: r = F::sqr(x); break;
    case 7: r = F::neg(x); break;
    default: r = F::zero();
  }
  out[i] = r;
}

// op 0: acc.madd(q affine, neg)  1: acc.add(q xyzz)  2: dbl(acc)  3: to_affine(acc)  4: mul_limbs(acc, k[8])
template <class F>
__global__ void k_dbg_pt(int op, const XYZZ<F>* acc, const uint32_t* q, int q_words, int neg, uint32_t* out, int out_words,
                         uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> a = acc[i];
  const uint32_t* qi = q + (size_t)i * q_words;
  uint32_t* oi = out + (size_t)i * out_words;
  if (op == 0) {
    pt_madd(a, *reinterpret_cast<const Affine<F>*>(qi), neg != 0);
    *reinterpret_cast<XYZZ<F>*>(oi) = a;
  } else if (op == 1) {
    pt_add(a, *reinterpret_cast<const XYZZ<F>*>(qi));
    *reinterpret_cast<XYZZ<F>*>(oi) = a;
  } else if (op == 2) {
    pt_dbl(a);
    *reinterpret_cast<XYZZ<F>*>(oi) = a;
  } else if (op == 3) {
    pt_to_affine(*reinterpret_cast<Affine<F>*>(oi), a);
  } else if (op == 4) {
    uint32_t k[8];
    for (int j = 0; j < 8; j++) k[j] = qi[j];
    *reinterpret_cast<XYZZ<F>*>(oi) = XYZZ<F>::mul_limbs(a, k, 8);
  }
}

template <class F>
static int run_fp(zkb_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  F *da, *db, *dout;
  ZKB_TRY(ws.alloc(&da, n));
  ZKB_TRY(ws.alloc(&db, n));
  ZKB_TRY(ws.alloc(&dout, n));
  ZKB_CUDA(ctx, cudaMemcpyAsync(da, a, n * sizeof(F), cudaMemcpyHostToDevice, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(db, b, n * sizeof(F), cudaMemcpyHostToDevice, st));
  ZKB_LAUNCH(ctx, (k_dbg_fp<F>), ceil_div(n, 64), 64, 0, st, op, da, db, dout, (uint32_t)n);
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, dout, n * sizeof(F), cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

template <class F>
static int run_pt(zkb_ctx* ctx, int op, const uint32_t* acc, const uint32_t* q, int neg, uint32_t* out, size_t n) {
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  const int pt_words = sizeof(XYZZ<F>) / 4, aff_words = sizeof(Affine<F>) / 4;
  const int q_words = op == 0 ? aff_words : op == 1 ? pt_words : op == 4 ? 8 : 1;
  const int out_words = op == 3 ? aff_words : pt_words;
  XYZZ<F>* dacc;
  uint32_t *dq, *dout;
  ZKB_TRY(ws.alloc(&dacc, n));
  ZKB_TRY(ws.alloc(&dq, n * q_words));
  ZKB_TRY(ws.alloc(&dout, n * out_words));
  ZKB_CUDA(ctx, cudaMemcpyAsync(dacc, acc, n * sizeof(XYZZ<F>), cudaMemcpyHostToDevice, st));
  if (op == 0 || op == 1 || op == 4)
    ZKB_CUDA(ctx, cudaMemcpyAsync(dq, q, n * q_words * 4, cudaMemcpyHostToDevice, st));
  ZKB_LAUNCH(ctx, (k_dbg_pt<F>), ceil_div(n, 64), 64, 0, st, op, dacc, dq, q_words, neg, dout, out_words, (uint32_t)n);
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, dout, n * out_words * 4, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

// field: 0 BnFr, 1 BlsFr, 2 BnFq, 3 BlsFq, 4 BnFq2, 5 BlsFq2; operands are n packed elements
int zkb_debug_fp_op(zkb_ctx* ctx, int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  if (!ctx || !a || !b || !out) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  switch (field) {
    case 0: return run_fp<Fp<BnFr>>(ctx, op, a, b, out, n);
    case 1: return run_fp<Fp<BlsFr>>(ctx, op, a, b, out, n);
    case 2: return run_fp<Fp<BnFq>>(ctx, op, a, b, out, n);
    case 3: return run_fp<Fp<BlsFq>>(ctx, op, a, b, out, n);
  }
  return set_err(ctx, ZKB_E_INVALID, "debug_fp_op: unknown field %d", field);
}

int zkb_debug_pt_op(zkb_ctx* ctx, int curve, int group, int op, const uint32_t* acc, const uint32_t* q, int neg,
                    uint32_t* out, size_t n) {
  if (!ctx || !acc || !out) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (curve == ZKB_BN254 && group == ZKB_G1) return run_pt<Fp<BnFq>>(ctx, op, acc, q, neg, out, n);
  if (curve == ZKB_BN254 && group == ZKB_G2) return run_pt<Fp2<BnFq>>(ctx, op, acc, q, neg, out, n);
  if (curve == ZKB_BLS12_381 && group == ZKB_G1) return run_pt<Fp<BlsFq>>(ctx, op, acc, q, neg, out, n);
  if (curve == ZKB_BLS12_381 && group == ZKB_G2) return run_pt<Fp2<BlsFq>>(ctx, op, acc, q, neg, out, n);
  return set_err(ctx, ZKB_E_INVALID, "debug_pt_op: unknown curve %d / group %d", curve, group);
}

}  // extern "C"
