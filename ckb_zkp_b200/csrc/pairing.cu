// zkb_multi_pairing (include/zkb.h): products of ate pairings for batched verification -- see pairing.cuh.
#include "common.cuh"
#include "pairing.cuh"

namespace zkb {

template <class PP>
static int multi_pairing_t(zkb_ctx* ctx, const uint64_t* g1_xy, const uint8_t* g1_inf, const uint64_t* g2_xy,
                           const uint8_t* g2_inf, size_t n_groups, size_t group_size, uint64_t* out_gt) {
  using Fq = Fp<typename PP::FqP>;
  using F12 = typename PairingT<PP>::F12;
  static_assert(sizeof(F12) == 12 * sizeof(Fq), "Fq12 is twelve packed Fq");
  const size_t n = n_groups * group_size;
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  Affine<Fq>* d_g1;
  Affine<Fp2<typename PP::FqP>>* d_g2;
  uint8_t *d_i1 = nullptr, *d_i2 = nullptr;
  F12 *d_ml, *d_out;
  ZKB_TRY(ws.alloc(&d_g1, n));
  ZKB_TRY(ws.alloc(&d_g2, n));
  ZKB_TRY(ws.alloc(&d_ml, n));
  ZKB_TRY(ws.alloc(&d_out, n_groups));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_g1, g1_xy, n * sizeof(*d_g1), cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_g2, g2_xy, n * sizeof(*d_g2), cudaMemcpyDefault, st));
  if (g1_inf) {
    ZKB_TRY(ws.alloc(&d_i1, n));
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_i1, g1_inf, n, cudaMemcpyDefault, st));
  }
  if (g2_inf) {
    ZKB_TRY(ws.alloc(&d_i2, n));
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_i2, g2_inf, n, cudaMemcpyDefault, st));
  }
  // 64-thread blocks: a pairing thread lives in local memory (an Fq12 is up to 576 bytes), small blocks spread a
  // modest batch over all SMs
  ZKB_LAUNCH(ctx, (k_miller_loops<PP>), ceil_div(n, 64), 64, 0, st, d_g1, d_i1, d_g2, d_i2, n, d_ml);
  // a long product is folded by chunks of 16 first, so that no thread multiplies more than 16 + 16 values in a row
  const size_t kChunk = 16;
  size_t group = group_size;
  F12* d_cur = d_ml;
  while (group > 2 * kChunk) {
    const size_t n_chunks = ceil_div(group, kChunk);
    F12* d_next;
    ZKB_TRY(ws.alloc(&d_next, n_groups * n_chunks));
    ZKB_LAUNCH(ctx, (k_gt_chunk_products<PP>), ceil_div(n_groups * n_chunks, 64), 64, 0, st, d_cur, n_groups, group, kChunk,
               n_chunks, d_next);
    d_cur = d_next;
    group = n_chunks;
  }
  ZKB_LAUNCH(ctx, (k_pairing_finish<PP>), ceil_div(n_groups, 64), 64, 0, st, d_cur, n_groups, group, d_out);
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_gt, d_out, n_groups * sizeof(F12), cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

extern "C" int zkb_multi_pairing(zkb_ctx* ctx, int curve, const uint64_t* g1_xy_mont, const uint8_t* g1_inf,
                                 const uint64_t* g2_xy_mont, const uint8_t* g2_inf, size_t n_groups, size_t group_size,
                                 uint64_t* out_gt_mont) {
  if (!ctx) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "multi_pairing: unknown curve %d", curve);
  if (n_groups == 0) return ZKB_OK;
  if (group_size == 0) return set_err(ctx, ZKB_E_INVALID, "multi_pairing: group_size must be at least 1");
  if (!g1_xy_mont || !g2_xy_mont || !out_gt_mont) return set_err(ctx, ZKB_E_INVALID, "multi_pairing: null buffer");
  if (n_groups > (size_t(1) << 24) || group_size > (size_t(1) << 24) || n_groups * group_size > (size_t(1) << 24))
    return set_err(ctx, ZKB_E_TOO_LARGE, "multi_pairing: more than 2^24 pairs in one call");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return curve == ZKB_BN254
             ? multi_pairing_t<BnPairing>(ctx, g1_xy_mont, g1_inf, g2_xy_mont, g2_inf, n_groups, group_size, out_gt_mont)
             : multi_pairing_t<BlsPairing>(ctx, g1_xy_mont, g1_inf, g2_xy_mont, g2_inf, n_groups, group_size, out_gt_mont);
}
