// C-ABI entry points for the Groth16 prove path (include/zkb.h); the per-curve device code
// is instantiated in groth16_bls.cu / groth16_bn.cu.
#include "groth16.cuh"

namespace zkb {

const Groth16Ops* groth16_ops_bls();
const Groth16Ops* groth16_ops_bn();

const Groth16Ops* groth16_ops(int curve) {
  if (curve == ZKB_BLS12_381) return groth16_ops_bls();
  if (curve == ZKB_BN254) return groth16_ops_bn();
  return nullptr;
}

static void free_buf(DevBuf* b) {
  if (b->p) cudaFree(b->p);
  b->p = nullptr;
  b->cap = 0;
}

void groth16_free_stage(zkb_ctx* ctx) {
  Groth16Stage* s = ctx->stage;
  if (!s) return;
  for (DevCsr* m : {&s->A, &s->B, &s->C}) {
    free_buf(&m->row_ptr);
    free_buf(&m->col_idx);
    free_buf(&m->coeff);
  }
  for (DevBuf* b : {&s->z, &s->z_repr, &s->va, &s->vb, &s->vc, &s->scratch}) free_buf(b);
  if (s->results) cudaFree(s->results);
  if (s->scal) cudaFree(s->scal);
  delete s;
  ctx->stage = nullptr;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_groth16_pk_create(zkb_ctx* ctx, int curve, const uint64_t* a_query, const uint8_t* a_inf, size_t a_len,
                          const uint64_t* b_g1_query, const uint8_t* b_g1_inf, size_t b_g1_len,
                          const uint64_t* b_g2_query, const uint8_t* b_g2_inf, size_t b_g2_len,
                          const uint64_t* h_query, const uint8_t* h_inf, size_t h_len, const uint64_t* l_query,
                          const uint8_t* l_inf, size_t l_len, const uint64_t* g1_singles, const uint64_t* g2_singles,
                          zkb_pk** out) {
  if (!ctx || !out) return ZKB_E_INVALID;
  *out = nullptr;
  const Groth16Ops* ops = groth16_ops(curve);
  if (!ops) return set_err(ctx, ZKB_E_INVALID, "pk_create: unknown curve %d", curve);
  if (!g1_singles || !g2_singles) return set_err(ctx, ZKB_E_INVALID, "pk_create: null vk elements");
  zkb_pk* pk = new zkb_pk();
  pk->ctx = ctx; pk->curve = curve;
  pk->a = pk->b_g1 = pk->b_g2 = pk->h = pk->l = nullptr;
  pk->g1_singles = pk->g2_singles = nullptr;
  const unsigned fl = ZKB_SRS_PRECOMPUTE;
  int rc = zkb_srs_upload(ctx, curve, ZKB_G1, a_query, a_inf, a_len, fl, &pk->a);
  if (rc == ZKB_OK) rc = zkb_srs_upload(ctx, curve, ZKB_G1, b_g1_query, b_g1_inf, b_g1_len, fl, &pk->b_g1);
  if (rc == ZKB_OK) rc = zkb_srs_upload(ctx, curve, ZKB_G2, b_g2_query, b_g2_inf, b_g2_len, fl, &pk->b_g2);
  if (rc == ZKB_OK) rc = zkb_srs_upload(ctx, curve, ZKB_G1, h_query, h_inf, h_len, fl, &pk->h);
  if (rc == ZKB_OK) rc = zkb_srs_upload(ctx, curve, ZKB_G1, l_query, l_inf, l_len, fl, &pk->l);
  if (rc == ZKB_OK) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaError_t e = cudaMalloc(&pk->g1_singles, 3 * ops->g1_affine_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&pk->g2_singles, 2 * ops->g2_affine_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(pk->g1_singles, g1_singles, 3 * ops->g1_affine_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pk->g2_singles, g2_singles, 2 * ops->g2_affine_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = set_err(ctx, ZKB_E_CUDA, "pk_create: %s", cudaGetErrorString(e));
  }
  if (rc != ZKB_OK) {
    zkb_groth16_pk_free(pk);
    return rc;
  }
  *out = pk;
  return ZKB_OK;
}

void zkb_groth16_pk_free(zkb_pk* pk) {
  if (!pk) return;
  zkb_srs_free(pk->a);
  zkb_srs_free(pk->b_g1);
  zkb_srs_free(pk->b_g2);
  zkb_srs_free(pk->h);
  zkb_srs_free(pk->l);
  if (pk->g1_singles) cudaFree(pk->g1_singles);
  if (pk->g2_singles) cudaFree(pk->g2_singles);
  delete pk;
}

int zkb_groth16_stage(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                      const uint64_t* z_mont, size_t n_inputs, size_t n_aux) {
  if (!ctx || !pk) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return groth16_ops(pk->curve)->stage(ctx, pk, A, B, C, z_mont, n_inputs, n_aux, 0);
}

int zkb_groth16_prove_staged(zkb_ctx* ctx, const zkb_pk* pk, const uint64_t r[4], const uint64_t s[4]) {
  if (!ctx || !pk) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return groth16_ops(pk->curve)->prove_staged(ctx, pk, r, s);
}

int zkb_groth16_fetch_proof(zkb_ctx* ctx, const zkb_pk* pk, uint64_t* proof_xy, uint8_t* proof_inf) {
  if (!ctx || !pk || !proof_xy || !proof_inf) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return groth16_ops(pk->curve)->fetch_proof(ctx, pk, proof_xy, proof_inf);
}

int zkb_groth16_prove(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                      const uint64_t* z_mont, size_t n_inputs, size_t n_aux, const uint64_t r[4], const uint64_t s[4],
                      uint64_t* proof_xy, uint8_t* proof_inf) {
  if (!ctx || !pk || !proof_xy || !proof_inf) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  const Groth16Ops* ops = groth16_ops(pk->curve);
  ZKB_TRY(ops->stage(ctx, pk, A, B, C, z_mont, n_inputs, n_aux, 1));
  ZKB_TRY(ops->prove_staged(ctx, pk, r, s));
  return ops->fetch_proof(ctx, pk, proof_xy, proof_inf);
}

int zkb_groth16_h(zkb_ctx* ctx, int curve, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C, const uint64_t* z_mont,
                  size_t n_inputs, size_t n_aux, uint64_t* h_canonical) {
  if (!ctx || !h_canonical) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const Groth16Ops* ops = groth16_ops(curve);
  if (!ops) return set_err(ctx, ZKB_E_INVALID, "groth16_h: unknown curve %d", curve);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  ZKB_TRY(ops->stage(ctx, nullptr, A, B, C, z_mont, n_inputs, n_aux, 0));
  ZKB_TRY(ops->compute_h(ctx, ctx->main));
  return ops->fetch_h(ctx, h_canonical);
}

}  // extern "C"
