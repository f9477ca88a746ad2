// C-ABI entry points for the Groth16 prove path (include/zkb.h); the per-curve device code
// is instantiated in groth16_bls.cu / groth16_bn.cu.
#include <vector>

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

// a one-shot call (host matrices deferred into the prove step) failed after staging: the deferred host pointers die
// with the caller's buffers, so the stage must not survive the call
static int unstage_on_error(zkb_ctx* ctx, int rc) {
  if (rc != ZKB_OK && ctx->stage) {
    ctx->stage->staged = false;
    ctx->stage->pending[0] = ctx->stage->pending[1] = ctx->stage->pending[2] = nullptr;
  }
  return rc;
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
  if (s->shard) cudaFree(s->shard);
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
  if (pk->q0_g1) cudaFree(pk->q0_g1);
  if (pk->q0_g2) cudaFree(pk->q0_g2);
  delete pk;
}

// contiguous, balanced partition of range(n): sizes differ by at most one, earlier ranks larger
// (the same rule as ckb_zkp_b200/parallel.py:shard_range)
static void shard_range(size_t n, int world, int rank, size_t* lo, size_t* hi) {
  size_t base = n / (size_t)world, extra = n % (size_t)world;
  size_t r = (size_t)rank;
  *lo = r * base + (r < extra ? r : extra);
  *hi = *lo + base + (r < extra ? 1 : 0);
}

int zkb_groth16_pk_create_sharded(zkb_ctx* ctx, int curve, const uint64_t* a_query, const uint8_t* a_inf, size_t a_len,
                                  const uint64_t* b_g1_query, const uint8_t* b_g1_inf, size_t b_g1_len,
                                  const uint64_t* b_g2_query, const uint8_t* b_g2_inf, size_t b_g2_len,
                                  const uint64_t* h_query, const uint8_t* h_inf, size_t h_len, const uint64_t* l_query,
                                  const uint8_t* l_inf, size_t l_len, const uint64_t* g1_singles, const uint64_t* g2_singles,
                                  int n_ranks, int rank, zkb_pk** out) {
  if (!ctx || !out) return ZKB_E_INVALID;
  *out = nullptr;
  const Groth16Ops* ops = groth16_ops(curve);
  if (!ops) return set_err(ctx, ZKB_E_INVALID, "pk_create: unknown curve %d", curve);
  if (!g1_singles || !g2_singles) return set_err(ctx, ZKB_E_INVALID, "pk_create: null vk elements");
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return set_err(ctx, ZKB_E_INVALID, "pk_create: bad rank %d of %d", rank, n_ranks);
  if (a_len == 0 || b_g1_len == 0 || b_g2_len == 0)
    return set_err(ctx, ZKB_E_INVALID, "groth16: empty query (index 0 is read by calculate_coeff)");
  zkb_pk* pk = new zkb_pk();
  pk->ctx = ctx; pk->curve = curve;
  pk->a = pk->b_g1 = pk->b_g2 = pk->h = pk->l = nullptr;
  pk->g1_singles = pk->g2_singles = nullptr;
  pk->sharded = true; pk->n_ranks = n_ranks; pk->rank = rank;
  const unsigned fl = ZKB_SRS_PRECOMPUTE;
  const size_t w1 = ops->g1_affine_bytes / 8, w2 = ops->g2_affine_bytes / 8;   // u64 words per affine point
  struct Q { const uint64_t* xy; const uint8_t* inf; size_t len; int group; size_t skip; zkb_srs** dst; };
  Q qs[5] = {{a_query, a_inf, a_len, ZKB_G1, 1, &pk->a},       {b_g1_query, b_g1_inf, b_g1_len, ZKB_G1, 1, &pk->b_g1},
             {b_g2_query, b_g2_inf, b_g2_len, ZKB_G2, 1, &pk->b_g2}, {h_query, h_inf, h_len, ZKB_G1, 0, &pk->h},
             {l_query, l_inf, l_len, ZKB_G1, 0, &pk->l}};
  int rc = ZKB_OK;
  for (int k = 0; k < 5 && rc == ZKB_OK; k++) {
    const Q& q = qs[k];
    size_t pairs = q.len - q.skip, lo, hi;
    shard_range(pairs, n_ranks, rank, &lo, &hi);
    pk->pair_lo[k] = lo;
    pk->pair_n[k] = hi - lo;
    const size_t w = q.group == ZKB_G1 ? w1 : w2;
    const uint64_t* xy = q.xy ? q.xy + (q.skip + lo) * w : nullptr;
    const uint8_t* inf = q.inf ? q.inf + q.skip + lo : nullptr;
    rc = zkb_srs_upload_shard(ctx, curve, q.group, xy, inf, hi - lo, lo, pairs, fl, q.dst);
  }
  if (rc == ZKB_OK) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t b1 = ops->g1_affine_bytes, b2 = ops->g2_affine_bytes;
    // index 0 of the coefficient queries; an identity there becomes the device's (0, 0) point
    std::vector<uint64_t> q0(2 * w1 + w2, 0);
    if (!a_inf[0]) memcpy(q0.data(), a_query, b1);
    if (!b_g1_inf[0]) memcpy(q0.data() + w1, b_g1_query, b1);
    if (!b_g2_inf[0]) memcpy(q0.data() + 2 * w1, b_g2_query, b2);
    cudaError_t e = cudaMalloc(&pk->g1_singles, 3 * b1);
    if (e == cudaSuccess) e = cudaMalloc(&pk->g2_singles, 2 * b2);
    if (e == cudaSuccess) e = cudaMalloc(&pk->q0_g1, 2 * b1);
    if (e == cudaSuccess) e = cudaMalloc(&pk->q0_g2, b2);
    if (e == cudaSuccess) e = cudaMemcpy(pk->g1_singles, g1_singles, 3 * b1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pk->g2_singles, g2_singles, 2 * b2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pk->q0_g1, q0.data(), 2 * b1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pk->q0_g2, q0.data() + 2 * w1, b2, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = set_err(ctx, ZKB_E_CUDA, "pk_create: %s", cudaGetErrorString(e));
  }
  if (rc != ZKB_OK) {
    zkb_groth16_pk_free(pk);
    return rc;
  }
  *out = pk;
  return ZKB_OK;
}

size_t zkb_groth16_partial_bytes(int curve) {
  const Groth16Ops* ops = groth16_ops(curve);
  return ops ? ops->partial_bytes : 0;
}

int zkb_groth16_prove_partial(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                              const uint64_t* z_mont, size_t n_inputs, size_t n_aux, const uint64_t r[4],
                              const uint64_t s[4], void* partial_out) {
  if (!ctx || !pk || !partial_out) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  const Groth16Ops* ops = groth16_ops(pk->curve);
  ZKB_TRY(ops->stage(ctx, pk, A, B, C, z_mont, n_inputs, n_aux, 1));
  ZKB_TRY(unstage_on_error(ctx, ops->prove_partial_staged(ctx, pk, r, s)));
  return ops->fetch_partial(ctx, partial_out);
}

int zkb_groth16_fold(zkb_ctx* ctx, const zkb_pk* pk, const void* partials, size_t count, const uint64_t r[4],
                     const uint64_t s[4], uint64_t* proof_xy, uint8_t* proof_inf) {
  if (!ctx || !pk || !partials || count == 0 || count > 4096 || !proof_xy || !proof_inf) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  const Groth16Ops* ops = groth16_ops(pk->curve);
  ZKB_TRY(ops->fold_partials(ctx, pk, partials, count, r, s, true));
  return ops->fetch_proof(ctx, pk, proof_xy, proof_inf);
}

int zkb_groth16_prove_sharded_staged(zkb_ctx* ctx, const zkb_pk* pk, const uint64_t r[4], const uint64_t s[4]) {
  if (!ctx || !pk) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  const Groth16Ops* ops = groth16_ops(pk->curve);
  ZKB_TRY(ops->prove_partial_staged(ctx, pk, r, s));
  return ops->fold_partials(ctx, pk, nullptr, 0, r, s, false);
}

int zkb_groth16_prove_sharded(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                              const uint64_t* z_mont, size_t n_inputs, size_t n_aux, const uint64_t r[4],
                              const uint64_t s[4], uint64_t* proof_xy, uint8_t* proof_inf) {
  if (!ctx || !pk || !proof_xy || !proof_inf) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  const Groth16Ops* ops = groth16_ops(pk->curve);
  ZKB_TRY(ops->stage(ctx, pk, A, B, C, z_mont, n_inputs, n_aux, 1));
  ZKB_TRY(unstage_on_error(ctx, ops->prove_partial_staged(ctx, pk, r, s)));
  ZKB_TRY(ops->fold_partials(ctx, pk, nullptr, 0, r, s, false));
  return ops->fetch_proof(ctx, pk, proof_xy, proof_inf);
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
  ZKB_TRY(unstage_on_error(ctx, ops->prove_staged(ctx, pk, r, s)));
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
