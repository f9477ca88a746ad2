// Groth16 prove path on the device, templated on the curve.  Included by one .cu per curve.
//
// Follows groth16/src/prover.rs:148-210 and groth16/src/r1cs_to_qap.rs:113-172 step by step;
// what changes is where each step runs (all of it on the GPU) and that the 1/N of the three
// inverse transforms and the coset shift of the following forward transforms are applied by
// the NTT passes themselves (ntt.cu).
#pragma once
#include "groth16.cuh"
#include "devutil.cuh"

namespace zkb {

int fr_convert_dev(zkb_ctx* ctx, cudaStream_t st, int curve, const void* d_in, void* d_out, size_t n, int mode);

// ------------------------------------------------------------------------------------------
// evaluate_constraint over all rows (r1cs_to_qap.rs:15-52,131-142,155-159): out[i] = <row_i, z>
// for i < n_rows; for the A matrix rows n_rows .. n_rows + n_inputs - 1 carry z_input[i]
// (:140-142); everything up to the domain size is zero.  One thread per row: rows hold 1-3
// terms, adjacent rows are adjacent in memory.
// ------------------------------------------------------------------------------------------
template <class FrP>
__global__ void __launch_bounds__(256)
k_spmv(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col_idx, const Fp<FrP>* __restrict__ coeff,
       const Fp<FrP>* __restrict__ z, Fp<FrP>* __restrict__ out, uint32_t n_rows, uint32_t n_inputs_tail,
       uint32_t domain) {
  using Fr = Fp<FrP>;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= domain) return;
  Fr acc = Fr::zero();
  if (i < n_rows) {
    uint32_t p0 = row_ptr[i], p1 = row_ptr[i + 1];
    const Fr one = Fr::one();
    for (uint32_t p = p0; p < p1; p++) {
      Fr c = ld_vec(&coeff[p]);
      Fr v = ld_vec(&z[col_idx[p]]);
      if (c != one) v = Fr::mul(v, c);      // the reference's is_one fast path (r1cs_to_qap.rs:41-45)
      acc = Fr::add(acc, v);
    }
  } else if (i < n_rows + n_inputs_tail) {
    acc = ld_vec(&z[i - n_rows]);
  }
  st_vec(&out[i], acc);
}

// ab = (a * b - c) * 1/Z(g)   (r1cs_to_qap.rs:150,164-168)
template <class FrP>
__global__ void __launch_bounds__(256)
k_qap_pointwise(Fp<FrP>* __restrict__ a, const Fp<FrP>* __restrict__ b, const Fp<FrP>* __restrict__ c,
                const Fp<FrP>* __restrict__ zinv_p, size_t n) {
  using Fr = Fp<FrP>;
  const Fr zinv = *zinv_p;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr x = Fr::mul(ld_vec_rw(&a[i]), ld_vec_rw(&b[i]));
    x = Fr::sub(x, ld_vec_rw(&c[i]));
    st_vec(&a[i], Fr::mul(x, zinv));
  }
}

// ------------------------------------------------------------------------------------------
// proof assembly (prover.rs:164-210)
// ------------------------------------------------------------------------------------------
template <class Fq, class Fq2>
struct G16Results {
  // MSM outputs
  XYZZ<Fq> msm_a, msm_b1, msm_h, msm_l;
  XYZZ<Fq2> msm_b2;
  // r*delta, s*delta, (r*s)*delta in G1; s*delta in G2
  XYZZ<Fq> r_delta, s_delta, rs_delta;
  XYZZ<Fq2> s_delta2;
  // g_a, g1_b and s*g_a, r*g1_b
  XYZZ<Fq> g_a, g1_b, s_g_a, r_g1_b;
  XYZZ<Fq2> g2_b;
  // proof, canonical affine
  Affine<Fq> proof_a;
  Affine<Fq2> proof_b;
  Affine<Fq> proof_c;
  uint32_t inf[4];
};

// scal: [r, s] canonical -> scal[2] = r * s mod p (canonical); four independent scalar
// multiplications, one warp-lane-0 each in separate blocks
template <class FrP, class Fq, class Fq2>
__global__ void k_g16_scalars(Fp<FrP> r, Fp<FrP> s, Fp<FrP>* scal, const Affine<Fq>* g1_singles,
                              const Affine<Fq2>* g2_singles, G16Results<Fq, Fq2>* res) {
  using Fr = Fp<FrP>;
  if (threadIdx.x) return;
  if (blockIdx.x == 0) scal[0] = r;
  if (blockIdx.x == 1) scal[1] = s;
  XYZZ<Fq> d1 = XYZZ<Fq>::from_affine(g1_singles[2]);
  if (blockIdx.x == 0) st_vec(&res->r_delta, XYZZ<Fq>::mul_limbs(d1, r.v, Fr::N));
  if (blockIdx.x == 1) st_vec(&res->s_delta, XYZZ<Fq>::mul_limbs(d1, s.v, Fr::N));
  if (blockIdx.x == 2) {
    Fr rs = Fr::from_mont(Fr::mul(Fr::to_mont(r), Fr::to_mont(s)));
    scal[2] = rs;
    st_vec(&res->rs_delta, XYZZ<Fq>::mul_limbs(d1, rs.v, Fr::N));
  }
  if (blockIdx.x == 3) {
    XYZZ<Fq2> d2 = XYZZ<Fq2>::from_affine(g2_singles[1]);
    st_vec(&res->s_delta2, XYZZ<Fq2>::mul_limbs(d2, s.v, Fr::N));
  }
}

// calculate_coeff (prover.rs:213-228): res = initial + query[0] + acc + vk_param, for A, B-G1, B-G2;
// then s*g_a and r*g1_b (prover.rs:192-193)
template <class FrP, class Fq, class Fq2>
__global__ void k_g16_coeffs(const Fp<FrP>* scal, const Affine<Fq>* a0, const Affine<Fq>* b1_0, const Affine<Fq2>* b2_0,
                             const Affine<Fq>* g1_singles, const Affine<Fq2>* g2_singles, G16Results<Fq, Fq2>* res) {
  using Fr = Fp<FrP>;
  if (threadIdx.x) return;
  Fr r = scal[0], s = scal[1];
  if (blockIdx.x == 0) {
    XYZZ<Fq> g = ld_vec_rw(&res->r_delta);
    pt_madd(g, *a0, false);
    { XYZZ<Fq> m = ld_vec_rw(&res->msm_a); pt_add(g, m); }
    pt_madd(g, g1_singles[0], false);    // alpha_g1
    st_vec(&res->g_a, g);
    st_vec(&res->s_g_a, XYZZ<Fq>::mul_limbs(g, s.v, Fr::N));
  }
  if (blockIdx.x == 1) {
    XYZZ<Fq> g = XYZZ<Fq>::inf();
    if (!r.is_zero()) {                  // the guard is on r (prover.rs:170)
      g = ld_vec_rw(&res->s_delta);
      pt_madd(g, *b1_0, false);
      { XYZZ<Fq> m = ld_vec_rw(&res->msm_b1); pt_add(g, m); }
      pt_madd(g, g1_singles[1], false);  // beta_g1
    }
    st_vec(&res->g1_b, g);
    st_vec(&res->r_g1_b, XYZZ<Fq>::mul_limbs(g, r.v, Fr::N));
  }
  if (blockIdx.x == 2) {
    XYZZ<Fq2> g = ld_vec_rw(&res->s_delta2);
    pt_madd(g, *b2_0, false);
    { XYZZ<Fq2> m = ld_vec_rw(&res->msm_b2); pt_add(g, m); }
    pt_madd(g, g2_singles[0], false);    // beta_g2
    st_vec(&res->g2_b, g);
  }
}

// g_c = s*g_a + r*g1_b - r*s*delta + l_acc + h_acc; into_affine of (g_a, g2_b, g_c)  (prover.rs:195-210)
// `first` selects the blocks: proof_a / proof_b (blocks 0, 1) only need the coefficients and are converted
// off the critical path; proof_c (block 2) waits for every MSM.
template <class Fq, class Fq2>
__global__ void k_g16_finish(G16Results<Fq, Fq2>* res, unsigned first) {
  if (threadIdx.x) return;
  const unsigned blk = blockIdx.x + first;
  if (blk == 0) {
    XYZZ<Fq> g = ld_vec_rw(&res->g_a);
    Affine<Fq> a;
    pt_to_affine(a, g);
    st_vec(&res->proof_a, a);
    res->inf[0] = g.is_inf();
  }
  if (blk == 1) {
    XYZZ<Fq2> g = ld_vec_rw(&res->g2_b);
    Affine<Fq2> a;
    pt_to_affine(a, g);
    st_vec(&res->proof_b, a);
    res->inf[1] = g.is_inf();
  }
  if (blk == 2) {
    XYZZ<Fq> g = ld_vec_rw(&res->s_g_a);
    XYZZ<Fq> t = ld_vec_rw(&res->r_g1_b);
    pt_add(g, t);
    t = ld_vec_rw(&res->rs_delta);
    t.neg_in_place();
    pt_add(g, t);
    t = ld_vec_rw(&res->msm_l);
    pt_add(g, t);
    t = ld_vec_rw(&res->msm_h);
    pt_add(g, t);
    Affine<Fq> a;
    pt_to_affine(a, g);
    st_vec(&res->proof_c, a);
    res->inf[2] = g.is_inf();
  }
}

// ------------------------------------------------------------------------------------------
// sharded proof (one process per GPU): every rank owns a slice of the pairs of each of the five MSMs.
// With A_k, B1_k, B2_k, L_k, H_k the partial sums of rank k, the proof of prover.rs:164-210 is
//   g_a  = Fa  + sum_k A_k                       Fa  = r*delta + a_query[0] + alpha_g1
//   g2_b = Fb2 + sum_k B2_k                      Fb2 = s*delta_g2 + b_g2_query[0] + beta_g2
//   g_c  = s*Fa + r*Fb1 - r*s*delta + sum_k C_k  Fb1 = s*delta + b_g1_query[0] + beta_g1
//   C_k  = s*A_k + r*B1_k + L_k + H_k
// (the r == 0 guard of prover.rs:170 only zeroes r*g1_b, which r = 0 does anyway).  The scalar multiplications
// by r and s are applied to the rank's own partials while its H / L accumulations still run, so after the ONE
// all-gather of (A_k, C_k, B2_k) only <= n_ranks additions and one inversion per proof element remain.
// ------------------------------------------------------------------------------------------
template <class Fq, class Fq2>
struct G16Partial {         // what one rank contributes to the all-gather
  XYZZ<Fq> a, c;
  XYZZ<Fq2> b2;
};
template <class Fq, class Fq2>
struct G16Shard {
  G16Partial<Fq, Fq2> part;
  XYZZ<Fq> msm_b1, msm_h, msm_l;   // local partial sums that only feed part.c
  XYZZ<Fq> s_a, r_b1;              // s * A_k, r * B1_k
  XYZZ<Fq> fa, s_fa, r_fb1;        // rank-independent terms
  XYZZ<Fq2> fb2;
};

// after k_g16_scalars: block 0: fa, s*fa   block 1: r*fb1   block 2: fb2
template <class FrP, class Fq, class Fq2>
__global__ void k_g16_fixed(const Fp<FrP>* scal, const Affine<Fq>* q0_g1, const Affine<Fq2>* q0_g2,
                            const Affine<Fq>* g1_singles, const Affine<Fq2>* g2_singles,
                            const G16Results<Fq, Fq2>* res, G16Shard<Fq, Fq2>* sh) {
  using Fr = Fp<FrP>;
  if (threadIdx.x) return;
  Fr r = scal[0], s = scal[1];
  if (blockIdx.x == 0) {
    XYZZ<Fq> g = ld_vec_rw(&res->r_delta);
    pt_madd(g, q0_g1[0], false);
    pt_madd(g, g1_singles[0], false);    // alpha_g1
    st_vec(&sh->fa, g);
    st_vec(&sh->s_fa, XYZZ<Fq>::mul_limbs(g, s.v, Fr::N));
  }
  if (blockIdx.x == 1) {
    XYZZ<Fq> g = ld_vec_rw(&res->s_delta);
    pt_madd(g, q0_g1[1], false);
    pt_madd(g, g1_singles[1], false);    // beta_g1
    st_vec(&sh->r_fb1, XYZZ<Fq>::mul_limbs(g, r.v, Fr::N));
  }
  if (blockIdx.x == 2) {
    XYZZ<Fq2> g = ld_vec_rw(&res->s_delta2);
    pt_madd(g, q0_g2[0], false);
    pt_madd(g, g2_singles[0], false);    // beta_g2
    st_vec(&sh->fb2, g);
  }
}
// block 0: s * A_k   block 1: r * B1_k
template <class FrP, class Fq, class Fq2>
__global__ void k_g16_local_mul(const Fp<FrP>* scal, G16Shard<Fq, Fq2>* sh) {
  using Fr = Fp<FrP>;
  if (threadIdx.x) return;
  Fr r = scal[0], s = scal[1];
  if (blockIdx.x == 0) { XYZZ<Fq> p = ld_vec_rw(&sh->part.a); st_vec(&sh->s_a, XYZZ<Fq>::mul_limbs(p, s.v, Fr::N)); }
  if (blockIdx.x == 1) { XYZZ<Fq> p = ld_vec_rw(&sh->msm_b1); st_vec(&sh->r_b1, XYZZ<Fq>::mul_limbs(p, r.v, Fr::N)); }
}
// C_k = s*A_k + r*B1_k + L_k + H_k
template <class Fq, class Fq2>
__global__ void k_g16_local_c(G16Shard<Fq, Fq2>* sh) {
  if (threadIdx.x | blockIdx.x) return;
  XYZZ<Fq> g = ld_vec_rw(&sh->s_a);
  XYZZ<Fq> t = ld_vec_rw(&sh->r_b1);
  pt_add(g, t);
  t = ld_vec_rw(&sh->msm_l);
  pt_add(g, t);
  t = ld_vec_rw(&sh->msm_h);
  pt_add(g, t);
  st_vec(&sh->part.c, g);
}
// fold the gathered partials in rank order, add the rank-independent terms, into_affine  (prover.rs:206-210)
template <class Fq, class Fq2>
__global__ void k_g16_fold_finish(const G16Partial<Fq, Fq2>* __restrict__ parts, uint32_t count,
                                  const G16Shard<Fq, Fq2>* __restrict__ sh, G16Results<Fq, Fq2>* res) {
  if (threadIdx.x) return;
  if (blockIdx.x == 0) {
    XYZZ<Fq> g = ld_vec_rw(&sh->fa);
    for (uint32_t k = 0; k < count; k++) { XYZZ<Fq> t = ld_vec_rw(&parts[k].a); pt_add(g, t); }
    Affine<Fq> a;
    pt_to_affine(a, g);
    st_vec(&res->proof_a, a);
    res->inf[0] = g.is_inf();
  }
  if (blockIdx.x == 1) {
    XYZZ<Fq2> g = ld_vec_rw(&sh->fb2);
    for (uint32_t k = 0; k < count; k++) { XYZZ<Fq2> t = ld_vec_rw(&parts[k].b2); pt_add(g, t); }
    Affine<Fq2> a;
    pt_to_affine(a, g);
    st_vec(&res->proof_b, a);
    res->inf[1] = g.is_inf();
  }
  if (blockIdx.x == 2) {
    XYZZ<Fq> g = ld_vec_rw(&sh->s_fa);
    XYZZ<Fq> t = ld_vec_rw(&sh->r_fb1);
    pt_add(g, t);
    t = ld_vec_rw(&res->rs_delta);
    t.neg_in_place();
    pt_add(g, t);
    for (uint32_t k = 0; k < count; k++) { t = ld_vec_rw(&parts[k].c); pt_add(g, t); }
    Affine<Fq> a;
    pt_to_affine(a, g);
    st_vec(&res->proof_c, a);
    res->inf[2] = g.is_inf();
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <class FrP, class FqP, int CURVE>
struct Groth16Impl {
  using Fr = Fp<FrP>;
  using Fq = Fp<FqP>;
  using Fq2 = Fp2<FqP>;
  using Res = G16Results<Fq, Fq2>;
  using Part = G16Partial<Fq, Fq2>;
  using Shard = G16Shard<Fq, Fq2>;
  static_assert(sizeof(Res) <= kStageBlockBytes && sizeof(Shard) <= kStageBlockBytes, "stage blocks hold either curve's layout");

  static int ensure(zkb_ctx* ctx, DevBuf* b, size_t bytes) {
    if (b->p && b->cap >= bytes) return ZKB_OK;
    if (b->p) ZKB_CUDA(ctx, cudaFree(b->p));
    b->p = nullptr;
    b->cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    ZKB_CUDA(ctx, cudaMalloc(&b->p, want));
    b->cap = want;
    return ZKB_OK;
  }

  static int upload_csr(zkb_ctx* ctx, cudaStream_t st, DevCsr* d, const zkb_csr* h) {
    if ((h->n_rows && !h->row_ptr) || (h->nnz && (!h->col_idx || !h->coeff_mont)))
      return set_err(ctx, ZKB_E_INVALID, "groth16: null matrix");
    if (h->n_rows >= (size_t(1) << 31) || h->nnz >= (size_t(1) << 32))
      return set_err(ctx, ZKB_E_INVALID, "groth16: matrix too large");
    ZKB_TRY(ensure(ctx, &d->row_ptr, (h->n_rows + 1) * 4));
    ZKB_TRY(ensure(ctx, &d->col_idx, h->nnz * 4));
    ZKB_TRY(ensure(ctx, &d->coeff, h->nnz * sizeof(Fr)));
    d->n_rows = h->n_rows;
    d->nnz = h->nnz;
    if (h->n_rows)
      ZKB_CUDA(ctx, cudaMemcpyAsync(d->row_ptr.p, h->row_ptr, (h->n_rows + 1) * 4, cudaMemcpyHostToDevice, st));
    else
      ZKB_CUDA(ctx, cudaMemsetAsync(d->row_ptr.p, 0, 4, st));
    if (h->nnz) {
      ZKB_CUDA(ctx, cudaMemcpyAsync(d->col_idx.p, h->col_idx, h->nnz * 4, cudaMemcpyHostToDevice, st));
      ZKB_CUDA(ctx, cudaMemcpyAsync(d->coeff.p, h->coeff_mont, h->nnz * sizeof(Fr), cudaMemcpyHostToDevice, st));
    }
    return ZKB_OK;
  }

  // defer_matrices: copy only the assignment now; prove_staged uploads A, B, C after it has started the
  // four assignment MSMs on their side streams, so the 200 MB of matrix traffic overlaps their compute
  static int stage(zkb_ctx* ctx, const zkb_pk*, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                   const uint64_t* z_mont, size_t n_inputs, size_t n_aux, int defer_matrices) {
    if (!ctx->stage) ctx->stage = new Groth16Stage();
    Groth16Stage* s = ctx->stage;
    s->staged = false;
    if (!A || !B || !C || !z_mont) return set_err(ctx, ZKB_E_INVALID, "groth16: null argument");
    if (A->n_rows != B->n_rows || A->n_rows != C->n_rows)
      return set_err(ctx, ZKB_E_INVALID, "groth16: A, B, C row counts differ");
    if (n_inputs == 0) return set_err(ctx, ZKB_E_INVALID, "groth16: input 0 (ONE) missing");
    // domain = EvaluationDomain::new(num_constraints + num_inputs)  (r1cs_to_qap.rs:123-126)
    size_t need = A->n_rows + n_inputs;
    unsigned log_n = ceil_log2(need);
    if ((int)log_n > FrP::TWO_ADICITY || log_n > 30)
      return set_err(ctx, ZKB_E_TOO_LARGE, "groth16: domain 2^%u exceeds the field's 2-adicity", log_n);
    cudaStream_t st = ctx->main;
    s->curve = CURVE;
    s->n_inputs = n_inputs; s->n_aux = n_aux; s->n_rows = A->n_rows;
    s->log_n = log_n; s->N = size_t(1) << log_n;
    size_t nz = n_inputs + n_aux;
    ZKB_TRY(ensure(ctx, &s->z, nz * sizeof(Fr)));
    ZKB_TRY(ensure(ctx, &s->z_repr, nz * sizeof(Fr)));
    ZKB_TRY(ensure(ctx, &s->va, s->N * sizeof(Fr)));
    ZKB_TRY(ensure(ctx, &s->vb, s->N * sizeof(Fr)));
    ZKB_TRY(ensure(ctx, &s->vc, s->N * sizeof(Fr)));
    ZKB_TRY(ensure(ctx, &s->scratch, s->N * sizeof(Fr)));
    if (!s->results) ZKB_CUDA(ctx, cudaMalloc(&s->results, kStageBlockBytes));
    if (!s->scal) ZKB_CUDA(ctx, cudaMalloc(&s->scal, sizeof(Fr) * 4));
    ZKB_CUDA(ctx, cudaMemcpyAsync(s->z.p, z_mont, nz * sizeof(Fr), cudaMemcpyHostToDevice, st));
    s->pending[0] = s->pending[1] = s->pending[2] = nullptr;
    if (defer_matrices) {
      for (const zkb_csr* m : {A, B, C})
        if ((m->n_rows && !m->row_ptr) || (m->nnz && (!m->col_idx || !m->coeff_mont)))
          return set_err(ctx, ZKB_E_INVALID, "groth16: null matrix");
      s->pending[0] = A; s->pending[1] = B; s->pending[2] = C;
    } else {
      ZKB_TRY(upload_csr(ctx, st, &s->A, A));
      ZKB_TRY(upload_csr(ctx, st, &s->B, B));
      ZKB_TRY(upload_csr(ctx, st, &s->C, C));
    }
    s->staged = true;
    return ZKB_OK;
  }

  // witness_map + into_repr: h (canonical) ends up in stage->va
  static int compute_h(zkb_ctx* ctx, cudaStream_t st) {
    Groth16Stage* s = ctx->stage;
    if (!s || !s->staged || s->curve != CURVE) return set_err(ctx, ZKB_E_INVALID, "groth16: nothing staged");
    NttDomain* dom;
    ZKB_TRY(ntt_get_domain(ctx, CURVE, s->log_n, &dom));
    const uint32_t N = (uint32_t)s->N;
    Fr *a = (Fr*)s->va.p, *b = (Fr*)s->vb.p, *c = (Fr*)s->vc.p;
    const unsigned blocks = ceil_div(N, 256);
    ZKB_LAUNCH(ctx, (k_spmv<FrP>), blocks, 256, 0, st, (const uint32_t*)s->A.row_ptr.p, (const uint32_t*)s->A.col_idx.p,
               (const Fr*)s->A.coeff.p, (const Fr*)s->z.p, a, (uint32_t)s->n_rows, (uint32_t)s->n_inputs, N);
    ZKB_LAUNCH(ctx, (k_spmv<FrP>), blocks, 256, 0, st, (const uint32_t*)s->B.row_ptr.p, (const uint32_t*)s->B.col_idx.p,
               (const Fr*)s->B.coeff.p, (const Fr*)s->z.p, b, (uint32_t)s->n_rows, 0u, N);
    ZKB_LAUNCH(ctx, (k_spmv<FrP>), blocks, 256, 0, st, (const uint32_t*)s->C.row_ptr.p, (const uint32_t*)s->C.col_idx.p,
               (const Fr*)s->C.coeff.p, (const Fr*)s->z.p, c, (uint32_t)s->n_rows, 0u, N);
    Fr* vecs[3] = {a, b, c};
    for (int v = 0; v < 3; v++) {
      ZKB_TRY(ntt_run(ctx, st, dom, vecs[v], s->scratch.p, ZKB_NTT_INVERSE));   // ifft_in_place       (:144-145,161)
      ZKB_TRY(ntt_run(ctx, st, dom, vecs[v], s->scratch.p, ZKB_NTT_COSET));     // coset_fft_in_place  (:147-148,162)
    }
    unsigned pw_blocks = blocks > (unsigned)ctx->sm_count * 8 ? ctx->sm_count * 8 : blocks;
    ZKB_LAUNCH(ctx, (k_qap_pointwise<FrP>), pw_blocks, 256, 0, st, a, (const Fr*)b, (const Fr*)c,
               (const Fr*)dom->consts + kConstZInv, (size_t)N);
    ZKB_TRY(ntt_run(ctx, st, dom, a, s->scratch.p, ZKB_NTT_INVERSE | ZKB_NTT_COSET));   // coset_ifft_in_place (:169)
    ZKB_TRY(fr_convert_dev(ctx, st, CURVE, a, a, N, 0));                                // into_repr (prover.rs:161)
    return ZKB_OK;
  }

  static int fetch_h(zkb_ctx* ctx, uint64_t* h) {
    Groth16Stage* s = ctx->stage;
    ZKB_CUDA(ctx, cudaMemcpyAsync(h, s->va.p, s->N * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->main));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return ZKB_OK;
  }

  static int prove_staged(zkb_ctx* ctx, const zkb_pk* pk, const uint64_t* r, const uint64_t* sc) {
    Groth16Stage* s = ctx->stage;
    if (!s || !s->staged || s->curve != CURVE) return set_err(ctx, ZKB_E_INVALID, "groth16: nothing staged");
    if (!pk || pk->curve != CURVE || pk->ctx != ctx) return set_err(ctx, ZKB_E_INVALID, "groth16: bad proving key");
    if (!r || !sc) return set_err(ctx, ZKB_E_INVALID, "groth16: null r/s");
    cudaStream_t st = ctx->main;
    // serial mode (zkb_set_serial): every kernel on the main stream, for per-kernel timing without overlap
    cudaStream_t side[kNumSideStreams];
    for (int i = 0; i < kNumSideStreams; i++) side[i] = ctx->serial ? st : ctx->side[i];
    const GroupOps* g1 = group_ops(CURVE, ZKB_G1);
    const GroupOps* g2 = group_ops(CURVE, ZKB_G2);
    Res* res = (Res*)s->results;
    Fr* scal = (Fr*)s->scal;
    if (pk->a->n == 0 || pk->b_g1->n == 0 || pk->b_g2->n == 0)
      return set_err(ctx, ZKB_E_INVALID, "groth16: empty query (index 0 is read by calculate_coeff)");
    Fr r_val, s_val;      // by-value kernel arguments: no staging buffer to race on
    memcpy(r_val.v, r, 32);
    memcpy(s_val.v, sc, 32);
    // the four delta multiples are independent of everything else: side stream, joined before the assembly
    ZKB_TRY(fork_streams(ctx, 1));
    ZKB_LAUNCH(ctx, (k_g16_scalars<FrP, Fq, Fq2>), 4, 32, 0, side[0], r_val, s_val, scal, (const Affine<Fq>*)pk->g1_singles,
               (const Affine<Fq2>*)pk->g2_singles, res);
    // assignment = into_repr(input[1..] ++ aux)  (prover.rs:150-158)
    const size_t n_assign = s->n_inputs - 1 + s->n_aux;
    ZKB_TRY(fr_convert_dev(ctx, st, CURVE, (const Fr*)s->z.p + 1, s->z_repr.p, n_assign, 0));
    const uint32_t* zr = (const uint32_t*)s->z_repr.p;
    auto clamp = [](size_t n, const zkb_srs* srs, size_t off) { size_t a = srs->n > off ? srs->n - off : 0; return n < a ? n : a; };
    // The four MSMs over the assignment are independent of each other and of witness_map: one
    // high-priority side stream each for their sorts and bucket reductions, while the accumulation
    // kernels of all five MSMs queue on the low-priority bulk stream in the order of the calls below
    // (b_g2, a, b_g1, h, l).  a, b_g1 and b_g2 go first because calculate_coeff's scalar multiplications
    // (s * g_a, r * g1_b) hang off them: they run on side[5] under the H and L accumulations, so that
    // after the last bucket reduction only the five additions and the inversion of proof.c remain.
    ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    for (int i = 1; i <= 5; i++) ZKB_CUDA(ctx, cudaStreamWaitEvent(side[i], ctx->ev_fork, 0));
    // sorts of all four assignment MSMs first, their accumulations after: a sort launched behind another MSM's
    // accumulation kernel would wait for its thousands of pending blocks (msm.cuh run_split)
    std::function<int()> acc_b2, acc_a, acc_b1, acc_l;
    static const int sorts_first = []() { const char* e = getenv("ZKB_SORTS_FIRST"); return e ? atoi(e) : 1; }();
    ZKB_TRY(g2->msm_run_split(ctx, side[1], pk->b_g2, 1, zr, clamp(n_assign, pk->b_g2, 1), 0, &res->msm_b2, sorts_first ? &acc_b2 : nullptr));
    ZKB_TRY(g1->msm_run_split(ctx, side[2], pk->a, 1, zr, clamp(n_assign, pk->a, 1), 0, &res->msm_a, sorts_first ? &acc_a : nullptr));
    ZKB_TRY(g1->msm_run_split(ctx, side[4], pk->b_g1, 1, zr, clamp(n_assign, pk->b_g1, 1), 0, &res->msm_b1, sorts_first ? &acc_b1 : nullptr));
    if (sorts_first) {
      ZKB_TRY(g1->msm_run_split(ctx, side[3], pk->l, 0, zr + (s->n_inputs - 1) * Fr::N, clamp(s->n_aux, pk->l, 0), 0, &res->msm_l, &acc_l));
      ZKB_TRY(acc_b2());
      ZKB_TRY(acc_a());
      ZKB_TRY(acc_b1());
    }
    for (int i : {0, 1, 2, 4}) {
      ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_join[i], side[i]));
      ZKB_CUDA(ctx, cudaStreamWaitEvent(side[5], ctx->ev_join[i], 0));
    }
    ZKB_LAUNCH(ctx, (k_g16_coeffs<FrP, Fq, Fq2>), 3, 32, 0, side[5], (const Fr*)scal, (const Affine<Fq>*)pk->a->table,
               (const Affine<Fq>*)pk->b_g1->table, (const Affine<Fq2>*)pk->b_g2->table,
               (const Affine<Fq>*)pk->g1_singles, (const Affine<Fq2>*)pk->g2_singles, res);
    ZKB_LAUNCH(ctx, (k_g16_finish<Fq, Fq2>), 2, 32, 0, side[5], res, 0u);
    if (s->pending[0]) {
      ZKB_TRY(upload_csr(ctx, st, &s->A, s->pending[0]));
      ZKB_TRY(upload_csr(ctx, st, &s->B, s->pending[1]));
      ZKB_TRY(upload_csr(ctx, st, &s->C, s->pending[2]));
      s->pending[0] = s->pending[1] = s->pending[2] = nullptr;
    }
    ZKB_TRY(compute_h(ctx, st));
    ZKB_TRY(g1->msm_run(ctx, st, pk->h, 0, (const uint32_t*)s->va.p, clamp(s->N, pk->h, 0), 0, &res->msm_h));
    if (sorts_first) ZKB_TRY(acc_l());
    else ZKB_TRY(g1->msm_run(ctx, side[3], pk->l, 0, zr + (s->n_inputs - 1) * Fr::N, clamp(s->n_aux, pk->l, 0), 0, &res->msm_l));
    for (int i : {3, 5}) {
      ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_join[i], side[i]));
      ZKB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
    }
    ZKB_LAUNCH(ctx, (k_g16_finish<Fq, Fq2>), 1, 32, 0, st, res, 2u);
    return ZKB_OK;
  }

  static int fetch_proof(zkb_ctx* ctx, const zkb_pk*, uint64_t* proof_xy, uint8_t* proof_inf) {
    Groth16Stage* s = ctx->stage;
    if (!s || !s->results) return set_err(ctx, ZKB_E_INVALID, "groth16: no proof computed");
    cudaStream_t st = ctx->main;
    Res* res = (Res*)s->results;
    constexpr size_t kProofBytes = 2 * sizeof(Affine<Fq>) + sizeof(Affine<Fq2>);
    static_assert(offsetof(Res, proof_b) - offsetof(Res, proof_a) == sizeof(Affine<Fq>), "proof layout");
    static_assert(offsetof(Res, proof_c) - offsetof(Res, proof_b) == sizeof(Affine<Fq2>), "proof layout");
    static_assert(offsetof(Res, inf) - offsetof(Res, proof_a) == kProofBytes, "proof layout");
    char* bounce = (char*)ctx->pinned;
    ZKB_CUDA(ctx, cudaMemcpyAsync(bounce, &res->proof_a, kProofBytes + 16, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(proof_xy, bounce, kProofBytes);
    const uint32_t* inf = (const uint32_t*)(bounce + kProofBytes);
    for (int i = 0; i < 3; i++) proof_inf[i] = inf[i] ? 1 : 0;
    return ZKB_OK;
  }

  // ---- sharded path ------------------------------------------------------------------------
  // this rank's partial (A_k, C_k, B2_k) -> stage->shard->part; everything enqueued, nothing synchronised
  static int prove_partial_staged(zkb_ctx* ctx, const zkb_pk* pk, const uint64_t* r, const uint64_t* sc) {
    Groth16Stage* s = ctx->stage;
    if (!s || !s->staged || s->curve != CURVE) return set_err(ctx, ZKB_E_INVALID, "groth16: nothing staged");
    if (!pk || pk->curve != CURVE || pk->ctx != ctx || !pk->sharded)
      return set_err(ctx, ZKB_E_INVALID, "groth16: not a sharded proving key of this context");
    if (!r || !sc) return set_err(ctx, ZKB_E_INVALID, "groth16: null r/s");
    cudaStream_t st = ctx->main;
    cudaStream_t side[kNumSideStreams];
    for (int i = 0; i < kNumSideStreams; i++) side[i] = ctx->serial ? st : ctx->side[i];
    const GroupOps* g1 = group_ops(CURVE, ZKB_G1);
    const GroupOps* g2 = group_ops(CURVE, ZKB_G2);
    if (!s->shard) ZKB_CUDA(ctx, cudaMalloc(&s->shard, kStageBlockBytes));
    Res* res = (Res*)s->results;
    Shard* sh = (Shard*)s->shard;
    Fr* scal = (Fr*)s->scal;
    Fr r_val, s_val;
    memcpy(r_val.v, r, 32);
    memcpy(s_val.v, sc, 32);
    ZKB_TRY(fork_streams(ctx, 1));
    ZKB_LAUNCH(ctx, (k_g16_scalars<FrP, Fq, Fq2>), 4, 32, 0, side[0], r_val, s_val, scal, (const Affine<Fq>*)pk->g1_singles,
               (const Affine<Fq2>*)pk->g2_singles, res);
    ZKB_LAUNCH(ctx, (k_g16_fixed<FrP, Fq, Fq2>), 3, 32, 0, side[0], (const Fr*)scal, (const Affine<Fq>*)pk->q0_g1,
               (const Affine<Fq2>*)pk->q0_g2, (const Affine<Fq>*)pk->g1_singles, (const Affine<Fq2>*)pk->g2_singles,
               (const Res*)res, sh);
    const size_t n_assign = s->n_inputs - 1 + s->n_aux;
    ZKB_TRY(fr_convert_dev(ctx, st, CURVE, (const Fr*)s->z.p + 1, s->z_repr.p, n_assign, 0));
    const uint32_t* zr = (const uint32_t*)s->z_repr.p;
    // local pair range of MSM `which` against a scalar vector of `avail` elements
    auto local_n = [&](int which, size_t avail) -> size_t {
      size_t lo = pk->pair_lo[which], hi = lo + pk->pair_n[which];
      if (hi > avail) hi = avail;
      return hi > lo ? hi - lo : 0;
    };
    ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    for (int i = 1; i <= 5; i++) ZKB_CUDA(ctx, cudaStreamWaitEvent(side[i], ctx->ev_fork, 0));
    ZKB_TRY(g2->msm_run(ctx, side[1], pk->b_g2, 0, zr + pk->pair_lo[2] * Fr::N, local_n(2, n_assign), 0, &sh->part.b2));
    ZKB_TRY(g1->msm_run(ctx, side[2], pk->a, 0, zr + pk->pair_lo[0] * Fr::N, local_n(0, n_assign), 0, &sh->part.a));
    ZKB_TRY(g1->msm_run(ctx, side[4], pk->b_g1, 0, zr + pk->pair_lo[1] * Fr::N, local_n(1, n_assign), 0, &sh->msm_b1));
    for (int i : {0, 2, 4}) {
      ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_join[i], side[i]));
      ZKB_CUDA(ctx, cudaStreamWaitEvent(side[5], ctx->ev_join[i], 0));
    }
    ZKB_LAUNCH(ctx, (k_g16_local_mul<FrP, Fq, Fq2>), 2, 32, 0, side[5], (const Fr*)scal, sh);
    if (s->pending[0]) {
      ZKB_TRY(upload_csr(ctx, st, &s->A, s->pending[0]));
      ZKB_TRY(upload_csr(ctx, st, &s->B, s->pending[1]));
      ZKB_TRY(upload_csr(ctx, st, &s->C, s->pending[2]));
      s->pending[0] = s->pending[1] = s->pending[2] = nullptr;
    }
    ZKB_TRY(compute_h(ctx, st));       // every rank computes the whole h (7 transforms), then multiplies its slice
    ZKB_TRY(g1->msm_run(ctx, st, pk->h, 0, (const uint32_t*)s->va.p + pk->pair_lo[3] * Fr::N, local_n(3, s->N), 0, &sh->msm_h));
    ZKB_TRY(g1->msm_run(ctx, side[3], pk->l, 0, zr + (s->n_inputs - 1 + pk->pair_lo[4]) * Fr::N, local_n(4, s->n_aux), 0,
                        &sh->msm_l));
    for (int i : {1, 3, 5}) {
      ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_join[i], side[i]));
      ZKB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
    }
    ZKB_LAUNCH(ctx, (k_g16_local_c<Fq, Fq2>), 1, 32, 0, st, sh);
    return ZKB_OK;
  }

  static int fetch_partial(zkb_ctx* ctx, void* out) {
    Groth16Stage* s = ctx->stage;
    if (!s || !s->shard) return set_err(ctx, ZKB_E_INVALID, "groth16: no partial computed");
    ZKB_CUDA(ctx, cudaMemcpyAsync(out, &((Shard*)s->shard)->part, sizeof(Part), cudaMemcpyDeviceToHost, ctx->main));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->main));
    return ZKB_OK;
  }

  // partials == nullptr: all-gather this rank's stage->shard->part over the communicator (the product path);
  // otherwise `count` partials from the host (transport supplied by the caller; the single-GPU tests).
  // recompute_fixed: run the r / s dependent kernels again (a fold that did not follow prove_partial_staged on this ctx).
  static int fold_partials(zkb_ctx* ctx, const zkb_pk* pk, const void* partials, size_t count, const uint64_t* r,
                           const uint64_t* sc, bool recompute_fixed) {
    Groth16Stage* s = ctx->stage;
    if (!pk || pk->curve != CURVE || pk->ctx != ctx || !pk->sharded)
      return set_err(ctx, ZKB_E_INVALID, "groth16: not a sharded proving key of this context");
    if (!s) { ctx->stage = new Groth16Stage(); s = ctx->stage; }
    if (!s->results) ZKB_CUDA(ctx, cudaMalloc(&s->results, kStageBlockBytes));
    if (!s->scal) ZKB_CUDA(ctx, cudaMalloc(&s->scal, sizeof(Fr) * 4));
    if (!s->shard) ZKB_CUDA(ctx, cudaMalloc(&s->shard, kStageBlockBytes));
    cudaStream_t st = ctx->main;
    Res* res = (Res*)s->results;
    Shard* sh = (Shard*)s->shard;
    if (recompute_fixed) {
      if (!r || !sc) return set_err(ctx, ZKB_E_INVALID, "groth16: null r/s");
      Fr r_val, s_val;
      memcpy(r_val.v, r, 32);
      memcpy(s_val.v, sc, 32);
      ZKB_LAUNCH(ctx, (k_g16_scalars<FrP, Fq, Fq2>), 4, 32, 0, st, r_val, s_val, (Fr*)s->scal, (const Affine<Fq>*)pk->g1_singles,
                 (const Affine<Fq2>*)pk->g2_singles, res);
      ZKB_LAUNCH(ctx, (k_g16_fixed<FrP, Fq, Fq2>), 3, 32, 0, st, (const Fr*)s->scal, (const Affine<Fq>*)pk->q0_g1,
                 (const Affine<Fq2>*)pk->q0_g2, (const Affine<Fq>*)pk->g1_singles, (const Affine<Fq2>*)pk->g2_singles,
                 (const Res*)res, sh);
    }
    void* d_all;
    if (partials) {
      ZKB_TRY(comm_gather_buffer(ctx, sizeof(Part) * count, &d_all));
      ZKB_CUDA(ctx, cudaMemcpyAsync(d_all, partials, sizeof(Part) * count, cudaMemcpyDefault, st));
    } else {
      count = (size_t)ctx->n_ranks;
      if (pk->n_ranks != ctx->n_ranks || pk->rank != ctx->rank)
        return set_err(ctx, ZKB_E_INVALID, "groth16: key sharded for rank %d of %d, communicator is rank %d of %d", pk->rank,
                       pk->n_ranks, ctx->rank, ctx->n_ranks);
      ZKB_TRY(comm_gather_buffer(ctx, sizeof(Part) * count, &d_all));
      ZKB_TRY(comm_allgather(ctx, st, &sh->part, d_all, sizeof(Part)));
    }
    ZKB_LAUNCH(ctx, (k_g16_fold_finish<Fq, Fq2>), 3, 32, 0, st, (const Part*)d_all, (uint32_t)count, (const Shard*)sh, res);
    return ZKB_OK;
  }

  static const Groth16Ops* ops() {
    static const Groth16Ops o = {&stage, &compute_h, &prove_staged, &fetch_proof, &fetch_h,
                                 sizeof(Affine<Fq>), sizeof(Affine<Fq2>),
                                 &prove_partial_staged, &fetch_partial, &fold_partials, sizeof(Part)};
    return &o;
  }
};

}  // namespace zkb
