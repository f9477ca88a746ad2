/*
 * zkb.h -- C ABI of the B200 proving backend for ckb-zkp's Groth16 / Marlin prove path.
 *
 * The reference (sec-bit/ckb-zkp @ 8f2141a) has no FFI: its "interface" for this path is
 * Rust generics resolved at compile time.  Each entry point below replaces one of those
 * call sites; INTEGRATION.md shows the Rust `extern "C"` binding and the patched bodies.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ZKB_E_* code; never throws/aborts;
 *     zkb_last_error() gives a human-readable message for the calling ctx.
 *   - caller owns all host buffers; the library owns device memory behind opaque handles.
 *   - field elements: little-endian u64 limbs exactly as ark-ff 0.2 stores them
 *     (Fr: 4 limbs for both curves; Fq: 4 limbs BN254, 6 limbs BLS12-381).
 *       "mont"      = Montgomery form, R = 2^(64*limbs)  (the in-memory form of `Fp256/Fp384`)
 *       "canonical" = plain integer < modulus             (the form `into_repr()` yields)
 *   - affine points: x || y in Montgomery form (Fq2 = c0 || c1), plus one infinity byte per
 *     point (ark `GroupAffine { x, y, infinity }` marshalled once at upload).  All group
 *     results are returned as canonical-form affine points, so byte compare == group equality.
 *   - one ctx per GPU per host thread (internally serialised).
 */
#ifndef ZKB_H
#define ZKB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKB_BN254 0
#define ZKB_BLS12_381 1
#define ZKB_G1 1
#define ZKB_G2 2

#define ZKB_OK 0
#define ZKB_E_INVALID (-1)   /* bad argument (null pointer, unknown curve, size overflow)     */
#define ZKB_E_CUDA (-2)      /* CUDA runtime error; see zkb_last_error                         */
#define ZKB_E_TOO_LARGE (-3) /* domain larger than the field's 2-adicity:
                                SynthesisError::PolynomialDegreeTooLarge (r1cs/src/error.rs:15) */
#define ZKB_E_NO_DEVICE (-4) /* no usable CUDA device: there is no CPU fallback                */

/* zkb_ntt flags */
#define ZKB_NTT_INVERSE 1u   /* ifft_in_place:  includes the 1/N scaling                       */
#define ZKB_NTT_COSET 2u     /* coset_fft / coset_ifft with g = Fr::multiplicative_generator() */

/* zkb_srs_upload flags */
#define ZKB_SRS_PRECOMPUTE 1u /* keep 2^(c*j) * P_i for every window j resident in HBM          */

typedef struct zkb_ctx zkb_ctx;
typedef struct zkb_srs zkb_srs;
typedef struct zkb_pk zkb_pk;

/* Sparse matrix in CSR form: the flattening of ProvingAssignment::{at,bt,ct}
 * (groth16/src/prover.rs:16-25): row i holds (coeff, column) pairs with
 * column = Input(i) -> i, Aux(i) -> num_inputs + i (groth16/src/r1cs_to_qap.rs:34-37).
 * Duplicate columns inside a row are allowed (r1cs/src/impl_lc.rs:58-70). */
typedef struct zkb_csr {
  size_t n_rows;
  size_t nnz;
  const uint32_t* row_ptr;   /* n_rows + 1 */
  const uint32_t* col_idx;   /* nnz */
  const uint64_t* coeff_mont;/* nnz * 4 limbs, Montgomery */
} zkb_csr;

/* ---- context ------------------------------------------------------------------------- */
int zkb_init(int device, zkb_ctx** out);
void zkb_destroy(zkb_ctx* ctx);
const char* zkb_last_error(zkb_ctx* ctx);
/* cudaStream_t of the ctx's main stream (for callers that time with CUDA events). */
void* zkb_stream(zkb_ctx* ctx);
int zkb_sync(zkb_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's `gpu_launches`). */
uint64_t zkb_launch_count(zkb_ctx* ctx);

/* CUDA-event timing of the dominant kernel (MSM bucket accumulation) on its launch stream:
 * enable resets the counters; read synchronises and returns the summed duration, the number of
 * launches and the algorithmic bytes (n * (32 + sizeof affine base) per MSM) they covered. */
/* Measurement aid: with on != 0 the Groth16 prove path enqueues every kernel on ONE stream, so the per-kernel
 * event timings of zkb_prof_* are kernel durations (no queueing behind concurrent kernels). */
int zkb_set_serial(zkb_ctx* ctx, int on);
int zkb_prof_enable(zkb_ctx* ctx, int on);
int zkb_prof_read(zkb_ctx* ctx, double* ms_total, uint64_t* launches, double* alg_bytes_total);

/* ---- bases resident in HBM --------------------------------------------------------------
 * Replaces the `&[G::Affine]` argument of ark_ec::msm::VariableBaseMSM::multi_scalar_mul
 * (call sites groth16/src/prover.rs:187,190,220; marlin/src/pc/kzg10.rs:109,118,137,146;
 * curve/src/lib.rs:44).  Uploaded once per Parameters / CommitterKey, reused by every proof. */
int zkb_srs_upload(zkb_ctx* ctx, int curve, int group, const uint64_t* xy_mont, const uint8_t* inf,
                   size_t n, unsigned flags, zkb_srs** out);
void zkb_srs_free(zkb_srs* srs);
size_t zkb_srs_len(const zkb_srs* srs);

/* ---- variable-base MSM --------------------------------------------------------------------
 * VariableBaseMSM::multi_scalar_mul(&bases[base_offset..base_offset+n], &scalars[..n]).
 * scalars: n * 4 limbs, canonical (what `into_repr()` produced in prover.rs:150-161).
 * out_xy: one affine point (Montgomery), out_inf: 1 if the result is the identity. */
int zkb_msm(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint64_t* scalars_canonical,
            size_t n, uint64_t* out_xy, uint8_t* out_inf);
/* Curve::vartime_multiscalar_mul (curve/src/lib.rs:38-45): scalars in Montgomery form,
 * `into_repr` is fused on the device. */
int zkb_msm_mont(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint64_t* scalars_mont,
                 size_t n, uint64_t* out_xy, uint8_t* out_inf);
/* Same, with the scalars already in device memory (device pointer) -- used to time the
 * kernel path alone. */
int zkb_msm_dev(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const void* d_scalars_canonical,
                size_t n, uint64_t* out_xy, uint8_t* out_inf);

/* k independent MSMs in one call -- the shape of the reference's commitment loops: PC::commit commits polynomial by
 * polynomial (marlin/src/pc/mod.rs:42-69), the `Curve::vartime_multiscalar_mul` consumers commit vector by vector
 * (spartan/src/commitments.rs:42-56, asvc/src/lib.rs:160-225).  MSM i is multi_scalar_mul(&srs[i][base_offsets[i]..],
 * scalars[i][..n[i]]) (zip semantics); the k sorts, accumulations and bucket reductions run concurrently on the
 * library's side streams, the k results are converted and fetched together.  scalars[i] may be host or device memory;
 * scalars_mont selects Montgomery (into_repr fused) or canonical input for all of them.  out_xy: k affine points of
 * the respective group back to back (all SRS handles of one call must belong to the same curve and group).
 * k <= 4096; a call whose MSMs are ALL short (n[i] <= 16, k >= 32: the g_ic of every proof of a batch verifier,
 * groth16/src/verifier.rs:27-30) may hold up to 2^20 of them and runs one thread per (scalar, base) term instead. */
int zkb_msm_batch(zkb_ctx* ctx, size_t k, const zkb_srs* const* srs, const size_t* base_offsets,
                  const uint64_t* const* scalars, const size_t* n, int scalars_mont, uint64_t* out_xy, uint8_t* out_inf);

/* ---- radix-2 NTT over Fr --------------------------------------------------------------------
 * EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place of ark-poly 0.2 as used in
 * groth16/src/r1cs_to_qap.rs:144-169: natural order in and out, 2^log_n elements of 4 limbs
 * in Montgomery form, in place. */
int zkb_ntt(zkb_ctx* ctx, int curve, uint64_t* data_mont, unsigned log_n, unsigned flags);
int zkb_ntt_dev(zkb_ctx* ctx, int curve, void* d_data_mont, unsigned log_n, unsigned flags);

/* ---- Groth16 ------------------------------------------------------------------------------ */
/* R1CStoQAP::witness_map (groth16/src/r1cs_to_qap.rs:113-172) followed by the into_repr sweep
 * of prover.rs:161.  z_mont = input_assignment ++ aux_assignment (n_inputs + n_aux elements,
 * z[0] = ONE).  h_canonical receives next_pow2(n_rows + n_inputs) * 4 limbs. */
int zkb_groth16_h(zkb_ctx* ctx, int curve, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                  const uint64_t* z_mont, size_t n_inputs, size_t n_aux, uint64_t* h_canonical);

/* Parameters<E> (groth16/src/lib.rs:81-91) made resident.  Point arrays are x||y Montgomery +
 * infinity bytes; singles are [alpha_g1, beta_g1, delta_g1] and [beta_g2, delta_g2]. */
int zkb_groth16_pk_create(zkb_ctx* ctx, int curve,
                          const uint64_t* a_query, const uint8_t* a_inf, size_t a_len,
                          const uint64_t* b_g1_query, const uint8_t* b_g1_inf, size_t b_g1_len,
                          const uint64_t* b_g2_query, const uint8_t* b_g2_inf, size_t b_g2_len,
                          const uint64_t* h_query, const uint8_t* h_inf, size_t h_len,
                          const uint64_t* l_query, const uint8_t* l_inf, size_t l_len,
                          const uint64_t* g1_singles /* alpha, beta, delta */, const uint64_t* g2_singles /* beta, delta */,
                          zkb_pk** out);
void zkb_groth16_pk_free(zkb_pk* pk);

/* create_proof (groth16/src/prover.rs:124-211) from "prover filled" (:146) to "Proof assembled"
 * (:206): witness_map, into_repr, 5 MSMs, final assembly, into_affine.
 * r, s: 4 limbs each, canonical.  proof_xy = A (G1) || B (G2) || C (G1) affine Montgomery;
 * proof_inf[3] = infinity flags. */
int zkb_groth16_prove(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                      const uint64_t* z_mont, size_t n_inputs, size_t n_aux,
                      const uint64_t r_canonical[4], const uint64_t s_canonical[4],
                      uint64_t* proof_xy, uint8_t* proof_inf);

/* Two-phase variant used to time the device path with inputs resident in HBM:
 * zkb_groth16_stage copies matrices and assignment to the device, zkb_groth16_prove_staged
 * runs the whole prove path from there (the result stays on the device until
 * zkb_groth16_fetch_proof). */
int zkb_groth16_stage(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                      const uint64_t* z_mont, size_t n_inputs, size_t n_aux);
int zkb_groth16_prove_staged(zkb_ctx* ctx, const zkb_pk* pk, const uint64_t r_canonical[4],
                             const uint64_t s_canonical[4]);
int zkb_groth16_fetch_proof(zkb_ctx* ctx, const zkb_pk* pk, uint64_t* proof_xy, uint8_t* proof_inf);

/* ---- multi-GPU: one process (one ctx) per GPU, NCCL over NVLink / NVSwitch ---------------------------
 * The reference runs in one address space (rayon, groth16/Cargo.toml:14); what shards here is what is independent
 * there: the pairs of every MSM (a plain sum, curve/src/lib.rs:38-45; the five Groth16 MSMs share only read-only
 * inputs, groth16/src/prover.rs:164-190; Marlin commits polynomial by polynomial, marlin/src/pc/mod.rs:42-69).
 * Each rank keeps a contiguous slice of the bases resident, computes its partial sum, and the ranks exchange ONE
 * all-gather of fixed-size partial points (NCCL has no EC-add reduction) which every rank folds in rank order:
 * all ranks return the identical canonical affine result.
 * Rendezvous: rank 0 calls zkb_comm_unique_id, the host distributes the 128 bytes by its own means
 * (torch.distributed / MPI / a file), every rank calls zkb_comm_init (collective).  NCCL is dlopen'ed
 * (libnccl.so.2); single-GPU users never need it. */
#define ZKB_COMM_ID_BYTES 128
int zkb_comm_unique_id(zkb_ctx* ctx, uint8_t id[ZKB_COMM_ID_BYTES]);
int zkb_comm_init(zkb_ctx* ctx, int n_ranks, int rank, const uint8_t id[ZKB_COMM_ID_BYTES]);
void zkb_comm_destroy(zkb_ctx* ctx);
int zkb_comm_rank(zkb_ctx* ctx);
int zkb_comm_size(zkb_ctx* ctx);
/* all-gathers this ctx has enqueued so far (bench.py reports it next to gpu_launches) */
uint64_t zkb_comm_collectives(zkb_ctx* ctx);

/* This rank's slice [global_lo, global_lo + n_local) of a logical SRS of global_n bases. */
int zkb_srs_upload_shard(zkb_ctx* ctx, int curve, int group, const uint64_t* xy_mont_local, const uint8_t* inf_local,
                         size_t n_local, size_t global_lo, size_t global_n, unsigned flags, zkb_srs** out);
/* VariableBaseMSM::multi_scalar_mul(&bases[base_offset..], &scalars[..n]) over the LOGICAL SRS, every rank passing the
 * same full-length scalar array (host or device; only the local slice is read) -- the call shape of the reference,
 * where every rank holds the same polynomial / assignment.  Collective: local partial, one all-gather, fold. */
int zkb_msm_sharded(zkb_ctx* ctx, const zkb_srs* srs_shard, size_t base_offset, const uint64_t* scalars, size_t n,
                    int scalars_mont, uint64_t* out_xy, uint8_t* out_inf);
/* Same exchange for a stand-alone MSM whose scalars are sharded like the bases: d_scalars_local (device, canonical)
 * pairs with the first n_local bases of the shard (BASELINE configs[2]). */
int zkb_msm_sharded_local(zkb_ctx* ctx, const zkb_srs* srs_shard, const void* d_scalars_local, size_t n_local,
                          uint64_t* out_xy, uint8_t* out_inf);
/* The two halves of zkb_msm_sharded for a host that brings its own transport: the rank's partial point
 * (zkb_partial_bytes(curve, group) bytes, XYZZ coordinates), and the fold of `count` partials in index order. */
size_t zkb_partial_bytes(int curve, int group);
int zkb_msm_partial(zkb_ctx* ctx, const zkb_srs* srs_shard, size_t base_offset, const uint64_t* scalars, size_t n,
                    int scalars_mont, void* partial_out);
int zkb_msm_fold(zkb_ctx* ctx, int curve, int group, const void* partials, size_t count, uint64_t* out_xy, uint8_t* out_inf);

/* Parameters<E> sharded over the ranks: the caller passes the WHOLE queries (every rank deserialises the same key
 * file) and the library keeps only this rank's slice of the pairs of each of the five MSMs resident. */
int zkb_groth16_pk_create_sharded(zkb_ctx* ctx, int curve,
                                  const uint64_t* a_query, const uint8_t* a_inf, size_t a_len,
                                  const uint64_t* b_g1_query, const uint8_t* b_g1_inf, size_t b_g1_len,
                                  const uint64_t* b_g2_query, const uint8_t* b_g2_inf, size_t b_g2_len,
                                  const uint64_t* h_query, const uint8_t* h_inf, size_t h_len,
                                  const uint64_t* l_query, const uint8_t* l_inf, size_t l_len,
                                  const uint64_t* g1_singles, const uint64_t* g2_singles, int n_ranks, int rank,
                                  zkb_pk** out);
/* ONE proof computed by all ranks together (strong scaling; same arguments and result as zkb_groth16_prove on every
 * rank): witness_map on every rank, the five MSMs over the rank's pairs, s * A_k + r * B1_k + L_k + H_k formed
 * locally, one all-gather of (A_k, C_k, B2_k), fold + into_affine. */
int zkb_groth16_prove_sharded(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                              const uint64_t* z_mont, size_t n_inputs, size_t n_aux,
                              const uint64_t r_canonical[4], const uint64_t s_canonical[4],
                              uint64_t* proof_xy, uint8_t* proof_inf);
/* after zkb_groth16_stage: the device path alone (result fetched with zkb_groth16_fetch_proof) */
int zkb_groth16_prove_sharded_staged(zkb_ctx* ctx, const zkb_pk* pk, const uint64_t r_canonical[4],
                                     const uint64_t s_canonical[4]);
/* The two halves for a host with its own transport: the rank's partial (zkb_groth16_partial_bytes(curve) bytes)
 * and the fold of `count` partials in rank order into the proof. */
size_t zkb_groth16_partial_bytes(int curve);
int zkb_groth16_prove_partial(zkb_ctx* ctx, const zkb_pk* pk, const zkb_csr* A, const zkb_csr* B, const zkb_csr* C,
                              const uint64_t* z_mont, size_t n_inputs, size_t n_aux,
                              const uint64_t r_canonical[4], const uint64_t s_canonical[4], void* partial_out);
int zkb_groth16_fold(zkb_ctx* ctx, const zkb_pk* pk, const void* partials, size_t count,
                     const uint64_t r_canonical[4], const uint64_t s_canonical[4], uint64_t* proof_xy, uint8_t* proof_inf);

/* ---- fixed-base batch multiplication (setup side; groth16/src/generator.rs:205-256) -------
 * out[i] = scalars[i] * G for one affine base point G; result-identical to ark FixedBaseMSM
 * followed by batch_normalization.  Used to mint synthetic SRS on the device. */
int zkb_fixed_base_mul(zkb_ctx* ctx, int curve, int group, const uint64_t* base_xy_mont,
                       const uint64_t* scalars_canonical, size_t n, uint64_t* out_xy, uint8_t* out_inf);

/* ---- key-file ingestion: ark-serialize 0.2 compressed points -> the ABI's point layout --------------------------
 * What `Parameters::<E>::deserialize` (groth16/src/lib.rs:81; cli/src/zkp_prove.rs:117-124) and the Marlin key types do
 * per point in ark-ec's `GroupAffine::deserialize`: x as canonical little-endian bytes (Fq2: c0 then c1; 32 / 48 bytes
 * per Fq), flags in the two top bits of the last byte (bit 7: y is the larger root, bit 6: infinity); y is recovered
 * with the Fq / Fq2 square root on the device.  out_status[i]: 0 ok, 1 coordinate not canonical (>= p), 2 x not on the
 * curve, 3 not in the prime-order subgroup (only with ZKB_DECOMPRESS_CHECK_SUBGROUP); a rejected point is returned as
 * the identity.  Buffers may be host or device memory.  The result feeds zkb_srs_upload / zkb_groth16_pk_create. */
#define ZKB_DECOMPRESS_CHECK_SUBGROUP 1u
int zkb_points_decompress(zkb_ctx* ctx, int curve, int group, const uint8_t* compressed, size_t n, unsigned flags,
                          uint64_t* out_xy_mont, uint8_t* out_inf, uint8_t* out_status);

/* ---- batched verification: products of pairings --------------------------------------------------------------------
 * What groth16/src/verifier.rs:18-44 (`verify_proof`: miller_loop over three pairs + final_exponentiation, compared
 * with alpha_g1_beta_g2) and marlin/src/pc/kzg10.rs `check` / `batch_check` obtain from ark-ec's PairingEngine, for many
 * proofs at once.  Pairs are laid out group after group: group g is pairs [g * group_size, (g + 1) * group_size) and
 *     out_gt[g] = prod_j a(P_j, Q_j)
 * with a = a reduced pairing chosen for the device: the plain ate f_{|x|,Q}(P)^(3 (q^12 - 1) / r) on BLS12-381, the optimal
 * ate on BN254 (a fixed power of the pairing ark-ec computes: equalities between products -- all these callers test --
 * hold or fail identically).
 * g1_xy / g2_xy: affine Montgomery points as everywhere in this ABI, *_inf optional identity flags (an all-zero point
 * is the identity too; a pair with an identity contributes 1).  G2 points must be in the order-r subgroup.
 * out_gt: n_groups x 12 x limbs(Fq) Montgomery limbs in the tower order of ark-ff's Fq12 (c0.c0.c0, c0.c0.c1, c0.c1.c0,
 * ... c1.c2.c1; Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v)).  One thread per pair, one per group: the call is
 * built for batches (thousands of pairs), a single 3-pair check takes tens of milliseconds.  Host or device buffers. */
int zkb_multi_pairing(zkb_ctx* ctx, int curve, const uint64_t* g1_xy_mont, const uint8_t* g1_inf,
                      const uint64_t* g2_xy_mont, const uint8_t* g2_inf, size_t n_groups, size_t group_size,
                      uint64_t* out_gt_mont);

/* ---- Fr helpers (device-side batch ops on host arrays; used by the host layers) ------------ */
/* out[i] = into_repr(in[i]) (mode 0) or from_repr(in[i]) (mode 1) */
int zkb_fr_convert(zkb_ctx* ctx, int curve, const uint64_t* in, uint64_t* out, size_t n, int mode);

/* ---- polynomial helpers of the Marlin prover (all Fr data Montgomery) -------------------------
 * Buffer arguments of this group, of zkb_fr_vec_op / zkb_fr_batch_inverse / zkb_fr_powers / zkb_spmv
 * (including the arrays a zkb_csr points at), of zkb_fr_convert and the scalar array of zkb_msm /
 * zkb_msm_mont may be HOST or DEVICE memory.  Host buffers are staged through stream-ordered scratch
 * (cudaMemcpyDefault) and the call returns when the result has landed.  Device buffers are used IN PLACE, and a
 * call of zkb_fr_vec_op / zkb_fr_batch_inverse / zkb_fr_prefix_product / zkb_fr_powers / zkb_spmv /
 * zkb_poly_lincomb / zkb_ntt_dev whose output lives in device memory returns WITHOUT synchronising: the work is
 * ordered on the ctx's stream (zkb_stream), later calls on the same ctx see it, and zkb_sync waits for it -- a
 * caller that keeps the round state resident in HBM thus pays neither copies nor a host round trip per primitive.
 * Scalars (z, coefficients) and results that feed the transcript (remainders, points) are host memory.
 * q = p / (x - z) and rem = p(z): KZG10::compute_witness_polynomial (marlin/src/pc/kzg10.rs:211-226)
 * and LabeledPolynomial::evaluate (marlin/src/lib.rs:147-156).  p has n coefficients (low degree
 * first), q receives n - 1 (may be NULL to evaluate only). */
int zkb_poly_div_linear(zkb_ctx* ctx, int curve, const uint64_t* p_mont, size_t n, const uint64_t z_mont[4],
                        uint64_t* q_mont, uint64_t rem_mont[4]);
/* k polynomial evaluations in one call: out[j] = polys[j](points[j]) -- the 21 evaluations zkp_marlin's prover sends
 * (marlin/src/lib.rs:147-156: `polynomial.evaluate(point)` in a loop over the query set).  The remainder trees are
 * enqueued back to back; one copy and one synchronisation return all values.  points_mont / out_mont: k * 4 limbs. */
int zkb_poly_eval_batch(zkb_ctx* ctx, int curve, size_t k, const uint64_t* const* polys_mont, const size_t* lens,
                        const uint64_t* points_mont, uint64_t* out_mont);

/* out[i] = sum_j coeffs[j] * polys[j][i - shifts[j]], i < out_len: the accumulation of PC::open
 * (marlin/src/pc/mod.rs:85-98; shift = supported_degree - degree_bound, :241-250).  k <= 64. */
int zkb_poly_lincomb(zkb_ctx* ctx, int curve, size_t k, const uint64_t* const* polys_mont, const size_t* lens,
                     const size_t* shifts, const uint64_t* coeffs_mont, uint64_t* out_mont, size_t out_len);
/* ark_ff::batch_inversion (marlin/src/ahp/prover.rs:365-367): out[i] = 1 / in[i], zeros stay zero. */
int zkb_fr_batch_inverse(zkb_ctx* ctx, int curve, const uint64_t* in_mont, uint64_t* out_mont, size_t n);

/* out[i] = in[0] * ... * in[i - 1], out[0] = 1: the grand-product accumulator z of PLONK's permutation argument
 * (plonk/src/ahp/indexer/permutation.rs:111-118: z.push(acc); acc *= perms[i]).  Host or device buffers. */
int zkb_fr_prefix_product(zkb_ctx* ctx, int curve, const uint64_t* in_mont, uint64_t* out_mont, size_t n);

/* Elementwise Fr vector operations (the pointwise loops of marlin/src/ahp/prover.rs:246-305,357-411):
 * op 0: a + b  1: a - b  2: a * b  3: s * a  4: a + s * b  5: s - a  6: a + s   (b / s may be NULL when unused) */
int zkb_fr_vec_op(zkb_ctx* ctx, int curve, int op, const uint64_t* a_mont, const uint64_t* b_mont, const uint64_t* s_mont,
                  uint64_t* out_mont, size_t n);
/* out[i] = scale * base^i (scale may be NULL = 1): domain elements (EvaluationDomain::elements) and coset powers */
int zkb_fr_powers(zkb_ctx* ctx, int curve, const uint64_t base_mont[4], const uint64_t* scale_mont, uint64_t* out_mont, size_t n);
/* y = M x for a CSR matrix with Montgomery coefficients: the sparse accumulations of
 * marlin/src/ahp/prover.rs:110-123 (z_A, z_B) and :259-269 (t, with the transposed matrices) */
int zkb_spmv(zkb_ctx* ctx, int curve, const zkb_csr* m, const uint64_t* x_mont, size_t n_cols, uint64_t* y_mont);

/* ---- host-side helpers of the Python host layer (no device work, usable without a GPU) ----------------------
 * marlin/src/fs_rng.rs builds its Fiat-Shamir generator from merlin (STROBE-128 over Keccak-f[1600]) and rand_chacha
 * (ChaCha20); a Rust host links those crates, the Python host layer (ckb_zkp_b200/fs_rng.py) uses these two. */
void zkb_host_keccak_f1600(uint64_t state[25]);
void zkb_host_chacha20_blocks(const uint8_t key[32], uint64_t counter, uint32_t* out_words, size_t n_blocks);

/* ---- diagnostics: single field / group operations of the device arithmetic on n operands, used by
 * the parity tests to check the GPU arithmetic against the CPU oracle in isolation.
 * field: 0 BN254 Fr, 1 BLS12-381 Fr, 2 BN254 Fq, 3 BLS12-381 Fq.
 * fp op: 0 mul 1 add 2 sub 3 inv 4 to_mont 5 from_mont 6 sqr 7 neg.
 * pt op: 0 acc += q (affine, optionally negated) 1 acc += q (XYZZ) 2 dbl 3 to_affine 4 acc * k[8]. */
int zkb_debug_fp_op(zkb_ctx* ctx, int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n);
int zkb_debug_pt_op(zkb_ctx* ctx, int curve, int group, int op, const uint32_t* acc, const uint32_t* q, int neg,
                    uint32_t* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* ZKB_H */
