// Groth16 prove path on the device: witness_map (r1cs_to_qap.rs:113-172), the into_repr
// sweeps, the five MSMs and the proof assembly of prover.rs:148-210.
#pragma once
#include "common.cuh"
#include "msm.cuh"
#include "ntt.cuh"

struct zkb_pk {
  zkb_ctx* ctx;
  int curve;
  zkb_srs *a, *b_g1, *b_g2, *h, *l;
  void* g1_singles;   // device Affine<Fq>[3]: alpha, beta, delta
  void* g2_singles;   // device Affine<Fq2>[2]: beta, delta
  // Sharded key (zkb_groth16_pk_create_sharded): every srs holds only this rank's slice of the MSM pairs
  // (a / b_g1 / b_g2: query[1 + lo .. 1 + hi) against assignment[lo .. hi); h, l: query[lo .. hi)), index 0 of
  // the three coefficient queries lives in q0_*.  Order of the arrays: a, b_g1, b_g2, h, l.
  bool sharded = false;
  int n_ranks = 1, rank = 0;
  size_t pair_lo[5] = {0, 0, 0, 0, 0}, pair_n[5] = {0, 0, 0, 0, 0};
  void* q0_g1 = nullptr;    // device Affine<Fq>[2]: a_query[0], b_g1_query[0]
  void* q0_g2 = nullptr;    // device Affine<Fq2>[1]: b_g2_query[0]
};

namespace zkb {

// growable device buffer (capacity in bytes)
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct DevCsr {
  DevBuf row_ptr, col_idx, coeff;
  size_t n_rows = 0, nnz = 0;
};

// The result / shard blocks of the stage are curve-typed structs (groth16_impl.cuh) but the stage outlives a change of
// curve on the same context: both blocks are allocated at this fixed size, which either curve's layout fits
constexpr size_t kStageBlockBytes = 32768;

struct Groth16Stage {
  int curve = -1;
  size_t n_inputs = 0, n_aux = 0, n_rows = 0, N = 0;
  unsigned log_n = 0;
  DevCsr A, B, C;
  DevBuf z;         // Fr[n_inputs + n_aux] Montgomery
  DevBuf z_repr;    // Fr[n_inputs + n_aux - 1] canonical (skips ONE)
  DevBuf va, vb, vc, scratch;   // Fr[N] each
  void* results = nullptr;   // device block holding MSM results and the proof (layout in groth16_impl.cuh)
  void* shard = nullptr;     // device block of the sharded prove path (G16Shard in groth16_impl.cuh)
  void* scal = nullptr;      // device Fr[4]: r, s, r*s (canonical), spare
  bool staged = false;
  // matrices whose upload is deferred into prove_staged (host pointers, valid for the duration of zkb_groth16_prove)
  const zkb_csr* pending[3] = {nullptr, nullptr, nullptr};
};

struct Groth16Ops {
  int (*stage)(zkb_ctx*, const zkb_pk*, const zkb_csr*, const zkb_csr*, const zkb_csr*, const uint64_t*, size_t, size_t,
               int defer_matrices);
  int (*compute_h)(zkb_ctx*, cudaStream_t);        // staged inputs -> h (canonical) in stage->va
  int (*prove_staged)(zkb_ctx*, const zkb_pk*, const uint64_t*, const uint64_t*);
  int (*fetch_proof)(zkb_ctx*, const zkb_pk*, uint64_t*, uint8_t*);
  int (*fetch_h)(zkb_ctx*, uint64_t*);
  size_t g1_affine_bytes, g2_affine_bytes;
  // sharded prove path: this rank's partial (device, stage->shard), then gather + fold into the proof
  int (*prove_partial_staged)(zkb_ctx*, const zkb_pk*, const uint64_t*, const uint64_t*);
  int (*fetch_partial)(zkb_ctx*, void* partial_out);
  int (*fold_partials)(zkb_ctx*, const zkb_pk*, const void* partials_host_or_null, size_t count, const uint64_t*,
                       const uint64_t*, bool recompute_fixed);
  size_t partial_bytes;
};
const Groth16Ops* groth16_ops(int curve);
void groth16_free_stage(zkb_ctx* ctx);

}  // namespace zkb
