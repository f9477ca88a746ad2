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

struct Groth16Stage {
  int curve = -1;
  size_t n_inputs = 0, n_aux = 0, n_rows = 0, N = 0;
  unsigned log_n = 0;
  DevCsr A, B, C;
  DevBuf z;         // Fr[n_inputs + n_aux] Montgomery
  DevBuf z_repr;    // Fr[n_inputs + n_aux - 1] canonical (skips ONE)
  DevBuf va, vb, vc, scratch;   // Fr[N] each
  void* results = nullptr;   // device block holding MSM results and the proof (layout in groth16_impl.cuh)
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
};
const Groth16Ops* groth16_ops(int curve);
void groth16_free_stage(zkb_ctx* ctx);

}  // namespace zkb
