// Radix-2 NTT over the scalar field, natural order in and out.
//
// Replaces ark-poly 0.2 `Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place`
// (un-vendored; call sites groth16/src/r1cs_to_qap.rs:144-169).  Same transform (same root
// of unity: 2-adic root squared down to the domain size, same coset generator g, 1/N on the
// inverse).  Schedule: Cooley-Tukey DIT split into passes of up to 10 stages; each pass
// stages a 1024-element tile in shared memory (struct-of-arrays, bank-conflict-free for
// unit-stride lanes), the first pass gathers its tile in bit-reversed order, later passes
// read 2^t-element contiguous runs.  Coset scaling, 1/N and the twiddles come from tables
// kept resident in HBM per (field, log_n).
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace zkb {

struct NttDomain {
  int curve;
  unsigned log_n;
  size_t n;
  void* tw;         // omega^i,        i < n/2
  void* tw_inv;     // omega^-i,       i < n/2
  void* coset;      // g^i,            i < n
  void* coset_inv;  // g^-i / n,       i < n
  void* consts;     // [omega, omega_inv, n_inv, g, g_inv, 1/(g^n - 1)]
};
enum { kConstOmega = 0, kConstOmegaInv, kConstNInv, kConstG, kConstGInv, kConstZInv, kConstGInvScaled, kNumConsts };

int ntt_get_domain(zkb_ctx* ctx, int curve, unsigned log_n, NttDomain** out);
// in-place transform of d_data (2^log_n elements) on stream st; d_scratch: same size
int ntt_run(zkb_ctx* ctx, cudaStream_t st, NttDomain* dom, void* d_data, void* d_scratch, unsigned flags);
void ntt_free_domains(zkb_ctx* ctx);

}  // namespace zkb
