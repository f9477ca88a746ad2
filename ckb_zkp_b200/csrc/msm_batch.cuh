// Bucket accumulation in AFFINE coordinates with batched inversion (Montgomery's trick), one inversion per thread.
//
// The XYZZ loop of msm.cuh (k_accumulate) spends 10 multiplications per bucket entry and runs at ~85 % of the
// IMAD.WIDE pipe, so only fewer multiplications make the MSM faster.  An affine addition costs 1 inversion + 3
// multiplications; sharing the inversion between k independent additions costs 3 more multiplications each, i.e.
// 6 per entry once the inversion is amortised.  Additions into the SAME bucket depend on each other, additions into
// different buckets do not, hence:
//
//   * a thread owns G bucket lists (G consecutive positions of the size-ordered bucket list, interleaved across the
//     warp so that the 32 lanes hold lists of the same length) and walks them in rounds: round r adds entry r of each
//     of its lists to that list's running affine sum (kept in global memory, it stays in L1 / L2);
//   * forward sweep over the G lists: d_g = x(P_g) - x(acc_g), prefix products in shared memory; ONE inversion of the
//     product per thread and round (Fp::inv_safegcd: division steps in batches of 30, ~25 k instructions of which only
//     ~4 k use the multiplier, so it hides in the issue slots the 4-cycle IMAD.WIDE leaves free); backward sweep:
//     1 / d_g = running inverse * prefix_g, lambda, x3, y3;
//   * no block-level synchronisation, no second pass over the entries, no extra sort: the kernel consumes the very same
//     (entries, offsets, order, sched) as k_accumulate and writes the same bucket array (Z = 1), so the chunked path
//     for giant buckets and the bucket reduction are unchanged.
//
// Rare operand pairs (acc == identity after a cancellation, x(P) == x(acc): doubling or P + (-P)) are flagged in the
// forward sweep (d_g := 1) and finished through the XYZZ formulas with their own inversion in the backward sweep.
// Measured against the formulation as a tree of pair levels (msm_affine.cuh, block-shared inversion: 44 % of the
// multiplier pipe, barrier and instruction-fetch stalls): DESIGN.md section 4b.
#pragma once
#include "common.cuh"
#include "curve.cuh"
#include "devutil.cuh"

namespace zkb {

constexpr int kBatchThreads = 64;            // threads per block: small blocks keep the tail of the grid short

template <class F>
struct BatchGeom {
  static constexpr int CH = sizeof(F) / 16;                  // 16-byte chunks of one field element
  // lists per thread: the prefix products (G * sizeof(F) per thread) bound it through shared memory
  static constexpr int G = sizeof(F) <= 48 ? 12 : 8;
  static constexpr size_t kSmemPerThread = (size_t)G * (sizeof(F) + 8);
  static constexpr size_t kSmem = kSmemPerThread * kBatchThreads;
};

// chunk c of element g of thread t: conflict-free 128-bit accesses across a warp
template <class F>
__device__ __forceinline__ void smem_put(uint4* base, int g, const F& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int c = 0; c < BatchGeom<F>::CH; c++) base[(g * BatchGeom<F>::CH + c) * kBatchThreads] = s[c];
}
template <class F>
__device__ __forceinline__ F smem_get(const uint4* base, int g) {
  F v;
  uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
  for (int c = 0; c < BatchGeom<F>::CH; c++) d[c] = base[(g * BatchGeom<F>::CH + c) * kBatchThreads];
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// F: the coordinate field with the multiplication as a call (CallVariant), same bytes as the table's field
template <class F>
__global__ void __launch_bounds__(kBatchThreads)
k_accumulate_batch(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets,
                   const uint32_t* __restrict__ order, const MsmSched* __restrict__ sched,
                   const Affine<F>* __restrict__ table, Affine<F>* __restrict__ acc, XYZZ<F>* __restrict__ bucket_acc) {
  constexpr int G = BatchGeom<F>::G;
  extern __shared__ uint4 batch_smem[];
  uint4* prefix = batch_smem + threadIdx.x;
  uint32_t* start = reinterpret_cast<uint32_t*>(batch_smem + (size_t)G * BatchGeom<F>::CH * kBatchThreads) + threadIdx.x;
  uint32_t* len = start + G * kBatchThreads;

  const uint32_t n_reg = sched->n_regular;
  const uint32_t warp_global = (blockIdx.x * kBatchThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t warp_first = warp_global * 32u * G;
  if (warp_first >= n_reg) return;                       // warp-uniform
  const uint32_t p0 = warp_first + lane;                 // list g of this thread sits at sorted position p0 + 32 g

  // ---- round 0: every running sum starts as the list's first point
  uint32_t rounds = 0;
  for (int g = 0; g < G; g++) {
    const uint32_t p = p0 + 32u * g;
    uint32_t s = 0, l = 0;
    if (p < n_reg) {
      const uint32_t b = order[p];
      s = offsets[b];
      l = offsets[b + 1] - s;
      const uint32_t e = entries[s];
      Affine<F> P = ld_vec(&table[e & 0x7fffffffu]);
      if (e >> 31) P.y = F::neg(P.y);
      st_vec(&acc[p], P);
      if (l > 1) {
        const char* nx = reinterpret_cast<const char*>(&table[entries[s + 1] & 0x7fffffffu]);
        prefetch_l2(nx);
        prefetch_l2(nx + sizeof(Affine<F>) - 1);
      }
    }
    start[g * kBatchThreads] = s;
    len[g * kBatchThreads] = l;
    rounds = l > rounds ? l : rounds;
  }
  rounds = __reduce_max_sync(0xffffffffu, rounds);

  for (uint32_t r = 1; r < rounds; r++) {
    // ---- forward: denominators and their prefix products
    F run = F::one();
    uint32_t special = 0;                                  // bit g: finished by the XYZZ formulas
    for (int g = 0; g < G; g++) {
      smem_put(prefix, g, run);
      if (r < len[g * kBatchThreads]) {
        const uint32_t pos = start[g * kBatchThreads] + r;
        const uint32_t e = entries[pos];
        if (r + 1 < len[g * kBatchThreads]) {              // the next round's point travels to L2 meanwhile
          const char* nx = reinterpret_cast<const char*>(&table[entries[pos + 1] & 0x7fffffffu]);
          prefetch_l2(nx);
          prefetch_l2(nx + sizeof(Affine<F>) - 1);
        }
        const F px = ld_vec(&table[e & 0x7fffffffu].x);
        const F ax = ld_vec_rw(&acc[p0 + 32u * g].x);
        const F d = F::sub(px, ax);
        if (d.is_zero() || ax.is_zero()) special |= 1u << g;
        else run = F::mul(run, d);
      }
    }
    F inv = F::inv_fast(run);
    // ---- backward: finish the additions
    for (int g = G - 1; g >= 0; g--) {
      if (r >= len[g * kBatchThreads]) continue;
      const uint32_t p = p0 + 32u * g;
      const uint32_t e = entries[start[g * kBatchThreads] + r];
      Affine<F> P = ld_vec(&table[e & 0x7fffffffu]);
      if (e >> 31) P.y = F::neg(P.y);
      Affine<F> A = ld_vec_rw(&acc[p]);
      if ((special >> g) & 1u) {                           // identity operand, doubling or cancellation
        XYZZ<F> t = XYZZ<F>::from_affine(A);
        pt_madd(t, P, false);
        pt_to_affine(A, t);
      } else {
        const F d = F::sub(P.x, A.x);
        const F inv_d = F::mul(inv, smem_get<F>(prefix, g));
        inv = F::mul(inv, d);
        const F lam = F::mul(F::sub(P.y, A.y), inv_d);
        const F x3 = F::sub(F::sub(F::sqr(lam), A.x), P.x);
        A.y = F::sub(F::mul(lam, F::sub(A.x, x3)), A.y);
        A.x = x3;
      }
      st_vec(&acc[p], A);
    }
  }

  // ---- the bucket array takes the sums with Z = 1
  for (int g = 0; g < G; g++) {
    const uint32_t p = p0 + 32u * g;
    if (p < n_reg && len[g * kBatchThreads]) {
      const Affine<F> A = ld_vec_rw(&acc[p]);
      st_vec(&bucket_acc[order[p]], XYZZ<F>::from_affine(A));
    }
  }
}

}  // namespace zkb
