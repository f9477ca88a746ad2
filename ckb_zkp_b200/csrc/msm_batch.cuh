// Bucket accumulation in AFFINE coordinates with batched inversion (Montgomery's trick), one inversion per thread and
// round, over bucket lists cut into chains of (almost) equal length.
//
// The XYZZ loop of msm.cuh (k_accumulate) spends 10 multiplications per bucket entry and runs at ~85 % of the
// IMAD.WIDE pipe, so only fewer multiplications make the MSM faster.  An affine addition costs 1 inversion + 3
// multiplications; sharing the inversion between k independent additions costs 3 more multiplications each, i.e.
// 6 per entry once the inversion is amortised.  Additions into the SAME running sum depend on each other, additions
// into different sums do not, hence:
//
//   * every regular bucket list (<= the big-bucket threshold) is cut into ceil(len / lmax) chains of equal length
//     (+-1): ~2^20 chains of 8..12 entries for a 2^20-point MSM.  Equal lengths are what keeps a lock-step schedule
//     full: with whole lists (Poisson lengths, mean 26, max ~55) the first version of this kernel ran one wave that
//     was half empty towards the end -- 13.6 ms against 6.0 ms for the XYZZ loop;
//   * the grid is one resident wave; the chains are dealt out evenly, G = chains / (32 * warps) per lane (in passes of
//     at most kBatchMaxG), interleaved across the warp so that lane accesses are contiguous;
//   * a warp walks its chains in rounds: round r adds entry r of each chain to the chain's running affine sum (global
//     memory, L1 / L2 resident).  Forward sweep over the G chains: d_g = x(P_g) - x(sum_g), prefix products to a
//     per-thread scratch line; ONE inversion of the product per thread and round (Fp::inv_safegcd: division steps in
//     batches of 30, add / logic instructions mostly, so it fills issue slots the 4-cycle IMAD.WIDE leaves free);
//     backward sweep: 1 / d_g = running inverse * prefix_g, lambda, x3, y3;
//   * k_chain_combine adds the chain sums of a bucket (mixed additions) into the XYZZ bucket array the reduction reads.
//
// No block-level synchronisation anywhere.  Rare operand pairs (sum == identity after a cancellation, x(P) == x(sum):
// doubling or P + (-P)) are flagged in the forward sweep (d_g := 1) and finished through the XYZZ formulas with their
// own inversion in the backward sweep.  Giant buckets keep the chunked path of msm.cuh.
#pragma once
#include "common.cuh"
#include "curve.cuh"
#include "devutil.cuh"

namespace zkb {

constexpr int kBatchThreads = 64;            // threads per block
constexpr int kBatchMaxG = 40;               // chains per lane and pass (shared memory: 8 bytes each)
constexpr size_t kBatchSmem = (size_t)kBatchMaxG * 8 * kBatchThreads;

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// seg_cnt[b] = number of chains of bucket b: 0 for empty and for giant buckets
static __global__ void k_chain_count(const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t big, uint32_t lmax,
                                     uint32_t* __restrict__ seg_cnt) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t len = offsets[b + 1] - offsets[b];
  seg_cnt[b] = (len == 0 || len > big) ? 0u : (len + lmax - 1) / lmax;
}
// chain_bucket[c] = b for the chains c of bucket b (seg_off = exclusive scan of the counts, nb + 1 entries)
static __global__ void k_chain_build(const uint32_t* __restrict__ seg_off, uint32_t nb, uint32_t* __restrict__ chain_bucket) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  for (uint32_t c = seg_off[b]; c < seg_off[b + 1]; c++) chain_bucket[c] = b;
}

// F: the coordinate field with the multiplication as a call (CallVariant), same bytes as the table's field.
// chain_sum: running sums, one per chain; prefix: kBatchMaxG field elements per thread of the grid.
template <class F>
__global__ void __launch_bounds__(kBatchThreads)
k_accumulate_chains(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets,
                    const uint32_t* __restrict__ seg_off, const uint32_t* __restrict__ chain_bucket, uint32_t nb,
                    const Affine<F>* __restrict__ table, Affine<F>* __restrict__ chain_sum, F* __restrict__ prefix) {
  extern __shared__ uint32_t batch_smem[];
  uint32_t* start = batch_smem + threadIdx.x;                          // [g][thread]
  uint32_t* len = start + kBatchMaxG * kBatchThreads;

  const uint32_t n_chains = seg_off[nb];
  const uint32_t n_warps = gridDim.x * (kBatchThreads / 32);
  const uint32_t warp = (blockIdx.x * kBatchThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  // even deal: every warp gets `per_warp` chains (a multiple of 32) in `passes` passes of 32 * G
  const uint32_t per_lane = (n_chains + n_warps * 32 - 1) / (n_warps * 32);
  const uint32_t passes = (per_lane + kBatchMaxG - 1) / kBatchMaxG;
  if (per_lane == 0) return;
  const uint32_t G = (per_lane + passes - 1) / passes;
  F* pre = prefix + ((size_t)warp * kBatchMaxG) * 32 + lane;           // element g at pre[g * 32]

  for (uint32_t pass = 0; pass < passes; pass++) {
    const uint32_t c0 = (warp * passes + pass) * 32 * G + lane;        // chain g of this lane: c0 + 32 g
    if (c0 - lane >= n_chains) break;                                  // warp-uniform
    // ---- round 0: every running sum starts as the chain's first point
    uint32_t rounds = 0;
    for (uint32_t g = 0; g < G; g++) {
      const uint32_t c = c0 + 32 * g;
      uint32_t s = 0, l = 0;
      if (c < n_chains) {
        const uint32_t b = chain_bucket[c];
        const uint32_t so = seg_off[b], ns = seg_off[b + 1] - so, k = c - so;
        const uint32_t off = offsets[b], blen = offsets[b + 1] - off;
        s = off + (uint32_t)(((uint64_t)k * blen) / ns);
        l = off + (uint32_t)(((uint64_t)(k + 1) * blen) / ns) - s;
        const uint32_t e = entries[s];
        Affine<F> P = ld_vec(&table[e & 0x7fffffffu]);
        if (e >> 31) P.y = F::neg(P.y);
        st_vec(&chain_sum[c], P);
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
      uint64_t special = 0;                                            // bit g: finished by the XYZZ formulas
      for (uint32_t g = 0; g < G; g++) {
        if (r >= len[g * kBatchThreads]) continue;
        st_vec(&pre[g * 32], run);
        const uint32_t pos = start[g * kBatchThreads] + r;
        const uint32_t e = entries[pos];
        if (r + 1 < len[g * kBatchThreads]) {                          // the next round's point travels to L2 meanwhile
          const char* nx = reinterpret_cast<const char*>(&table[entries[pos + 1] & 0x7fffffffu]);
          prefetch_l2(nx);
          prefetch_l2(nx + sizeof(Affine<F>) - 1);
        }
        const F px = ld_vec(&table[e & 0x7fffffffu].x);
        const F ax = ld_vec_rw(&chain_sum[c0 + 32 * g].x);
        const F d = F::sub(px, ax);
        if (d.is_zero() || ax.is_zero()) special |= 1ull << g;
        else run = F::mul(run, d);
      }
      F inv = F::inv_fast(run);
      // ---- backward: finish the additions
      for (uint32_t g = G; g-- > 0;) {
        if (r >= len[g * kBatchThreads]) continue;
        const uint32_t c = c0 + 32 * g;
        const uint32_t e = entries[start[g * kBatchThreads] + r];
        Affine<F> P = ld_vec(&table[e & 0x7fffffffu]);
        if (e >> 31) P.y = F::neg(P.y);
        Affine<F> A = ld_vec_rw(&chain_sum[c]);
        if ((special >> g) & 1ull) {                                   // identity operand, doubling or cancellation
          XYZZ<F> t = XYZZ<F>::from_affine(A);
          pt_madd(t, P, false);
          pt_to_affine(A, t);
        } else {
          const F d = F::sub(P.x, A.x);
          const F inv_d = F::mul(inv, ld_vec_rw(&pre[g * 32]));
          inv = F::mul(inv, d);
          const F lam = F::mul(F::sub(P.y, A.y), inv_d);
          const F x3 = F::sub(F::sub(F::sqr(lam), A.x), P.x);
          A.y = F::sub(F::mul(lam, F::sub(A.x, x3)), A.y);
          A.x = x3;
        }
        st_vec(&chain_sum[c], A);
      }
    }
  }
}

// bucket_acc[b] = sum of the chain sums of bucket b (buckets without chains -- empty or giant -- are left alone)
template <class F>
__global__ void __launch_bounds__(128)
k_chain_combine(const uint32_t* __restrict__ seg_off, uint32_t nb, const Affine<F>* __restrict__ chain_sum,
                XYZZ<F>* __restrict__ bucket_acc) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t c0 = seg_off[b], c1 = seg_off[b + 1];
  if (c0 == c1) return;
  XYZZ<F> acc = XYZZ<F>::from_affine(ld_vec_rw(&chain_sum[c0]));
  for (uint32_t c = c0 + 1; c < c1; c++) {
    const Affine<F> p = ld_vec_rw(&chain_sum[c]);
    pt_madd(acc, p, false);
  }
  st_vec(&bucket_acc[b], acc);
}

}  // namespace zkb
