// Bucket accumulation in AFFINE coordinates with batched inversion (Montgomery's trick): one inversion per thread and
// round, over bucket lists cut into short chains of equal length, in two levels.
//
// The XYZZ loop of msm.cuh (k_accumulate) spends 10 multiplications per bucket entry and runs at ~85 % of the
// IMAD.WIDE pipe, so only fewer multiplications make the MSM faster.  An affine addition costs 1 inversion + 3
// multiplications; sharing the inversion between k independent additions costs 3 more multiplications each, i.e.
// 6 per entry once the inversion is amortised over enough additions.  Additions into the SAME running sum depend on
// each other, additions into different sums do not, hence:
//
//   * level 0 cuts every regular bucket list (<= the big-bucket threshold) into ceil(len / L) chains of equal length
//     (+-1) -- ~2 M chains of <= 8 entries for a 2^20-point MSM -- and sums each chain; level 1 does the same with the
//     chain sums of a bucket as its list (3-4 of them on average, so one chain per bucket); what is left after two
//     levels (only lists above L^2 entries) is added by k_chain_combine with XYZZ mixed additions, which also converts
//     to the bucket array the reduction reads.  Chains, not whole lists, because a lock-step schedule needs equal
//     lengths (whole lists: Poisson lengths, mean 26, max ~55 -- version 1 ran one half-empty wave, 13.6 ms against
//     4.9 ms for the XYZZ loop) and because the number of independent chains per lane, G, is what amortises the
//     inversion (version 2, one level with L = 12: G = 34, the division-step inversion still took 30 % of the
//     multiplier cycles, 5.5 + 0.5 ms);
//   * the grid is one resident wave; the chains of a level are dealt out evenly, G = chains / (32 * warps) per lane (in
//     passes of at most g_max), interleaved across the warp so that lane accesses are contiguous;
//   * a warp walks its chains in rounds: round r adds item r of each chain to the chain's running affine sum (global
//     memory).  Forward sweep over the G chains: d_g = x(P_g) - x(sum_g), prefix products to a per-lane scratch line;
//     ONE inversion of the product per lane and round (Fp::inv_safegcd); backward sweep: 1 / d_g = running inverse *
//     prefix_g, lambda, x3, y3.  The operands of the next chain are loaded before the multiplications of the current
//     one (the multiplication is an opaque call: the compiler does not move loads across it on its own).
//
// No block-level synchronisation anywhere.  Rare operand pairs (sum or item == identity after a cancellation,
// x(P) == x(sum): doubling or P + (-P)) are flagged in the forward sweep (d_g := 1) and finished through the XYZZ
// formulas with their own inversion in the backward sweep.  Giant buckets keep the chunked path of msm.cuh.
#pragma once
#include "common.cuh"
#include "curve.cuh"
#include "devutil.cuh"

namespace zkb {

constexpr int kBatchThreads = 64;            // threads per block
constexpr uint32_t kBatchGMax = 64;          // chains per lane and pass (prefix scratch: one field element each; mask bits)

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// seg_cnt[b] = number of chains of list b: 0 for empty lists and (level 0) for giant buckets
static __global__ void k_chain_count(const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t big, uint32_t lmax,
                                     uint32_t* __restrict__ seg_cnt) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t len = offsets[b + 1] - offsets[b];
  seg_cnt[b] = (len == 0 || len > big) ? 0u : (len + lmax - 1) / lmax;
}
// rec[c] = (first item, length) of chain c: list b is cut into seg_off[b + 1] - seg_off[b] parts of equal length (+-1)
static __global__ void k_chain_build(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ seg_off, uint32_t nb,
                                     uint2* __restrict__ rec) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t c0 = seg_off[b], ns = seg_off[b + 1] - c0;
  if (ns == 0) return;
  const uint32_t off = offsets[b], len = offsets[b + 1] - off;
  uint32_t s = off;
  for (uint32_t k = 0; k < ns; k++) {
    const uint32_t e = off + (uint32_t)(((uint64_t)(k + 1) * len) / ns);
    rec[c0 + k] = make_uint2(s, e - s);
    s = e;
  }
}

// One level.  GATHER: item i is (negate | table index) entries[i] into the base table `pts`; else item i is the dense
// point pts[i] (a chain sum of the level below; may be the identity).  F: the coordinate field with the multiplication
// as a call (CallVariant), same bytes as the table's field.  chain_sum: one running sum per chain; prefix: g_max
// field elements per lane of the grid.
template <class F, bool GATHER>
__global__ void __launch_bounds__(kBatchThreads)
k_accumulate_chains(const uint32_t* __restrict__ entries, const Affine<F>* __restrict__ pts, const uint2* __restrict__ rec,
                    const uint32_t* __restrict__ n_chains_ptr, Affine<F>* __restrict__ chain_sum, F* __restrict__ prefix,
                    uint32_t g_max) {
  const uint32_t n_chains = *n_chains_ptr;
  const uint32_t n_warps = gridDim.x * (kBatchThreads / 32);
  const uint32_t warp = (blockIdx.x * kBatchThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  // even deal: every lane gets `per_lane` chains in `passes` passes of G
  const uint32_t per_lane = (n_chains + n_warps * 32 - 1) / (n_warps * 32);
  if (per_lane == 0) return;
  const uint32_t passes = (per_lane + g_max - 1) / g_max;
  const uint32_t G = (per_lane + passes - 1) / passes;
  F* pre = prefix + ((size_t)warp * g_max) * 32 + lane;                // element g at pre[g * 32]

  auto item_index = [&](uint32_t i) -> uint32_t { return GATHER ? (entries[i] & 0x7fffffffu) : i; };
  auto chain_rec = [&](uint32_t c) -> uint2 { return c < n_chains ? rec[c] : make_uint2(0u, 0u); };

  for (uint32_t pass = 0; pass < passes; pass++) {
    const uint32_t c0 = (warp * passes + pass) * 32 * G + lane;        // chain g of this lane: c0 + 32 g
    if (c0 - lane >= n_chains) break;                                  // warp-uniform
    // ---- round 0: every running sum starts as the chain's first item
    uint32_t rounds = 0;
    for (uint32_t g = 0; g < G; g++) {
      const uint32_t c = c0 + 32 * g;
      const uint2 rc = chain_rec(c);
      if (rc.y) {
        Affine<F> P;
        if (GATHER) {
          const uint32_t e = entries[rc.x];
          P = ld_vec(&pts[e & 0x7fffffffu]);
          if (e >> 31) P.y = F::neg(P.y);
        } else {
          P = ld_vec_rw(&pts[rc.x]);
        }
        st_vec(&chain_sum[c], P);
        if (GATHER && rc.y > 1) {
          const char* nx = reinterpret_cast<const char*>(&pts[item_index(rc.x + 1)]);
          prefetch_l2(nx);
          prefetch_l2(nx + sizeof(Affine<F>) - 1);
        }
      }
      rounds = rc.y > rounds ? rc.y : rounds;
    }
    rounds = __reduce_max_sync(0xffffffffu, rounds);

    for (uint32_t r = 1; r < rounds; r++) {
      // ---- forward: denominators and their prefix products; operands of chain g + 1 load while chain g multiplies
      F run = F::one();
      uint64_t special = 0;                                            // bit g: finished by the XYZZ formulas
      bool act = false;
      F px, ax;
      auto load_fwd = [&](uint32_t g, bool& a, F& x_item, F& x_sum) {
        const uint32_t c = c0 + 32 * g;
        const uint2 rc = chain_rec(c);
        a = r < rc.y;
        if (a) {
          const uint32_t idx = item_index(rc.x + r);
          if (GATHER && r + 1 < rc.y) {                                // the next round's point travels to L2 meanwhile
            const char* nx = reinterpret_cast<const char*>(&pts[item_index(rc.x + r + 1)]);
            prefetch_l2(nx);
            prefetch_l2(nx + sizeof(Affine<F>) - 1);
          }
          x_item = GATHER ? ld_vec(&pts[idx].x) : ld_vec_rw(&pts[idx].x);
          x_sum = ld_vec_rw(&chain_sum[c].x);
        }
      };
      load_fwd(0, act, px, ax);
      for (uint32_t g = 0; g < G; g++) {
        bool act_n = false;
        F px_n, ax_n;
        if (g + 1 < G) load_fwd(g + 1, act_n, px_n, ax_n);
        if (act) {
          st_vec(&pre[g * 32], run);
          const F d = F::sub(px, ax);
          if (d.is_zero() || ax.is_zero() || (!GATHER && px.is_zero())) special |= 1ull << g;
          else run = F::mul(run, d);
        }
        act = act_n;
        if (act_n) { px = px_n; ax = ax_n; }
      }
      F inv = F::inv_fast(run);
      // ---- backward: finish the additions; operands of chain g - 1 load while chain g multiplies
      Affine<F> P, A;
      F pf;
      bool neg = false;
      auto load_bwd = [&](uint32_t g, bool& a, Affine<F>& item, Affine<F>& sum, F& pfx, bool& ng) {
        const uint32_t c = c0 + 32 * g;
        const uint2 rc = chain_rec(c);
        a = r < rc.y;
        if (a) {
          if (GATHER) {
            const uint32_t e = entries[rc.x + r];
            item = ld_vec(&pts[e & 0x7fffffffu]);
            ng = (e >> 31) != 0;
          } else {
            item = ld_vec_rw(&pts[rc.x + r]);
            ng = false;
          }
          sum = ld_vec_rw(&chain_sum[c]);
          pfx = ld_vec_rw(&pre[g * 32]);
        }
      };
      // (Fq2 points are 48 registers each: there the look-ahead would spill, and 17 base-field multiplications per
      // addition hide the load latency better anyway)
      constexpr bool kLookAhead = sizeof(F) <= 48;
      if (kLookAhead) load_bwd(G - 1, act, P, A, pf, neg);
      for (uint32_t g = G; g-- > 0;) {
        bool act_n = false, neg_n = false;
        Affine<F> P_n, A_n;
        F pf_n;
        if (kLookAhead) {
          if (g > 0) load_bwd(g - 1, act_n, P_n, A_n, pf_n, neg_n);
        } else {
          load_bwd(g, act, P, A, pf, neg);
        }
        if (act) {
          if (neg) P.y = F::neg(P.y);
          if ((special >> g) & 1ull) {                                 // identity operand, doubling or cancellation
            XYZZ<F> t = XYZZ<F>::from_affine(A);
            pt_madd(t, P, false);
            pt_to_affine(A, t);
          } else {
            const F d = F::sub(P.x, A.x);
            const F inv_d = F::mul(inv, pf);
            inv = F::mul(inv, d);
            const F lam = F::mul(F::sub(P.y, A.y), inv_d);
            const F x3 = F::sub(F::sub(F::sqr(lam), A.x), P.x);
            A.y = F::sub(F::mul(lam, F::sub(A.x, x3)), A.y);
            A.x = x3;
          }
          st_vec(&chain_sum[c0 + 32 * g], A);
        }
        if (kLookAhead) {
          act = act_n;
          if (act_n) { P = P_n; A = A_n; pf = pf_n; neg = neg_n; }
        }
      }
    }
  }
}

// bucket_acc[b] = sum of the (few) remaining chain sums of bucket b; buckets without chains -- empty or giant -- are
// left alone
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
