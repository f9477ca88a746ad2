// Batched-affine pair levels of the bucket accumulation.
//
// Every bucket's sorted entry list is reduced as a binary tree: one level replaces the k points
// of a bucket by ceil(k / 2) points, adding neighbours (2j, 2j + 1) in affine coordinates and
// carrying an odd leftover over.  All additions of a level are independent, so their
// denominators (x1 - x0, or 2y for a doubling) share inversions by Montgomery's trick: a thread
// owns M consecutive output slots and multiplies its M denominators into a running product (stored
// per slot); the 128 products of a block are multiplied up a tree in shared memory, the root is
// inverted ONCE with the shift-and-subtract Euclid of field.cuh (add / logic pipe: it overlaps the
// multiplier work of the other resident blocks), the inverses travel back down the tree, and each
// thread walks back through its slots finishing every addition with 1/d = running inverse * prefix.
// Cost per addition: 5 multiplications + 1 squaring (+ 1/M of an inversion) instead of the
// 8 + 2 of the XYZZ mixed addition, and no bucket is "big": a bucket with a million entries is
// half a million independent pairs.  After the levels the (much shorter) lists go through the
// XYZZ accumulation of msm.cuh unchanged.
//
// A level is: k_pair_sizes -> exclusive scan (msm.cuh) -> k_pair_level.
// Identity = affine (0, 0), as everywhere on the device; P + (-P), doublings and identities are
// handled (they do not occur for honest inputs, but bases are not required to be distinct).
#pragma once
#include "common.cuh"
#include "curve.cuh"
#include "devutil.cuh"

namespace zkb {

constexpr int kPairThreads = 128;             // threads per block = leaves of the shared inversion tree
constexpr uint32_t kPairMinM = 4;             // output slots per thread and pass: lower / upper limit
constexpr uint32_t kPairMaxM = 128;
constexpr uint32_t kPairFlag = 0x80000000u;   // meta: the slot adds two points (else it copies one)

// cnt_out[b] = ceil(k_b / 2)
static __global__ void k_pair_sizes(const uint32_t* __restrict__ off_in, uint32_t nb, uint32_t* __restrict__ cnt_out) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb) cnt_out[b] = (off_in[b + 1] - off_in[b] + 1) >> 1;
}

// pull [p, p + bytes) towards L1 (no destination register: the data is loaded normally one iteration later)
static __device__ int g_pair_prefetch = 0;             // 0 none, 1 L1, 2 L2 (tuning switch, set from ZKB_PAIR_PF)
__device__ __forceinline__ void prefetch_l1(const void* p, uint32_t bytes) {
  const char* c = (const char*)p;
  const int mode = g_pair_prefetch;
  if (mode == 1) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c));
    if ((((uintptr_t)c) & 127u) + bytes > 128u) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + bytes - 1));
  } else if (mode == 2) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c));
    if ((((uintptr_t)c) & 127u) + bytes > 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + bytes - 1));
  }
}

// where the points of a level come from: level 0 gathers them from the base table through the sorted
// (negate | index) entries, later levels read the previous level's dense output
template <class F, bool GATHER>
struct PairSrc {
  const uint32_t* entries;
  const Affine<F>* pts;
  __device__ __forceinline__ const Affine<F>* at(uint32_t pos, bool& neg) const {
    if (GATHER) {
      uint32_t e = __ldg(entries + pos);
      neg = (e >> 31) != 0;
      return pts + (e & 0x7fffffffu);
    }
    neg = false;
    return pts + pos;
  }
  __device__ __forceinline__ F coord(const F* p) const { return GATHER ? ld_vec(p) : ld_vec_rw(p); }
  __device__ __forceinline__ Affine<F> point(uint32_t pos) const {
    bool neg;
    const Affine<F>* a = at(pos, neg);
    Affine<F> p = GATHER ? ld_vec(a) : ld_vec_rw(a);
    if (neg) p.y = F::neg(p.y);
    return p;
  }
  __device__ __forceinline__ void prefetch_x(uint32_t pos) const {
    bool neg;
    prefetch_l1(&at(pos, neg)->x, sizeof(F));
  }
  __device__ __forceinline__ void prefetch(uint32_t pos) const {
    bool neg;
    prefetch_l1(at(pos, neg), sizeof(Affine<F>));
  }
  // x coordinates of points[pos], points[pos + 1] (what the denominator needs in the common case)
  __device__ __forceinline__ void fetch_x(uint32_t pos, F& x0, F& x1) const {
    bool n0, n1;
    x0 = coord(&at(pos, n0)->x);
    x1 = coord(&at(pos + 1, n1)->x);
  }
  // denominator of points[pos] + points[pos + 1] given their x; false when the sum needs no inversion
  // (an identity operand, or P + (-P)).  The y coordinates are only read on the rare paths.
  __device__ __forceinline__ bool denominator(uint32_t pos, const F& x0, const F& x1, F& d) const {
    bool n0, n1;
    if (x0.is_zero() && coord(&at(pos, n0)->y).is_zero()) return false;
    if (x1.is_zero() && coord(&at(pos + 1, n1)->y).is_zero()) return false;
    d = F::sub(x1, x0);
    if (!d.is_zero()) return true;
    F y0 = coord(&at(pos, n0)->y), y1 = coord(&at(pos + 1, n1)->y);
    if (n0) y0 = F::neg(y0);
    if (n1) y1 = F::neg(y1);
    if (y0 == y1 && !y0.is_zero()) { d = F::dbl(y0); return true; }
    return false;
  }
};

// r = p + q in affine coordinates, 1/d taken from the shared inversion chain (see k_pair_level)
template <class F>
__device__ __forceinline__ void pair_finish(Affine<F>& r, const Affine<F>& q, F& inv_run, const F* prefix_prev, bool chain_more) {
  if (r.is_inf()) { r = q; return; }
  if (q.is_inf()) return;
  F d = F::sub(q.x, r.x), num;
  if (d.is_zero()) {
    if (r.y == q.y && !r.y.is_zero()) {     // doubling: lambda = 3 x^2 / 2 y
      d = F::dbl(r.y);
      F xx = F::sqr(r.x);
      num = F::add(F::dbl(xx), xx);
    } else {                                // P + (-P)
      r = Affine<F>::inf();
      return;
    }
  } else {
    num = F::sub(q.y, r.y);
  }
  F inv_d = inv_run;
  if (chain_more) {
    inv_d = F::mul(inv_run, *prefix_prev);
    inv_run = F::mul(inv_run, d);
  }
  F lam = F::mul(num, inv_d);
  F x3 = F::sub(F::sub(F::sqr(lam), r.x), q.x);
  r.y = F::sub(F::mul(lam, F::sub(r.x, x3)), r.y);
  r.x = x3;
}

// One level.  off_in / off_out: bucket offsets of the input / output lists (nb + 1 each).  The grid is
// sized to one resident wave; every thread takes M = ceil(E / threads) consecutive output slots (E is only
// known on the device), in several passes when that exceeds kPairMaxM.  Per pass and block: forward
// products per thread -> product tree over the block's 128 threads in shared memory -> ONE inversion
// (thread 0) -> inverses pushed back down the tree -> backward pass finishing the additions.
template <class F, bool GATHER>
__global__ void __launch_bounds__(kPairThreads, sizeof(F) <= 48 ? 3 : 2)
k_pair_level(PairSrc<F, GATHER> src, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out,
             uint32_t nb, uint32_t* __restrict__ meta, F* __restrict__ prefix, Affine<F>* __restrict__ out) {
  __shared__ F tree[2 * kPairThreads];            // node i: children 2i, 2i + 1; leaves at kPairThreads + tid
  const uint32_t tid = threadIdx.x;
  const uint32_t E = off_out[nb];
  const uint32_t T = gridDim.x * kPairThreads;
  uint32_t M = (E + T - 1) / T;
  if (M > kPairMaxM) {                            // several passes of equal length
    const uint32_t passes = (M + kPairMaxM - 1) / kPairMaxM;
    M = (M + passes - 1) / passes;
  }
  if (M < kPairMinM) M = kPairMinM;
  const uint64_t per_pass = (uint64_t)T * M;
  for (uint64_t base = (uint64_t)blockIdx.x * kPairThreads * M; base < E; base += per_pass) {   // block-uniform
    const uint64_t first = base + (uint64_t)tid * M;
    const uint32_t o0 = (uint32_t)(first < E ? first : E);
    const uint32_t cnt = E - o0 < M ? E - o0 : M;

    // ---- forward: slot -> source position, running product of the denominators
    F run = F::one();
    if (cnt) {
      uint32_t lo = 0, hi = nb;                   // off_out[lo] <= o0 < off_out[hi]
      while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (off_out[mid] <= o0) lo = mid; else hi = mid;
      }
      uint32_t b = lo, b_beg = off_out[b], b_end = off_out[b + 1], in_beg = off_in[b], in_end = off_in[b + 1];
      auto locate = [&](uint32_t o) -> uint32_t {   // meta word of slot o (slots are visited in order)
        while (o >= b_end) {                      // next non-empty bucket
          b++;
          b_beg = b_end;
          b_end = off_out[b + 1];
          in_beg = off_in[b];
          in_end = off_in[b + 1];
        }
        const uint32_t s0 = in_beg + 2 * (o - b_beg);
        return s0 | (s0 + 1 < in_end ? kPairFlag : 0u);
      };
      uint32_t m = locate(o0);
      for (uint32_t i = 0; i < cnt; i++) {
        const uint32_t o = o0 + i;
        uint32_t m_next = 0;
        if (i + 1 < cnt) {                        // pull the next slot's operands towards L1 while this one multiplies
          m_next = locate(o + 1);
        }
        meta[o] = m;
        if (m & kPairFlag) {
          F x0, x1, d;
          src.fetch_x(m & ~kPairFlag, x0, x1);
          if (src.denominator(m & ~kPairFlag, x0, x1, d)) run = F::mul(run, d);
        }
        st_vec(&prefix[o], run);
        m = m_next;
      }
    }

    // ---- one inversion per block: product tree up, inverse, inverses down
    tree[kPairThreads + tid] = run;
    __syncthreads();
    for (uint32_t s = kPairThreads / 2; s >= 1; s >>= 1) {
      if (tid < s) tree[s + tid] = F::mul(tree[2 * (s + tid)], tree[2 * (s + tid) + 1]);
      __syncthreads();
    }
    if (tid == 0) tree[1] = F::inv(tree[1]);
    __syncthreads();
    for (uint32_t s = 1; s < kPairThreads; s <<= 1) {
      if (tid < s) {
        const uint32_t i = s + tid;
        const F iv = tree[i], l = tree[2 * i], r = tree[2 * i + 1];
        tree[2 * i] = F::mul(iv, r);
        tree[2 * i + 1] = F::mul(iv, l);
      }
      __syncthreads();
    }
    F inv_run = tree[kPairThreads + tid];

    // ---- backward: finish the additions.  The operands of slot i - 1 (two gathered points and the prefix
    // product) are pulled towards L1 while slot i computes; meta and entries are walked contiguously,
    // so reading them early to form the addresses costs L1 hits only.
    if (cnt) {
      uint32_t i = cnt;
      while (i-- > 0) {
        const uint32_t o = o0 + i;
        const uint32_t m = meta[o];
        Affine<F> r = src.point(m & ~kPairFlag);
        if (m & kPairFlag) {
          Affine<F> q = src.point((m & ~kPairFlag) + 1);
          F pre = i > 0 ? ld_vec_rw(&prefix[o - 1]) : F::one();
          pair_finish(r, q, inv_run, &pre, i > 0);
        }
        st_vec(&out[o], r);
      }
    }
    __syncthreads();                              // the tree is reused by the next pass
  }
}

}  // namespace zkb
