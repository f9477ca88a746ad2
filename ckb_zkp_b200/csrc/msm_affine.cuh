// Batched-affine pair levels of the bucket accumulation.
//
// Every bucket's sorted entry list is reduced as a binary tree: one level replaces the k points
// of a bucket by ceil(k / 2) points, adding neighbours (2j, 2j + 1) in affine coordinates and
// carrying an odd leftover over.  All additions of a level are independent, so their
// denominators (x1 - x0, or 2y for a doubling) share inversions by Montgomery's trick: a thread
// owns M output slots and multiplies its M denominators into a running product (stored per
// slot); the 128 products of a block are multiplied up a tree in shared memory, the root is
// inverted ONCE with the shift-and-subtract Euclid of field.cuh (add / logic pipe: it overlaps the
// multiplier work of the other resident blocks), the inverses travel back down the tree, and each
// thread walks back through its slots finishing every addition with 1/d = running inverse * prefix.
// Cost per addition: 5 multiplications + 1 squaring (+ 1/M of an inversion) instead of the
// 8 + 2 of the XYZZ mixed addition, and no bucket is "big": a bucket with a million entries is
// half a million independent pairs.  After the levels the (much shorter) lists go through the
// XYZZ accumulation of msm.cuh unchanged.
//
// Memory schedule of one level (what keeps the multiplier fed): a block owns a contiguous range of
// 128 * M slots per pass.  Phase 0 resolves every slot to its two source indices (8-byte record;
// contiguous slots per thread, so the bucket walk is cheap).  The forward and backward phases then
// visit the range TRANSPOSED -- iteration i of thread t is slot base + 128 i + t -- so records,
// prefix products and results stream through HBM coalesced, and the operands of the next slot(s)
// are copied into a per-thread shared-memory ring with cp.async while the current slot multiplies
// (no registers held, no scoreboard stall: the gathers of level 0 are ~1 us away).
//
// Work distribution (round 2): the slots of a level are cut into chunks of 128 * m_j slots whose lengths m_j cycle
// through a fixed pattern of DIFFERENT values, and the resident blocks draw chunks from an atomic counter.  Blocks of
// equal length started together reach their (single-thread, ~150 k cycle) inversion at the same moment and leave the
// multiplier pipe idle for a third of the time -- measured 41-61 % pipe utilisation in round 1 with equal passes on
// one resident wave; unequal chunks drift out of phase at once, so while one block of an SM inverts the others multiply.
// Dynamic chunks also make the kernel indifferent to how many of its blocks are resident when several MSMs share the
// machine (the five MSMs of a proof run on five streams).
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
constexpr uint32_t kPairNone = 0xffffffffu;   // second source of a slot that only carries one point over
// chunk j of a level holds 128 * scale * kPairUnits[j % 8] slots (64 units per period of 8 chunks)
constexpr int kPairPeriod = 8;
constexpr uint32_t kPairPeriodUnits = 64;
__host__ __device__ __forceinline__ uint32_t pair_chunk_units(uint32_t j) {
  // 5, 11, 7, 9, 6, 10, 8, 8 packed in nibbles (lowest nibble = j % 8 == 0)
  return (0x88A697B5u >> (4 * (j % kPairPeriod))) & 0xFu;
}
__host__ __device__ __forceinline__ uint32_t pair_chunk_start_units(uint32_t j) {
  // exclusive prefix sums of the pattern: 0, 5, 16, 23, 32, 38, 48, 56
  const uint32_t k = j % kPairPeriod;
  const uint32_t pre = k == 0 ? 0u : k == 1 ? 5u : k == 2 ? 16u : k == 3 ? 23u : k == 4 ? 32u : k == 5 ? 38u : k == 6 ? 48u : 56u;
  return (j / kPairPeriod) * kPairPeriodUnits + pre;
}

// cnt_out[b] = ceil(k_b / 2)
static __global__ void k_pair_sizes(const uint32_t* __restrict__ off_in, uint32_t nb, uint32_t* __restrict__ cnt_out) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb) cnt_out[b] = (off_in[b + 1] - off_in[b] + 1) >> 1;
}

// ---- cp.async (16-byte, L2 only) -----------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-thread operand ring in shared memory: 16-byte chunk c of stage s of thread t sits at
// ((s * CH + c) * 128 + t) * 16, so a warp's 128-bit accesses are conflict free.
template <class F>
struct PairRing {
  static constexpr int CHX = sizeof(F) / 16;            // chunks of one coordinate
  static constexpr int CHP = 2 * CHX;                   // chunks of one point
  static constexpr int CH_FULL = 2 * CHP;               // backward stage: two points
  static constexpr int CH_X = 2 * CHX;                  // forward stage: two x coordinates
  static constexpr int STAGES_FULL = 2;
  static constexpr int STAGES_X = 4;                    // same bytes: 4 * CH_X == 2 * CH_FULL
  static constexpr int kChunks = STAGES_FULL * CH_FULL;
  static constexpr size_t kBytes = (size_t)kChunks * kPairThreads * 16;      // >= 2 * kPairThreads * sizeof(F) (the tree)
  uint4* base;                                          // block's ring + threadIdx.x
  __device__ __forceinline__ uint4* chunk(int idx) const { return base + idx * kPairThreads; }
  __device__ __forceinline__ F load(int first_chunk) const {
    F r;
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int c = 0; c < CHX; c++) d[c] = *chunk(first_chunk + c);
    return r;
  }
};

// r = p + q in affine coordinates, 1/d taken from the shared inversion chain (see k_pair_level)
template <class F>
__device__ __forceinline__ void pair_finish(Affine<F>& r, const Affine<F>& q, F& inv_run, const F& prefix_prev, bool chain_more) {
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
    inv_d = F::mul(inv_run, prefix_prev);
    inv_run = F::mul(inv_run, d);
  }
  F lam = F::mul(num, inv_d);
  F x3 = F::sub(F::sub(F::sqr(lam), r.x), q.x);
  r.y = F::sub(F::mul(lam, F::sub(r.x, x3)), r.y);
  r.x = x3;
}

// denominator of P0 + P1 given their x; false when the sum needs no inversion (an identity operand, or
// P + (-P)).  The y coordinates are only read (from global memory) on the rare paths.
template <class F>
__device__ __forceinline__ bool pair_denominator(const Affine<F>* pts, uint2 rec, const F& x0, const F& x1, F& d) {
  const Affine<F>* a0 = pts + (rec.x & 0x7fffffffu);
  const Affine<F>* a1 = pts + (rec.y & 0x7fffffffu);
  if (x0.is_zero() && ld_vec_rw(&a0->y).is_zero()) return false;
  if (x1.is_zero() && ld_vec_rw(&a1->y).is_zero()) return false;
  d = F::sub(x1, x0);
  if (!d.is_zero()) return true;
  F y0 = ld_vec_rw(&a0->y), y1 = ld_vec_rw(&a1->y);
  if (rec.x >> 31) y0 = F::neg(y0);
  if (rec.y >> 31) y1 = F::neg(y1);
  if (y0 == y1 && !y0.is_zero()) { d = F::dbl(y0); return true; }
  return false;
}

// One level.  entries != nullptr (level 0): the input lists are the sorted (negate | table index) entries and
// pts is the base table; else the lists are the previous level's dense points.  off_in / off_out: bucket
// offsets of the input / output lists (nb + 1 each).  The grid is (at most) one resident wave of blocks that draw
// chunks of 128 * M slots (M = scale * pattern[j % 8]) from *work_counter (zeroed by the host) until the E output
// slots of the level are used up.
template <class F>
__global__ void __launch_bounds__(kPairThreads, sizeof(F) <= 48 ? 3 : 2)
k_pair_level(const uint32_t* __restrict__ entries, const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ off_in,
             const uint32_t* __restrict__ off_out, uint32_t nb, uint2* __restrict__ recs, F* __restrict__ prefix,
             Affine<F>* __restrict__ out, uint32_t* __restrict__ work_counter, uint32_t scale) {
  using Ring = PairRing<F>;
  // dynamic shared memory (PairRing<F>::kBytes): the operand ring; between the forward and the backward phase,
  // while the ring is idle, its first bytes hold the inversion tree (node i: children 2i, 2i + 1; leaves at
  // kPairThreads + tid)
  extern __shared__ uint4 pair_smem[];
  __shared__ uint32_t chunk_sh;
  F* tree = reinterpret_cast<F*>(pair_smem);
  const uint32_t tid = threadIdx.x;
  Ring ring{pair_smem + tid};
  const uint32_t E = off_out[nb];
  for (;;) {                                      // block-uniform
    if (tid == 0) chunk_sh = atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t chunk = chunk_sh;
    const uint64_t base64 = (uint64_t)kPairThreads * scale * pair_chunk_start_units(chunk);
    if (base64 >= E) break;
    const uint32_t M = scale * pair_chunk_units(chunk);
    const uint32_t base = (uint32_t)base64;
    const uint32_t range_end = E - base < kPairThreads * M ? E : base + kPairThreads * M;

    // ---- phase 0: slot -> (source 0, source 1) records; thread t resolves slots [base + t M, base + (t + 1) M)
    {
      const uint32_t o0 = base + tid * M < range_end ? base + tid * M : range_end;
      const uint32_t o1 = o0 + M < range_end ? o0 + M : range_end;
      if (o0 < o1) {
        uint32_t lo = 0, hi = nb;                 // off_out[lo] <= o0 < off_out[hi]
        while (hi - lo > 1) {
          uint32_t mid = (lo + hi) >> 1;
          if (off_out[mid] <= o0) lo = mid; else hi = mid;
        }
        uint32_t b = lo, b_beg = off_out[b], b_end = off_out[b + 1], in_beg = off_in[b], in_end = off_in[b + 1];
        for (uint32_t o = o0; o < o1; o++) {
          while (o >= b_end) {                    // next non-empty bucket
            b++;
            b_beg = b_end;
            b_end = off_out[b + 1];
            in_beg = off_in[b];
            in_end = off_in[b + 1];
          }
          const uint32_t s0 = in_beg + 2 * (o - b_beg);
          uint2 rec;
          rec.x = entries ? __ldg(entries + s0) : s0;
          rec.y = s0 + 1 < in_end ? (entries ? __ldg(entries + s0 + 1) : s0 + 1) : kPairNone;
          recs[o] = rec;
        }
      }
    }
    __syncthreads();                              // records are read back transposed below

    // slots of this thread: o(i) = base + tid + 128 i, i < cnt
    const uint32_t cnt = base + tid < range_end ? (range_end - base - tid + kPairThreads - 1) / kPairThreads : 0;
    auto slot = [&](uint32_t i) { return base + tid + i * kPairThreads; };

    // ---- forward: running product of the denominators; x coordinates staged STAGES_X - 1 slots ahead
    F run = F::one();
    {
      auto stage_x = [&](uint32_t i) {            // copy the two x of slot i into ring stage i % STAGES_X
        if (i < cnt) {
          const uint2 rec = recs[slot(i)];
          if (rec.y != kPairNone) {
            const int s = (int)(i % Ring::STAGES_X) * Ring::CH_X;
            const uint4* g0 = reinterpret_cast<const uint4*>(&pts[rec.x & 0x7fffffffu].x);
            const uint4* g1 = reinterpret_cast<const uint4*>(&pts[rec.y & 0x7fffffffu].x);
#pragma unroll
            for (int c = 0; c < Ring::CHX; c++) {
              cp_async16(ring.chunk(s + c), g0 + c);
              cp_async16(ring.chunk(s + Ring::CHX + c), g1 + c);
            }
          }
        }
        cp_async_commit();                        // one group per slot, empty or not: wait counts stay uniform
      };
#pragma unroll
      for (int k = 0; k < Ring::STAGES_X - 1; k++) stage_x(k);
      for (uint32_t i = 0; i < cnt; i++) {
        stage_x(i + Ring::STAGES_X - 1);
        cp_async_wait<Ring::STAGES_X - 1>();      // slot i has landed
        const uint2 rec = recs[slot(i)];
        if (rec.y != kPairNone) {
          const int s = (int)(i % Ring::STAGES_X) * Ring::CH_X;
          const F x0 = ring.load(s), x1 = ring.load(s + Ring::CHX);
          F d;
          if (pair_denominator(pts, rec, x0, x1, d)) run = F::mul(run, d);
        }
        st_vec(&prefix[slot(i)], run);
      }
      cp_async_wait<0>();
    }

    // ---- one inversion per block: product tree up, inverse, inverses down
    __syncthreads();                              // all threads are done with the ring: the tree takes its place
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
    __syncthreads();                              // every leaf is read before the ring overwrites the tree

    // ---- backward: finish the additions; both points of the next slot staged while this one computes
    {
      auto stage_full = [&](uint32_t i) {         // i counts down; i >= cnt (wrapped below 0) means nothing left
        if (i < cnt) {
          const int s = (int)(i & 1) * Ring::CH_FULL;
          const uint2 rec = recs[slot(i)];
          const uint4* g0 = reinterpret_cast<const uint4*>(&pts[rec.x & 0x7fffffffu]);
#pragma unroll
          for (int c = 0; c < Ring::CHP; c++) cp_async16(ring.chunk(s + c), g0 + c);
          if (rec.y != kPairNone) {
            const uint4* g1 = reinterpret_cast<const uint4*>(&pts[rec.y & 0x7fffffffu]);
#pragma unroll
            for (int c = 0; c < Ring::CHP; c++) cp_async16(ring.chunk(s + Ring::CHP + c), g1 + c);
          }
        }
        cp_async_commit();
      };
      if (cnt) stage_full(cnt - 1);
      F pre = cnt > 1 ? ld_vec_rw(&prefix[slot(cnt - 2)]) : F::one();
      for (uint32_t i = cnt; i-- > 0;) {
        stage_full(i - 1);                        // i == 0 wraps to 0xffffffff >= cnt: empty group
        const F pre_next = i > 1 ? ld_vec_rw(&prefix[slot(i - 2)]) : F::one();
        cp_async_wait<1>();                       // slot i has landed
        const uint2 rec = recs[slot(i)];
        const int s = (int)(i & 1) * Ring::CH_FULL;
        Affine<F> r{ring.load(s), ring.load(s + Ring::CHX)};
        if (rec.x >> 31) r.y = F::neg(r.y);
        if (rec.y != kPairNone) {
          Affine<F> q{ring.load(s + Ring::CHP), ring.load(s + Ring::CHP + Ring::CHX)};
          if (rec.y >> 31) q.y = F::neg(q.y);
          pair_finish(r, q, inv_run, pre, i > 0);
        }
        st_vec(&out[slot(i)], r);
        pre = pre_next;
      }
      cp_async_wait<0>();
    }
    __syncthreads();                              // tree, ring, records and chunk_sh are reused by the next chunk
  }
}

}  // namespace zkb
