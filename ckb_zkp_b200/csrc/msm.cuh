// Variable-base multi-scalar multiplication on the device.
//
// Replaces ark_ec::msm::VariableBaseMSM::multi_scalar_mul (ark-ec 0.2, un-vendored; call
// sites groth16/src/prover.rs:187,190,220, marlin/src/pc/kzg10.rs:109,118,137,146,
// curve/src/lib.rs:44).  Same mathematical function sum_i s_i * P_i; the schedule is
// re-designed for B200:
//
//   * signed c-bit digits (half the buckets of the reference's unsigned windows);
//   * with ZKB_SRS_PRECOMPUTE the bases 2^(c*j) * P_i of every window j are resident in
//     HBM (180 GB makes a 13x copy of the SRS cheap), so ALL windows share ONE bucket
//     array: no per-window reduction and no doubling chain at the end;
//   * counting sort of (bucket, base index) pairs: histogram -> exclusive scan -> scatter;
//   * bucket accumulation: one thread per bucket, buckets ordered by decreasing size so the
//     32 lanes of a warp run the same trip count; buckets above a runtime threshold
//     (boolean-heavy witnesses put a large share of all entries into the "digit = 1" bucket)
//     are cut into 2048-entry chunks, one block each, and folded by a second block-level pass;
//   * bucket reduction sum_b (b + 1) * S_b through the row / column sums of the bucket array
//     viewed as a matrix [hi][lo]:  sum_hi (hi * L) * R_hi + sum_lo (lo + 1) * C_lo.
//
// Entry layout (uint32): bit 31 = negate, bits 0..30 = index into the base table.
#pragma once
#include <functional>
#include <memory>

#include "common.cuh"
#include "curve.cuh"
#include "devutil.cuh"

struct zkb_srs {
  zkb_ctx* ctx;
  int curve, group;
  size_t n;
  int c, W;          // window bits, number of windows
  int precomp;       // table holds W * n points (window-major) when set
  void* table;       // Affine<F>[(precomp ? W : 1) * n]
  uint8_t* inf;      // n flags
  // multi-GPU: this SRS holds bases [global_lo, global_lo + n) of a logical SRS of global_n bases
  // (zkb_srs_upload_shard); an unsharded SRS has global_lo = 0, global_n = n
  size_t global_lo, global_n;
};

namespace zkb {

constexpr int kSizeBins = 4096;         // regular buckets hold 1 .. big_threshold <= kSizeBins entries
constexpr int kMinBigBucket = 256;      // lower clamp of the runtime big-bucket threshold
constexpr int kChunkThreads = 128;      // threads per big-bucket chunk block
constexpr int kChunkPer = 32;           // entries per thread in a chunk
constexpr int kChunk = kChunkThreads * kChunkPer;
constexpr int kSegK = 8;                // points summed per thread in the row / column passes
constexpr int kScalarLimbs = 8;

// ------------------------------------------------------------------------------------------
// window geometry
// ------------------------------------------------------------------------------------------
struct MsmGeom {
  int c, W;
  uint32_t B;          // buckets per window = 2^(c-1)
  int precomp;
  uint32_t n_srs;      // stride between windows in the precomputed table
  uint32_t n_sets;     // bucket sets: 1 (precomp) or W
};

inline int msm_windows(int scalar_bits, int c) { return (scalar_bits + 1 + c - 1) / c; }

// ------------------------------------------------------------------------------------------
// digit extraction (histogram pass and scatter pass)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t window_bits(const uint32_t* s, int bit, int c) {
  int limb = bit >> 5, sh = bit & 31;
  if (limb >= kScalarLimbs) return 0;
  uint64_t w = s[limb];
  if (limb + 1 < kScalarLimbs) w |= (uint64_t)s[limb + 1] << 32;
  return (uint32_t)(w >> sh) & ((1u << c) - 1u);
}

template <bool SCATTER, class FrField>
__global__ void k_digits(const uint32_t* __restrict__ scalars, uint32_t n, const uint8_t* __restrict__ inf,
                         MsmGeom g, uint32_t base_offset, int scalars_mont,
                         uint32_t* __restrict__ counters, uint32_t* __restrict__ entries) {
  const uint32_t half = 1u << (g.c - 1);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (inf[base_offset + i]) continue;
    FrField sc = ld_vec(reinterpret_cast<const FrField*>(scalars) + i);
    if (scalars_mont) sc = FrField::from_mont(sc);       // fused into_repr (curve/src/lib.rs:39-42)
    if (sc.is_zero()) continue;
    uint32_t carry = 0;
    for (int j = 0; j < g.W; j++) {
      uint32_t raw = window_bits(sc.v, j * g.c, g.c) + carry;
      uint32_t mag, neg;
      if (raw > half) { mag = (1u << g.c) - raw; neg = 1; carry = 1; }
      else { mag = raw; neg = 0; carry = 0; }
      if (mag == 0) continue;
      uint32_t bucket = (g.precomp ? 0u : (uint32_t)j * g.B) + (mag - 1);
      if (SCATTER) {
        uint32_t pos = atomicAdd(&counters[bucket], 1u);
        uint32_t idx = (g.precomp ? (uint32_t)j * g.n_srs : 0u) + base_offset + i;
        entries[pos] = idx | (neg << 31);
      } else {
        atomicAdd(&counters[bucket], 1u);
      }
    }
  }
}

// Tiny MSMs over a precomputed table (Marlin's hiding commitments are 2-point MSMs over a 1.5 M-point key): one
// thread per (pair, window) multiplies the window's table point by its signed digit (<= c bits of double-and-add);
// the n * W terms are then summed by k_sum_points.  No bucket array, no sort, no reduction passes.
constexpr uint32_t kSmallMsmTerms = 512;
template <class F, class FrField>
__global__ void __launch_bounds__(128)
k_small_msm_terms(const uint32_t* __restrict__ scalars, uint32_t n, const uint8_t* __restrict__ inf, MsmGeom g,
                  uint32_t base_offset, int scalars_mont, const Affine<F>* __restrict__ table, XYZZ<F>* __restrict__ terms) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (uint32_t)g.W) return;
  const uint32_t i = t / (uint32_t)g.W, j = t % (uint32_t)g.W;
  XYZZ<F> r = XYZZ<F>::inf();
  if (!inf[base_offset + i]) {
    FrField sc = ld_vec(reinterpret_cast<const FrField*>(scalars) + i);
    if (scalars_mont) sc = FrField::from_mont(sc);
    const uint32_t half = 1u << (g.c - 1);
    uint32_t carry = 0, mag = 0, neg = 0;
    for (uint32_t w = 0; w <= j; w++) {                 // signed digits up to window j (same recoding as k_digits)
      uint32_t raw = window_bits(sc.v, (int)w * g.c, g.c) + carry;
      if (raw > half) { mag = (1u << g.c) - raw; neg = 1; carry = 1; }
      else { mag = raw; neg = 0; carry = 0; }
    }
    if (mag) {
      Affine<F> p = ld_vec(&table[(size_t)j * g.n_srs + base_offset + i]);
      if (neg) p.y = F::neg(p.y);
      r = XYZZ<F>::mul_u32(XYZZ<F>::from_affine(p), mag);
    }
  }
  st_vec(&terms[t], r);
}

// ------------------------------------------------------------------------------------------
// exclusive scan of the histogram (3 small kernels; <= 2^23 counters)
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanTile = kScanThreads * kScanPer;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  uint32_t warp_off = wid ? warp_sums[wid - 1] : 0;
  *total = warp_sums[(blockDim.x >> 5) - 1];
  __syncthreads();
  return warp_off + x - v;
}

static __global__ void k_scan_tiles(uint32_t* data, uint32_t n, uint32_t* tile_sums) {
  uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
  uint32_t v[kScanPer], s = 0;
#pragma unroll
  for (int i = 0; i < kScanPer; i++) { v[i] = base + i < n ? data[base + i] : 0; s += v[i]; }
  uint32_t total;
  uint32_t off = block_exclusive_scan(s, &total);
#pragma unroll
  for (int i = 0; i < kScanPer; i++) { if (base + i < n) data[base + i] = off; off += v[i]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
static __global__ void k_scan_sums(uint32_t* tile_sums, uint32_t n_tiles) {   // one block of 1024, <= 4096 tiles
  uint32_t v[4], s = 0;
  uint32_t base = threadIdx.x * 4;
#pragma unroll
  for (int i = 0; i < 4; i++) { v[i] = base + i < n_tiles ? tile_sums[base + i] : 0; s += v[i]; }
  uint32_t total;
  uint32_t off = block_exclusive_scan(s, &total);
#pragma unroll
  for (int i = 0; i < 4; i++) { if (base + i < n_tiles) tile_sums[base + i] = off; off += v[i]; }
  if (threadIdx.x == 0) tile_sums[n_tiles] = total;
}
// offsets[i] += tile offset; offsets[n] = total; cursor = copy of offsets
static __global__ void k_scan_finish(uint32_t* offsets, uint32_t n, const uint32_t* tile_sums, uint32_t n_tiles,
                              uint32_t* cursor) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t v = offsets[i] + tile_sums[i / kScanTile];
    offsets[i] = v;
    if (cursor) cursor[i] = v;
  }
  if (i == 0) offsets[n] = tile_sums[n_tiles];
}

}  // namespace zkb
#include "msm_affine.cuh"
namespace zkb {

// ------------------------------------------------------------------------------------------
// bucket scheduling: order the regular buckets by decreasing size, list the big ones
// ------------------------------------------------------------------------------------------
struct MsmSched {
  uint32_t n_regular;     // buckets with 1 .. big_threshold entries
  uint32_t n_big;         // buckets above big_threshold
  uint32_t n_chunks;      // total chunk blocks of the big buckets
  uint32_t pad;
};
}  // namespace zkb
#include "msm_batch.cuh"
namespace zkb {

// A bucket is "big" when one thread adding it sequentially would stretch the kernel's critical path:
// a lone warp retires a mixed addition in ~7 us, the whole grid ~2.6 G of them per second, so a bucket
// may hold up to ~entries / 32768 entries before it is worth cutting it into chunks.
// With few buckets (small c) the AVERAGE list is already that long; only lists well above the average are outliers
// (a 2^19-point G2 MSM at c = 16 sent half of its entries through the chunk path: 27 of 43 ms).
inline uint32_t msm_big_threshold(size_t max_entries, uint32_t n_buckets) {
  size_t t = max_entries >> 15;
  const size_t avg4 = 4 * (max_entries / (n_buckets ? n_buckets : 1) + 1);
  if (t < avg4) t = avg4;
  if (t < (size_t)kMinBigBucket) t = kMinBigBucket;
  if (t > (size_t)kSizeBins) t = kSizeBins;
  return (uint32_t)t;
}

// size_hist[big - size] counts regular buckets (descending size order); big buckets are compacted
// into big_list with their chunk ranges
static __global__ void k_size_hist(const uint32_t* __restrict__ offsets, uint32_t n_buckets, uint32_t big,
                                   uint32_t* size_hist, MsmSched* sched, uint32_t* big_list, uint32_t* big_chunk_off) {
  __shared__ uint32_t h[kSizeBins];
  for (uint32_t i = threadIdx.x; i < big; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += gridDim.x * blockDim.x) {
    uint32_t sz = offsets[b + 1] - offsets[b];
    if (sz > big) {
      uint32_t idx = atomicAdd(&sched->n_big, 1u);
      uint32_t nch = (sz + kChunk - 1) / kChunk;
      big_list[idx] = b;
      big_chunk_off[idx] = atomicAdd(&sched->n_chunks, nch);
    } else if (sz) {
      atomicAdd(&h[big - sz], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < big; i += blockDim.x)
    if (h[i]) atomicAdd(&size_hist[i], h[i]);
}
// exclusive scan of the size bins (one block of 1024 threads, 4 bins each) -> cursors; total -> n_regular
static __global__ void k_size_scan(uint32_t* size_hist, uint32_t big, MsmSched* sched) {
  uint32_t v[4], sum = 0;
  uint32_t base = threadIdx.x * 4;
#pragma unroll
  for (int i = 0; i < 4; i++) { v[i] = base + i < big ? size_hist[base + i] : 0; sum += v[i]; }
  uint32_t total;
  uint32_t off = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int i = 0; i < 4; i++) { if (base + i < big) size_hist[base + i] = off; off += v[i]; }
  if (threadIdx.x == 0) sched->n_regular = total;
}
static __global__ void k_size_scatter(const uint32_t* __restrict__ offsets, uint32_t n_buckets, uint32_t big,
                                      uint32_t* size_cursor, uint32_t* __restrict__ order) {
  // warp-aggregated atomics: lanes with the same size share one atomicAdd
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t sz = b < n_buckets ? offsets[b + 1] - offsets[b] : 0;
  bool regular = sz != 0 && sz <= big;
  uint32_t mask = __ballot_sync(0xffffffffu, regular);
  if (!regular) return;
  uint32_t peers = __match_any_sync(mask, sz);
  int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&size_cursor[big - sz], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  uint32_t rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
  order[base + rank] = b;
}
// chunk -> (big bucket slot, chunk index inside the bucket)
static __global__ void k_big_chunk_map(const uint32_t* __restrict__ offsets, const MsmSched* sched,
                                       const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ big_chunk_off,
                                       uint32_t* __restrict__ chunk_slot) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sched->n_big) return;
  uint32_t b = big_list[i];
  uint32_t nch = (offsets[b + 1] - offsets[b] + kChunk - 1) / kChunk;
  uint32_t off = big_chunk_off[i];
  for (uint32_t k = 0; k < nch; k++) chunk_slot[off + k] = i;
}

// ------------------------------------------------------------------------------------------
// bucket accumulation
// ------------------------------------------------------------------------------------------
// one thread per regular bucket, in order of decreasing size.  DIRECT: the lists are the dense output of
// the pair levels (entry = position, no sign, identities possible) instead of (negate | table index) entries.
template <class F, int MIN_BLOCKS, bool DIRECT = false>
__global__ void __launch_bounds__(128, MIN_BLOCKS)
k_accumulate(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets,
             const uint32_t* __restrict__ order, const MsmSched* __restrict__ sched,
             const Affine<F>* __restrict__ table, XYZZ<F>* __restrict__ bucket_acc) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= sched->n_regular) return;
  uint32_t b = order[t];
  uint32_t pos = offsets[b], end = offsets[b + 1];
  XYZZ<F> acc = XYZZ<F>::inf();
  // G1: the next base travels in registers while this one is added.  G2 points are 48 registers and the XYZZ
  // accumulator 96, so there the look-ahead is a prefetch instruction (L2) and the load happens in place.
  constexpr bool kRegPrefetch = sizeof(F) <= 48;
  if (kRegPrefetch) {
    uint32_t e = DIRECT ? pos : entries[pos];
    Affine<F> p = ld_vec(&table[e & 0x7fffffffu]);
    for (;;) {
      pos++;
      uint32_t e_next = 0;
      Affine<F> p_next;
      bool more = pos < end;
      if (more) {                                   // fetch the next base while this one is added
        e_next = DIRECT ? pos : entries[pos];
        p_next = ld_vec(&table[e_next & 0x7fffffffu]);
      }
      if (!DIRECT || !p.is_inf()) acc.madd_xy(p.x, p.y, (e >> 31) != 0);
      if (!more) break;
      e = e_next;
      p = p_next;
    }
  } else {
    uint32_t e = DIRECT ? pos : entries[pos];
    for (;;) {
      pos++;
      const bool more = pos < end;
      const uint32_t e_next = more ? (DIRECT ? pos : entries[pos]) : 0;
      if (more) {
        const char* nx = reinterpret_cast<const char*>(&table[e_next & 0x7fffffffu]);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + sizeof(Affine<F>) - 1));
      }
      const Affine<F> p = ld_vec(&table[e & 0x7fffffffu]);
      if (!DIRECT || !p.is_inf()) acc.madd_xy(p.x, p.y, (e >> 31) != 0);
      if (!more) break;
      e = e_next;
    }
  }
  st_vec(&bucket_acc[b], acc);
}

// block-level tree sum of one XYZZ per thread (result valid in thread 0)
template <class F, int THREADS>
__device__ __forceinline__ void block_sum(XYZZ<F>& acc, XYZZ<F>* sh) {
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      XYZZ<F> a = sh[threadIdx.x], b = sh[threadIdx.x + s];
      pt_add(a, b);
      sh[threadIdx.x] = a;
    }
    __syncthreads();
  }
  acc = sh[0];
}

// one block per 2048-entry chunk of a big bucket -> partial[chunk]
template <class F, bool DIRECT = false>
__global__ void __launch_bounds__(kChunkThreads)
k_big_chunks(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets, const MsmSched* __restrict__ sched,
             const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ big_chunk_off,
             const uint32_t* __restrict__ chunk_slot, const Affine<F>* __restrict__ table, XYZZ<F>* __restrict__ partial) {
  __shared__ XYZZ<F> sh[kChunkThreads];
  for (uint32_t ch = blockIdx.x; ch < sched->n_chunks; ch += gridDim.x) {
    uint32_t slot = chunk_slot[ch];
    uint32_t b = big_list[slot];
    uint32_t k = ch - big_chunk_off[slot];
    uint32_t lo = offsets[b] + k * kChunk, hi = min(lo + (uint32_t)kChunk, offsets[b + 1]);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t pos = lo + threadIdx.x; pos < hi; pos += kChunkThreads) {    // coalesced entry reads
      uint32_t e = DIRECT ? pos : entries[pos];
      Affine<F> p = ld_vec(&table[e & 0x7fffffffu]);
      if (!DIRECT || !p.is_inf()) acc.madd_xy(p.x, p.y, (e >> 31) != 0);
    }
    block_sum<F, kChunkThreads>(acc, sh);
    if (threadIdx.x == 0) st_vec(&partial[ch], acc);
    __syncthreads();
  }
}
// one block per big bucket: sum its chunk partials -> bucket_acc[b]
template <class F>
__global__ void __launch_bounds__(kChunkThreads)
k_big_fold(const uint32_t* __restrict__ offsets, const MsmSched* __restrict__ sched, const uint32_t* __restrict__ big_list,
           const uint32_t* __restrict__ big_chunk_off, const XYZZ<F>* __restrict__ partial, XYZZ<F>* __restrict__ bucket_acc) {
  __shared__ XYZZ<F> sh[kChunkThreads];
  for (uint32_t slot = blockIdx.x; slot < sched->n_big; slot += gridDim.x) {
    uint32_t b = big_list[slot];
    uint32_t nch = (offsets[b + 1] - offsets[b] + kChunk - 1) / kChunk;
    const XYZZ<F>* src = partial + big_chunk_off[slot];
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = threadIdx.x; k < nch; k += kChunkThreads) {
      XYZZ<F> q = ld_vec_rw(&src[k]);
      pt_add(acc, q);
    }
    block_sum<F, kChunkThreads>(acc, sh);
    if (threadIdx.x == 0) st_vec(&bucket_acc[b], acc);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// bucket reduction through row / column sums
// ------------------------------------------------------------------------------------------
// out[set][o] = sum_{k < K} in[set][(o / inner) * outer_stride + (o % inner) * inner_stride + k * step]
template <class F>
__global__ void __launch_bounds__(128)
k_seg_sum(const XYZZ<F>* __restrict__ in, size_t in_set_stride, XYZZ<F>* __restrict__ out, uint32_t n_out, uint32_t inner,
          uint32_t inner_stride, uint32_t outer_stride, uint32_t step, uint32_t K) {
  uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const XYZZ<F>* src = in + (size_t)blockIdx.y * in_set_stride + (size_t)(o / inner) * outer_stride +
                       (size_t)(o % inner) * inner_stride;
  XYZZ<F> acc = ld_vec_rw(src);
  for (uint32_t k = 1; k < K; k++) {
    XYZZ<F> q = ld_vec_rw(src + (size_t)k * step);
    pt_add(acc, q);
  }
  st_vec(&out[(size_t)blockIdx.y * n_out + o], acc);
}
// The same pass for BOTH axes of the bucket matrix in one launch (blockIdx.z = axis): the row sums and the column sums
// are independent chains of latency-bound passes (~17 us per dependent G1 addition, ~55 us for G2), so running them side
// by side halves the serial depth of the reduction.  An axis that is already done has n_out == 0.
template <class F>
struct SegPass {
  const XYZZ<F>* in;
  size_t in_set_stride;
  XYZZ<F>* out;
  uint32_t n_out, inner, inner_stride, outer_stride, step, K;
};
template <class F>
__global__ void __launch_bounds__(128)
k_seg_sum2(SegPass<F> pa, SegPass<F> pb) {
  const SegPass<F>& p = blockIdx.z ? pb : pa;
  uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= p.n_out) return;
  const XYZZ<F>* src = p.in + (size_t)blockIdx.y * p.in_set_stride + (size_t)(o / p.inner) * p.outer_stride +
                       (size_t)(o % p.inner) * p.inner_stride;
  XYZZ<F> acc = ld_vec_rw(src);
  for (uint32_t k = 1; k < p.K; k++) {
    XYZZ<F> q = ld_vec_rw(src + (size_t)k * p.step);
    pt_add(acc, q);
  }
  st_vec(&p.out[(size_t)blockIdx.y * p.n_out + o], acc);
}
// weighted[set][i] = (i * L) * rows[set][i] for i < H;  weighted[set][H + j] = (j + 1) * cols[set][j] for j < L
template <class F>
__global__ void __launch_bounds__(128)
k_weight_rows_cols(const XYZZ<F>* __restrict__ rows, const XYZZ<F>* __restrict__ cols, uint32_t H, uint32_t L,
                   XYZZ<F>* __restrict__ weighted) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H + L) return;
  uint32_t set = blockIdx.y;
  XYZZ<F> p = i < H ? ld_vec_rw(&rows[(size_t)set * H + i]) : ld_vec_rw(&cols[(size_t)set * L + (i - H)]);
  uint32_t k = i < H ? i * L : (i - H) + 1;
  XYZZ<F> r = XYZZ<F>::inf();
  if (k && !p.is_inf()) r = XYZZ<F>::mul_u32(p, k);
  st_vec(&weighted[(size_t)set * (H + L) + i], r);
}

// out[set][blockIdx.x] = sum of a slice of in[set][...]
template <class F>
__global__ void __launch_bounds__(64)
k_sum_points(const XYZZ<F>* __restrict__ in, uint32_t n_per_set, XYZZ<F>* __restrict__ out, uint32_t per_thread) {
  __shared__ XYZZ<F> sh[64];
  uint32_t set = blockIdx.y;
  const XYZZ<F>* src = in + (size_t)set * n_per_set;
  uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) * per_thread;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t i = first; i < first + per_thread && i < n_per_set; i++) { XYZZ<F> q = ld_vec_rw(&src[i]); pt_add(acc, q); }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 32; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      XYZZ<F> a = sh[threadIdx.x], b = sh[threadIdx.x + s];
      pt_add(a, b);
      sh[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_vec(&out[(size_t)set * gridDim.x + blockIdx.x], sh[0]);
}

// combine window sums (high -> low, c doublings between windows); result in out[0]
template <class F>
__global__ void k_window_combine(const XYZZ<F>* __restrict__ sums, uint32_t n_sets, int c, XYZZ<F>* __restrict__ out) {
  if (threadIdx.x | blockIdx.x) return;
  XYZZ<F> total = ld_vec_rw(&sums[n_sets - 1]);
  for (int j = (int)n_sets - 2; j >= 0; j--) {
    for (int k = 0; k < c; k++) pt_dbl(total);
    XYZZ<F> q = ld_vec_rw(&sums[j]);
    pt_add(total, q);
  }
  st_vec(out, total);
}

// XYZZ -> canonical affine + infinity flag
template <class F>
__global__ void k_to_affine(const XYZZ<F>* __restrict__ in, uint32_t n, Affine<F>* __restrict__ out_xy, uint8_t* __restrict__ out_inf) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> p = ld_vec_rw(&in[i]);
  Affine<F> a;
  pt_to_affine(a, p);
  st_vec(&out_xy[i], a);
  out_inf[i] = p.is_inf() ? 1 : 0;
}

// sum of `count` points in index order (= rank order of the all-gather of per-rank partials) -> out_pt and/or
// canonical affine.  One thread: count <= 8 ranks, ~14 multiplications per addition.
template <class F>
__global__ void k_fold_points(const XYZZ<F>* __restrict__ in, uint32_t count, uint32_t stride_bytes,
                              XYZZ<F>* __restrict__ out_pt, Affine<F>* __restrict__ out_xy, uint8_t* __restrict__ out_inf) {
  if (threadIdx.x | blockIdx.x) return;
  const char* base = reinterpret_cast<const char*>(in);
  XYZZ<F> acc = ld_vec_rw(reinterpret_cast<const XYZZ<F>*>(base));
  for (uint32_t i = 1; i < count; i++) {
    XYZZ<F> q = ld_vec_rw(reinterpret_cast<const XYZZ<F>*>(base + (size_t)i * stride_bytes));
    pt_add(acc, q);
  }
  if (out_pt) st_vec(out_pt, acc);
  if (out_xy) {
    Affine<F> a;
    pt_to_affine(a, acc);
    st_vec(out_xy, a);
    *out_inf = acc.is_inf() ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// SRS ingestion: apply infinity flags, precompute 2^(c*j) * P_i
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void k_apply_inf(Affine<F>* pts, const uint8_t* inf, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && inf[i]) st_vec(&pts[i], Affine<F>::inf());
}
template <class F>
__global__ void __launch_bounds__(128)
k_precompute(Affine<F>* table, uint32_t n, int c, int W) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = ld_vec_rw(&table[i]);
  XYZZ<F> q = XYZZ<F>::from_affine(p);
  for (int j = 1; j < W; j++) {
    for (int k = 0; k < c; k++) pt_dbl(q);
    Affine<F> a;
    pt_to_affine(a, q);
    st_vec(&table[(size_t)j * n + i], a);
    q = XYZZ<F>::from_affine(a);       // keep Z = 1 so the next inversion input stays small
  }
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
// Window width from a cost model in field multiplications: n * W mixed additions (10 mul) plus the
// bucket reduction, ~2.3 full additions (14 mul) per bucket of every bucket set.  Checked against a
// sweep of c on the 2^20-constraint proof (c = 16..22 -> 80, 53, 48, 47, 43, 49, 51 ms).
inline int msm_pick_c(size_t n, int precomp, int scalar_bits) {
  if (n < 2) n = 2;
  int best_c = 4;
  double best = 1e300;
  const int c_max = precomp ? 23 : 16;
  for (int c = 4; c <= c_max; c++) {
    int W = msm_windows(scalar_bits, c);
    double sets = precomp ? 1.0 : (double)W;
    double cost = (double)n * W * 10.0 + sets * (double)(size_t(1) << (c - 1)) * 32.0;
    // a narrow top window funnels its n entries into a few buckets, which then take the slower chunked
    // path (measured: c = 19 loses to c = 20 at n = 2^19..2^20 although it has fewer buckets)
    int top_bits = scalar_bits + 1 - c * (W - 1);
    if (top_bits < c - 5) cost += (double)n * 20.0;
    if (!precomp) cost += (double)W * c * 9.0;      // doubling chain of the window combine (negligible)
    if (cost < best) { best = cost; best_c = c; }
  }
  // Mid-size MSMs over window tables (the shards of a sharded proof, Marlin's commitments): the model above picks
  // c = 16 for 2^15 <= n <= 2^19, but one thread per bucket then has only 2^15 threads -- less than one wave of the
  // machine -- walking lists of 100+ entries.  Measured (BLS12-381, ms per MSM, c = 16 / 20; profiles/r2u_window_sweep.txt):
  // G1 2^17 2.39 / 2.20, 2^18 3.38 / 2.83, 2^19 5.44 / 4.15; G2 2^17 8.22 / 7.34, 2^18 11.39 / 9.59 (2^16: 6.11 / 6.36).
  // Widths whose top window is narrow (17, 18, 19 for 255 bits) are worse than both.
  if (precomp && n >= (size_t(1) << 17) && best_c < 20) best_c = 20;
  if (const char* e = getenv(precomp ? "ZKB_MSM_C" : "ZKB_MSM_C_NOPRE")) {
    int v = atoi(e);
    if (v >= 2 && v <= 23) best_c = v;
  }
  return best_c;
}

// Batched-affine bucket accumulation (msm_batch.cuh) instead of the XYZZ loop.  ZKB_MSM_BATCH=0/1 overrides (read per
// call: the parity tests run both).
inline bool msm_batch_affine() {
  // Default off: measured on B200 (2^20 BLS12-381 G1) 13.6 ms against 6.0 ms for the XYZZ loop -- with one list per
  // slot the 2^19 lists of unequal length (Poisson, mean 26, max ~55) leave the one-wave grid half empty towards the
  // end, and G = 12 additions per inversion make the inversion as expensive as the additions.
  const char* e = getenv("ZKB_MSM_BATCH");
  return e ? atoi(e) != 0 : false;
}

// Pair levels before the XYZZ accumulation: each level moves half of the remaining additions to the
// cheaper batched-affine form but costs a scan and a kernel of its own, so stop when the lists are short
// (average load: entries per bucket).  ZKB_MSM_PAIR_LEVELS overrides (0 = XYZZ only).
inline int msm_pair_levels(size_t max_entries, uint32_t n_buckets) {
  if (max_entries >= (size_t(1) << 31)) return 0;       // meta packs the source position into 31 bits
  const char* e = getenv("ZKB_MSM_PAIR_LEVELS");        // read per call: the tests sweep it
  const int forced = e ? atoi(e) : -1;
  if (forced >= 0) return forced > 12 ? 12 : forced;
  // Default: none.  Measured on B200 (2^20 BLS12-381 G1): the level kernels reach ~55 % of the
  // multiplier pipe against 83 % for the XYZZ loop (one dependent multiplication chain per thread instead of
  // three), which cancels the 40 % saving in multiplications: 5.1 ms with three levels vs 4.9 ms without.
  (void)n_buckets;
  return 0;
}

// resident blocks per SM of a level kernel.  The dynamic shared-memory opt-in is a per-DEVICE function attribute, so it
// is set on every call (a process may hold contexts on several GPUs); the occupancy answer is the same on all of them.
inline int pair_level_occupancy(const void* kernel, size_t smem_bytes) {
  int blocks = 0;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, kPairThreads, smem_bytes) != cudaSuccess || blocks < 1)
    blocks = 1;
  return blocks;
}
// slots per thread and chunk = scale * (5 .. 11): large enough to amortise the block's inversion and tree, small
// enough for >= ~6 chunks per resident block
inline uint32_t pair_level_scale(size_t slots_bound, unsigned resident_blocks) {
  const char* e = getenv("ZKB_PAIR_SCALE");             // read per call: the tuning script sweeps it
  const int forced = e ? atoi(e) : 0;
  if (forced >= 1 && forced <= 16) return (uint32_t)forced;
  size_t per_block = slots_bound / ((size_t)resident_blocks * kPairThreads * 8 * 6 + 1);
  return per_block < 1 ? 1u : per_block > 4 ? 4u : (uint32_t)per_block;
}

template <class F, class FrP>
struct MsmEngine {
  using Fr = Fp<FrP>;
  using Pt = XYZZ<F>;
  using Aff = Affine<F>;

  static int srs_build(zkb_ctx* ctx, zkb_srs* srs, const void* h_xy, const uint8_t* h_inf, unsigned flags) {
    cudaStream_t st = ctx->main;
    size_t n = srs->n;
    srs->precomp = (flags & ZKB_SRS_PRECOMPUTE) ? 1 : 0;
    size_t n_eff = 0;                    // identity bases never produce bucket entries
    for (size_t i = 0; i < n; i++) n_eff += h_inf[i] ? 0 : 1;
    srs->c = msm_pick_c(n_eff, srs->precomp, FrP::BITS);
    if (srs->precomp && sizeof(F) > sizeof(Fp<typename F::Params>)) {          // G2 (Fq2 coordinates): its own override
      if (const char* e = getenv("ZKB_MSM_C_G2")) {
        int v = atoi(e);
        if (v >= 2 && v <= 23) srs->c = v;
      }
    }
    srs->W = msm_windows(FrP::BITS, srs->c);
    if (srs->precomp && (size_t)srs->W * n >= (size_t(1) << 31))
      return set_err(ctx, ZKB_E_INVALID, "precomputed table too large for 31-bit indices");
    if (n >= (size_t(1) << 31)) return set_err(ctx, ZKB_E_INVALID, "too many bases");
    size_t copies = srs->precomp ? srs->W : 1;
    ZKB_CUDA(ctx, cudaMalloc(&srs->table, sizeof(Aff) * copies * (n ? n : 1)));
    ZKB_CUDA(ctx, cudaMalloc((void**)&srs->inf, n ? n : 1));
    if (n == 0) return ZKB_OK;
    ZKB_CUDA(ctx, cudaMemcpyAsync(srs->table, h_xy, sizeof(Aff) * n, cudaMemcpyHostToDevice, st));
    ZKB_CUDA(ctx, cudaMemcpyAsync(srs->inf, h_inf, n, cudaMemcpyHostToDevice, st));
    ZKB_LAUNCH(ctx, (k_apply_inf<F>), ceil_div(n, 256), 256, 0, st, (Aff*)srs->table, srs->inf, (uint32_t)n);
    if (srs->precomp && srs->W > 1)
      ZKB_LAUNCH(ctx, (k_precompute<F>), ceil_div(n, 128), 128, 0, st, (Aff*)srs->table, (uint32_t)n, srs->c, srs->W);
    ZKB_CUDA(ctx, cudaStreamSynchronize(st));
    return ZKB_OK;
  }

  // d_result: one XYZZ point (device).  All work is enqueued on `st`.
  static int run(zkb_ctx* ctx, cudaStream_t st, const zkb_srs* srs, size_t base_offset, const uint32_t* d_scalars,
                 size_t n, int scalars_mont, void* d_result_v) {
    return run_split(ctx, st, srs, base_offset, d_scalars, n, scalars_mont, d_result_v, nullptr);
  }
  // Two-phase form: with `deferred` != nullptr only the sort (digits, counting sort, bucket schedule) is enqueued now and
  // *deferred receives the closure that enqueues the accumulation and the bucket reduction.  A caller that runs several
  // MSMs on several streams enqueues all the sorts first: kernels are dispatched roughly in launch order, so a sort
  // launched after another MSM's accumulation kernel waits behind its thousands of pending blocks (the 2.5 ms "sort
  // phase" after the G2 accumulation in the round-2 proof timeline).
  static int run_split(zkb_ctx* ctx, cudaStream_t st, const zkb_srs* srs, size_t base_offset, const uint32_t* d_scalars,
                       size_t n, int scalars_mont, void* d_result_v, std::function<int()>* deferred) {
    if (deferred) *deferred = []() -> int { return ZKB_OK; };
    Pt* d_result = (Pt*)d_result_v;
    if (base_offset + n > srs->n) return set_err(ctx, ZKB_E_INVALID, "msm: base range out of bounds");
    if (n == 0) {
      ZKB_CUDA(ctx, cudaMemsetAsync(d_result, 0, sizeof(Pt), st));
      return ZKB_OK;
    }
    MsmGeom g;
    g.c = srs->c; g.W = srs->W; g.precomp = srs->precomp;
    if (!g.precomp) {
      // without window tables nothing ties the window width to the SRS: choose it for THIS call's n, so that a short
      // MSM over a slice of a long key (a Marlin commitment to a low-degree polynomial, a Pedersen row commitment) does
      // not pay the bucket reduction of the full-length one
      g.c = msm_pick_c(n, 0, FrP::BITS);
      g.W = msm_windows(FrP::BITS, g.c);
    }
    g.B = 1u << (g.c - 1);
    g.n_srs = (uint32_t)srs->n; g.n_sets = g.precomp ? 1u : (uint32_t)g.W;
    if (g.precomp && n * (size_t)g.W <= kSmallMsmTerms) {
      Scratch ws(ctx, st);
      const uint32_t n_terms = (uint32_t)n * (uint32_t)g.W;
      Pt *terms, *tmp;
      ZKB_TRY(ws.alloc(&terms, n_terms));
      ZKB_TRY(ws.alloc(&tmp, ceil_div(n_terms, 64) + 1));
      ZKB_LAUNCH(ctx, (k_small_msm_terms<F, Fr>), ceil_div(n_terms, 128), 128, 0, st, d_scalars, (uint32_t)n, srs->inf, g,
                 (uint32_t)base_offset, scalars_mont, (const Aff*)srs->table, terms);
      uint32_t n_per = n_terms;
      Pt *pin = terms, *pout = tmp;
      while (n_per > 1) {
        uint32_t n_blocks = ceil_div(n_per, 64);
        ZKB_LAUNCH(ctx, (k_sum_points<F>), dim3(n_blocks, 1), 64, 0, st, pin, n_per, pout, 1u);
        Pt* nx = pin;
        pin = pout;
        pout = nx;
        n_per = n_blocks;
      }
      ZKB_CUDA(ctx, cudaMemcpyAsync(d_result, pin, sizeof(Pt), cudaMemcpyDeviceToDevice, st));
      return ZKB_OK;
    }
    const uint32_t n_buckets = g.B * g.n_sets;
    if (n_buckets > (1u << 23)) return set_err(ctx, ZKB_E_INVALID, "msm: too many buckets");
    const size_t max_entries = n * (size_t)g.W;
    if (max_entries >= (size_t(1) << 32)) return set_err(ctx, ZKB_E_INVALID, "msm: too many entries");

    auto wsp = std::make_shared<Scratch>(ctx, st);       // lives until the deferred phase has been enqueued
    Scratch& ws = *wsp;
    uint32_t *offsets, *cursor, *tile_sums, *entries;
    const uint32_t n_tiles = ceil_div(n_buckets, kScanTile);
    ZKB_TRY(ws.alloc(&offsets, (size_t)n_buckets + 1));
    ZKB_TRY(ws.alloc(&cursor, n_buckets));
    ZKB_TRY(ws.alloc(&tile_sums, (size_t)n_tiles + 1));
    ZKB_TRY(ws.alloc(&entries, max_entries));
    ZKB_CUDA(ctx, cudaMemsetAsync(offsets, 0, sizeof(uint32_t) * ((size_t)n_buckets + 1), st));

    const unsigned dig_blocks = min(ceil_div(n, 256), (unsigned)(ctx->sm_count * 8));
    ZKB_LAUNCH(ctx, (k_digits<false, Fr>), dig_blocks, 256, 0, st, d_scalars, (uint32_t)n, srs->inf, g,
               (uint32_t)base_offset, scalars_mont, offsets, (uint32_t*)nullptr);
    ZKB_LAUNCH(ctx, k_scan_tiles, n_tiles, kScanThreads, 0, st, offsets, n_buckets, tile_sums);
    ZKB_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, st, tile_sums, n_tiles);
    ZKB_LAUNCH(ctx, k_scan_finish, ceil_div(n_buckets, 256), 256, 0, st, offsets, n_buckets, tile_sums, n_tiles, cursor);
    ZKB_LAUNCH(ctx, (k_digits<true, Fr>), dig_blocks, 256, 0, st, d_scalars, (uint32_t)n, srs->inf, g,
               (uint32_t)base_offset, scalars_mont, cursor, entries);

    // ---- batched-affine pair levels (msm_affine.cuh): each halves the bucket lists
    const uint32_t* acc_off = offsets;          // lists that the XYZZ accumulation below consumes
    const Aff* acc_pts = (const Aff*)srs->table;
    size_t acc_max = max_entries;               // upper bound of their total length
    const int levels = msm_pair_levels(max_entries, n_buckets);
    if (levels > 0) {
      auto halved = [&](size_t e) { return (e + (e < n_buckets ? e : (size_t)n_buckets) + 1) / 2; };
      const size_t e1 = halved(max_entries), e2 = halved(e1);
      uint32_t *off_a, *off_b, *work_counters;
      uint2* recs;
      F* prefix;
      Aff *pts_a, *pts_b;
      ZKB_TRY(ws.alloc(&off_a, (size_t)n_buckets + 1));
      ZKB_TRY(ws.alloc(&off_b, (size_t)n_buckets + 1));
      ZKB_TRY(ws.alloc(&recs, e1));
      ZKB_TRY(ws.alloc(&prefix, e1));
      ZKB_TRY(ws.alloc(&pts_a, e1));
      ZKB_TRY(ws.alloc(&pts_b, levels > 1 ? e2 : 1));
      ZKB_TRY(ws.alloc(&work_counters, (size_t)levels));
      ZKB_CUDA(ctx, cudaMemsetAsync(work_counters, 0, sizeof(uint32_t) * (size_t)levels, st));
      const uint32_t* off_in = offsets;
      size_t e_out = e1;
      for (int lvl = 0; lvl < levels; lvl++) {
        uint32_t* off_out = (lvl & 1) ? off_b : off_a;
        Aff* pts_out = (lvl & 1) ? pts_b : pts_a;
        ZKB_LAUNCH(ctx, k_pair_sizes, ceil_div(n_buckets, 256), 256, 0, st, off_in, n_buckets, off_out);
        ZKB_LAUNCH(ctx, k_scan_tiles, n_tiles, kScanThreads, 0, st, off_out, n_buckets, tile_sums);
        ZKB_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, st, tile_sums, n_tiles);
        ZKB_LAUNCH(ctx, k_scan_finish, ceil_div(n_buckets, 256), 256, 0, st, off_out, n_buckets, tile_sums, n_tiles,
                   (uint32_t*)nullptr);
        // one resident wave; the kernel derives the slots per thread from the list length on the device
        ZKB_TRY(on_bulk_stream(ctx, st, [&](cudaStream_t bs) -> int {
          const int occ = pair_level_occupancy((const void*)k_pair_level<F>, PairRing<F>::kBytes);
          const unsigned blocks = (unsigned)(ctx->sm_count * occ);
          prof_begin(ctx, bs);
          ZKB_LAUNCH(ctx, (k_pair_level<F>), blocks, kPairThreads, PairRing<F>::kBytes, bs,
                     lvl == 0 ? (const uint32_t*)entries : (const uint32_t*)nullptr, lvl == 0 ? (const Aff*)srs->table : acc_pts,
                     off_in, (const uint32_t*)off_out, n_buckets, recs, prefix, pts_out, work_counters + lvl,
                     pair_level_scale(e_out, blocks));
          prof_end(ctx, bs, 0.0);
          return ZKB_OK;
        }));
        off_in = off_out;
        acc_off = off_out;
        acc_pts = pts_out;
        acc_max = e_out;
        e_out = halved(e_out);
      }
    }
    const bool direct = levels > 0;

    // schedule: regular buckets by decreasing size, big buckets in chunks
    uint32_t *size_hist, *order, *big_list, *big_chunk_off, *chunk_slot;
    MsmSched* sched;
    const uint32_t big = msm_big_threshold(acc_max, n_buckets);
    const uint32_t max_big = (uint32_t)(acc_max / (big + 1)) + 1;
    const uint32_t max_chunks = (uint32_t)(acc_max / kChunk) + max_big;
    ZKB_TRY(ws.alloc(&size_hist, (size_t)kSizeBins + 4));       // bins followed by the MsmSched block
    sched = reinterpret_cast<MsmSched*>(size_hist + kSizeBins);
    ZKB_TRY(ws.alloc(&order, n_buckets));
    ZKB_TRY(ws.alloc(&big_list, max_big));
    ZKB_TRY(ws.alloc(&big_chunk_off, max_big));
    ZKB_TRY(ws.alloc(&chunk_slot, max_chunks));
    ZKB_CUDA(ctx, cudaMemsetAsync(size_hist, 0, sizeof(uint32_t) * (kSizeBins + 4), st));
    {
      unsigned hist_blocks = ceil_div(n_buckets, 1024);
      if (hist_blocks > (unsigned)ctx->sm_count * 2) hist_blocks = ctx->sm_count * 2;
      ZKB_LAUNCH(ctx, k_size_hist, hist_blocks, 1024, 0, st, acc_off, n_buckets, big, size_hist, sched, big_list,
                 big_chunk_off);
    }
    ZKB_LAUNCH(ctx, k_size_scan, 1, 1024, 0, st, size_hist, big, sched);
    ZKB_LAUNCH(ctx, k_size_scatter, ceil_div(n_buckets, 256), 256, 0, st, acc_off, n_buckets, big, size_hist, order);
    ZKB_LAUNCH(ctx, k_big_chunk_map, ceil_div(max_big, 256), 256, 0, st, acc_off, sched, big_list, big_chunk_off, chunk_slot);

    // accumulate
    Pt *bucket_acc, *partial;
    ZKB_TRY(ws.alloc(&bucket_acc, n_buckets));
    ZKB_TRY(ws.alloc(&partial, max_chunks));
    ZKB_CUDA(ctx, cudaMemsetAsync(bucket_acc, 0, sizeof(Pt) * (size_t)n_buckets, st));     // empty buckets = identity
    // batched-affine accumulation (msm_batch.cuh): two levels of equal-length chains, then the XYZZ combine
    const bool batch_affine = levels == 0 && msm_batch_affine();
    uint32_t *seg_off0 = nullptr, *seg_off1 = nullptr;
    uint2 *rec0 = nullptr, *rec1 = nullptr;
    Aff *chain_sum0 = nullptr, *chain_sum1 = nullptr;
    F* chain_prefix = nullptr;
    unsigned batch_blocks = 0;
    if (batch_affine) {
      using FC = typename CallVariant<F>::type;
      auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
      int occ = 0;
      ZKB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_accumulate_chains<FC, true>,
                                                                  kBatchThreads, 0));
      int bps = env_int("ZKB_BATCH_BPS", 4);                // 8 warps per SM keep the multiplier pipe > 90 % busy (chains.cu)
      if (bps > occ) bps = occ;
      if (bps < 1) bps = 1;
      batch_blocks = (unsigned)(ctx->sm_count * bps);
      const size_t lanes = (size_t)batch_blocks * kBatchThreads;
      int l0 = env_int("ZKB_BATCH_L0", 8), l1 = env_int("ZKB_BATCH_L1", 8);
      if (l0 < 2) l0 = 2;
      if (l1 < 2) l1 = 2;
      const size_t max_chains0 = acc_max / (size_t)l0 + n_buckets + 1, max_chains1 = max_chains0 / (size_t)l1 + n_buckets + 1;
      ZKB_TRY(ws.alloc(&seg_off0, (size_t)n_buckets + 1));
      ZKB_TRY(ws.alloc(&seg_off1, (size_t)n_buckets + 1));
      ZKB_TRY(ws.alloc(&rec0, max_chains0));
      ZKB_TRY(ws.alloc(&rec1, max_chains1));
      ZKB_TRY(ws.alloc(&chain_sum0, max_chains0));
      ZKB_TRY(ws.alloc(&chain_sum1, max_chains1));
      ZKB_TRY(ws.alloc(&chain_prefix, lanes * kBatchGMax));
      // the chain structure of both levels depends on the list lengths only: built before the accumulation starts
      const uint32_t* lists = acc_off;
      uint32_t* seg[2] = {seg_off0, seg_off1};
      uint2* recs[2] = {rec0, rec1};
      for (int lvl = 0; lvl < 2; lvl++) {
        ZKB_LAUNCH(ctx, k_chain_count, ceil_div(n_buckets, 256), 256, 0, st, lists, n_buckets, lvl == 0 ? big : 0xffffffffu,
                   (uint32_t)(lvl == 0 ? l0 : l1), seg[lvl]);
        ZKB_LAUNCH(ctx, k_scan_tiles, n_tiles, kScanThreads, 0, st, seg[lvl], n_buckets, tile_sums);
        ZKB_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, st, tile_sums, n_tiles);
        ZKB_LAUNCH(ctx, k_scan_finish, ceil_div(n_buckets, 256), 256, 0, st, seg[lvl], n_buckets, tile_sums, n_tiles,
                   (uint32_t*)nullptr);
        ZKB_LAUNCH(ctx, k_chain_build, ceil_div(n_buckets, 256), 256, 0, st, lists, (const uint32_t*)seg[lvl], n_buckets,
                   recs[lvl]);
        lists = seg[lvl];
      }
    }
    auto phase2 = [=]() -> int {
    Scratch& ws = *wsp;
    // the accumulation kernel fills the machine: it goes to the low-priority bulk stream (common.cuh)
    ZKB_TRY(on_bulk_stream(ctx, st, [&](cudaStream_t bs) -> int {
      using FC = typename CallVariant<F>::type;
      static_assert(sizeof(Affine<FC>) == sizeof(Aff) && sizeof(XYZZ<FC>) == sizeof(Pt), "call variant layout");
      // tuning switches (defaults chosen from measurements): multiplication as a call,
      // and a register cap that trades a few spills for a fourth resident block per SM
      const int use_call = []() { const char* e = getenv("ZKB_ACC_CALL"); return e ? atoi(e) : 0; }();
      static const int occ4 = []() { const char* e = getenv("ZKB_ACC_OCC4"); return e ? atoi(e) : 0; }();
      prof_begin(ctx, bs);
      if (batch_affine) {
        // 6 instead of 10 multiplications per entry: level 0 gathers from the table, level 1 sums the chain sums
        ZKB_LAUNCH(ctx, (k_accumulate_chains<FC, true>), batch_blocks, kBatchThreads, 0, bs, entries,
                   (const Affine<FC>*)srs->table, (const uint2*)rec0, (const uint32_t*)(seg_off0 + n_buckets),
                   (Affine<FC>*)chain_sum0, (FC*)chain_prefix, kBatchGMax);
        ZKB_LAUNCH(ctx, (k_accumulate_chains<FC, false>), batch_blocks, kBatchThreads, 0, bs, (const uint32_t*)nullptr,
                   (const Affine<FC>*)chain_sum0, (const uint2*)rec1, (const uint32_t*)(seg_off1 + n_buckets),
                   (Affine<FC>*)chain_sum1, (FC*)chain_prefix, kBatchGMax);
        ZKB_LAUNCH(ctx, (k_chain_combine<FC>), ceil_div(n_buckets, 128), 128, 0, bs, (const uint32_t*)seg_off1, n_buckets,
                   (const Affine<FC>*)chain_sum1, (XYZZ<FC>*)bucket_acc);
      } else if (direct)
        ZKB_LAUNCH(ctx, (k_accumulate<F, 1, true>), ceil_div(n_buckets, 128), 128, 0, bs, (const uint32_t*)nullptr, acc_off,
                   order, sched, acc_pts, bucket_acc);
      else if (use_call & (sizeof(F) <= 48 ? 1 : 2))      // bit 0: G1, bit 1: G2
        ZKB_LAUNCH(ctx, (k_accumulate<FC, (sizeof(F) <= 48 ? 3 : 2)>), ceil_div(n_buckets, 128), 128, 0, bs, entries, offsets,
                   order, sched, (const Affine<FC>*)srs->table, (XYZZ<FC>*)bucket_acc);
      else if (occ4)
        ZKB_LAUNCH(ctx, (k_accumulate<F, 4>), ceil_div(n_buckets, 128), 128, 0, bs, entries, offsets, order, sched,
                   (const Aff*)srs->table, bucket_acc);
      else
        ZKB_LAUNCH(ctx, (k_accumulate<F, 1>), ceil_div(n_buckets, 128), 128, 0, bs, entries, offsets, order, sched,
                   (const Aff*)srs->table, bucket_acc);
      prof_end(ctx, bs, (double)n * (32.0 + sizeof(Aff)));   // one read of each (scalar, base) pair
      return ZKB_OK;
    }));
    {
      // grids are upper bounds read against device-side counts (no host round trip)
      unsigned chunk_blocks = max_chunks < (unsigned)ctx->sm_count * 8 ? max_chunks : ctx->sm_count * 8;
      unsigned fold_blocks = max_big < (unsigned)ctx->sm_count * 4 ? max_big : ctx->sm_count * 4;
      if (direct)
        ZKB_LAUNCH(ctx, (k_big_chunks<F, true>), chunk_blocks, kChunkThreads, 0, st, (const uint32_t*)nullptr, acc_off, sched,
                   big_list, big_chunk_off, chunk_slot, acc_pts, partial);
      else
        ZKB_LAUNCH(ctx, (k_big_chunks<F>), chunk_blocks, kChunkThreads, 0, st, entries, offsets, sched, big_list,
                   big_chunk_off, chunk_slot, (const Aff*)srs->table, partial);
      ZKB_LAUNCH(ctx, (k_big_fold<F>), fold_blocks, kChunkThreads, 0, st, acc_off, sched, big_list, big_chunk_off,
                 (const Pt*)partial, bucket_acc);
    }

    // reduce: bucket b = hi * L + lo carries weight b + 1 = hi * L + (lo + 1)
    const unsigned lo_bits = (unsigned)(g.c - 1) / 2, hi_bits = (unsigned)(g.c - 1) - lo_bits;
    const uint32_t L = 1u << lo_bits, H = 1u << hi_bits;
    Pt *tmp_a, *tmp_b, *rows, *cols, *weighted;
    ZKB_TRY(ws.alloc(&tmp_a, (size_t)g.n_sets * (g.B / 2 + 1)));
    ZKB_TRY(ws.alloc(&tmp_b, (size_t)g.n_sets * (g.B / 4 + 1)));
    ZKB_TRY(ws.alloc(&rows, (size_t)g.n_sets * H));
    ZKB_TRY(ws.alloc(&cols, (size_t)g.n_sets * L));
    ZKB_TRY(ws.alloc(&weighted, (size_t)g.n_sets * (H + L)));
    auto reduce_axis = [&](bool along_rows, Pt* final_out) -> int {
      // along_rows: sum over lo (length L) for every hi; else sum over hi (length H) for every lo
      uint32_t len = along_rows ? L : H;          // remaining length of the reduced axis
      const uint32_t other = along_rows ? H : L;
      const Pt* in = bucket_acc;
      size_t in_set_stride = g.B;
      if (len == 1) {
        ZKB_LAUNCH(ctx, (k_seg_sum<F>), dim3(ceil_div(other, 128), g.n_sets), 128, 0, st, in, in_set_stride, final_out,
                   other, other, 1u, 1u, 1u, 1u);
        return ZKB_OK;
      }
      Pt* bufs[2] = {tmp_a, tmp_b};
      int flip = 0;
      while (len > 1) {
        // 4-way passes while the axis is long (many short waves instead of two long ones: the 8-way first pass was 1.15
        // waves of 7 dependent additions), one last pass of up to 8
        const uint32_t k_env = []() { const char* e = getenv("ZKB_SEG_K"); return e ? (uint32_t)atoi(e) : 0u; }();
        const uint32_t k_pass = k_env >= 2 && k_env <= 16 && !(k_env & (k_env - 1)) ? k_env : 4u;
        uint32_t K = len <= (uint32_t)kSegK ? len : k_pass;
        uint32_t new_len = len / K;                 // powers of two throughout
        uint32_t n_out = new_len * other;
        Pt* out = new_len == 1 ? final_out : bufs[flip];
        if (along_rows)      // element (hi, j) of a [other][len] matrix; output [other][new_len]
          ZKB_LAUNCH(ctx, (k_seg_sum<F>), dim3(ceil_div(n_out, 128), g.n_sets), 128, 0, st, in, in_set_stride, out, n_out,
                     new_len, K, len, 1u, K);
        else                 // element (i, lo) of a [len][other] matrix; output [new_len][other]
          ZKB_LAUNCH(ctx, (k_seg_sum<F>), dim3(ceil_div(n_out, 128), g.n_sets), 128, 0, st, in, in_set_stride, out, n_out,
                     other, 1u, K * other, other, K);
        in = out;
        in_set_stride = n_out;
        len = new_len;
        flip ^= 1;
      }
      return ZKB_OK;
    };
    static const int fused_axes = []() { const char* e = getenv("ZKB_REDUCE_FUSED"); return e ? atoi(e) : 1; }();
    if (fused_axes && lo_bits >= 2) {
      // both axes side by side: pass i sums 4 (the last pass of an axis 2 when its bit count is odd) along each axis
      Pt *tmp_c, *tmp_d;
      ZKB_TRY(ws.alloc(&tmp_c, (size_t)g.n_sets * (g.B / 4 + 1)));
      ZKB_TRY(ws.alloc(&tmp_d, (size_t)g.n_sets * (g.B / 16 + 1)));
      struct Axis { uint32_t len, other; const Pt* in; size_t in_set_stride; Pt* bufs[2]; Pt* final_out; int flip; bool rows; };
      Axis ax[2] = {{L, H, bucket_acc, g.B, {tmp_a, tmp_b}, rows, 0, true}, {H, L, bucket_acc, g.B, {tmp_c, tmp_d}, cols, 0, false}};
      while (ax[0].len > 1 || ax[1].len > 1) {
        SegPass<F> ps[2];
        unsigned blocks = 1;
        for (int z = 0; z < 2; z++) {
          Axis& a = ax[z];
          SegPass<F>& sp = ps[z];
          memset(&sp, 0, sizeof sp);
          if (a.len <= 1) continue;                       // this axis is done: n_out == 0
          const uint32_t K = a.len >= 4 ? 4u : a.len;
          const uint32_t new_len = a.len / K, n_out = new_len * a.other;
          Pt* out = new_len == 1 ? a.final_out : a.bufs[a.flip];
          sp.in = a.in; sp.in_set_stride = a.in_set_stride; sp.out = out; sp.n_out = n_out; sp.K = K;
          if (a.rows) { sp.inner = new_len; sp.inner_stride = K; sp.outer_stride = a.len; sp.step = 1u; }
          else { sp.inner = a.other; sp.inner_stride = 1u; sp.outer_stride = K * a.other; sp.step = a.other; }
          a.in = out; a.in_set_stride = n_out; a.len = new_len; a.flip ^= 1;
          if (ceil_div(n_out, 128) > blocks) blocks = ceil_div(n_out, 128);
        }
        ZKB_LAUNCH(ctx, (k_seg_sum2<F>), dim3(blocks, g.n_sets, 2), 128, 0, st, ps[0], ps[1]);
      }
    } else {
      ZKB_TRY(reduce_axis(true, rows));
      ZKB_TRY(reduce_axis(false, cols));
    }
    ZKB_LAUNCH(ctx, (k_weight_rows_cols<F>), dim3(ceil_div(H + L, 128), g.n_sets), 128, 0, st, (const Pt*)rows,
               (const Pt*)cols, H, L, weighted);
    uint32_t n_per = H + L;
    Pt *pin = weighted, *pout = tmp_a;
    while (n_per > 1) {
      uint32_t per_thread = n_per > 64 * 64 ? 4 : 1;
      uint32_t n_blocks = ceil_div(n_per, 64 * per_thread);
      ZKB_LAUNCH(ctx, (k_sum_points<F>), dim3(n_blocks, g.n_sets), 64, 0, st, pin, n_per, pout, per_thread);
      pin = pout;
      pout = pout == tmp_a ? tmp_b : tmp_a;
      n_per = n_blocks;
    }
    if (g.n_sets > 1) {
      ZKB_LAUNCH(ctx, (k_window_combine<F>), 1, 32, 0, st, pin, g.n_sets, g.c, d_result);
    } else {
      ZKB_CUDA(ctx, cudaMemcpyAsync(d_result, pin, sizeof(Pt), cudaMemcpyDeviceToDevice, st));
    }
    return ZKB_OK;
    };   // phase2
    if (deferred) {
      *deferred = phase2;
      return ZKB_OK;
    }
    return phase2();
  }

  // MSM with host output (affine canonical)
  static int run_to_host(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint32_t* d_scalars, size_t n,
                         int scalars_mont, uint64_t* out_xy, uint8_t* out_inf) {
    cudaStream_t st = ctx->main;
    Scratch ws(ctx, st);
    Pt* d_pt;
    Aff* d_aff;
    uint8_t* d_inf;
    ZKB_TRY(ws.alloc(&d_pt, 1));
    ZKB_TRY(ws.alloc(&d_aff, 1));
    ZKB_TRY(ws.alloc(&d_inf, 16));
    ZKB_TRY(run(ctx, st, srs, base_offset, d_scalars, n, scalars_mont, d_pt));
    ZKB_LAUNCH(ctx, (k_to_affine<F>), 1, 32, 0, st, d_pt, 1u, d_aff, d_inf);
    ZKB_CUDA(ctx, cudaMemcpyAsync(out_xy, d_aff, sizeof(Aff), cudaMemcpyDeviceToHost, st));
    ZKB_CUDA(ctx, cudaMemcpyAsync(out_inf, d_inf, 1, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA(ctx, cudaStreamSynchronize(st));
    return ZKB_OK;
  }
};

// ------------------------------------------------------------------------------------------
// type-erased per-(curve, group) operations (one translation unit each, see group_*.cu)
// ------------------------------------------------------------------------------------------
// one MSM of a small-batch call (zkb_msm_batch with many short MSMs): bases table[base_offset ..], canonical or
// Montgomery scalars (device), its terms occupy [term0, term0 + len) of the scratch
struct SmallMsmJob {
  const void* table;
  const uint8_t* inf;
  const uint32_t* scalars;
  uint32_t base_offset, len, term0, pad;
};

struct GroupOps {
  size_t affine_bytes, xyzz_bytes;
  int (*srs_build)(zkb_ctx*, zkb_srs*, const void*, const uint8_t*, unsigned);
  int (*msm_run)(zkb_ctx*, cudaStream_t, const zkb_srs*, size_t, const uint32_t*, size_t, int, void*);
  int (*msm_to_host)(zkb_ctx*, const zkb_srs*, size_t, const uint32_t*, size_t, int, uint64_t*, uint8_t*);
  // msm_run in two phases: enqueues the sort now, *deferred enqueues accumulation + reduction when called
  int (*msm_run_split)(zkb_ctx*, cudaStream_t, const zkb_srs*, size_t, const uint32_t*, size_t, int, void*,
                       std::function<int()>* deferred);
  // out = k * P for `count` (scalar, point) pairs, one thread each (small counts: proof assembly)
  int (*fixed_base_mul)(zkb_ctx*, cudaStream_t, const void* d_base_affine, const uint32_t* d_scalars, size_t n,
                        void* d_out_affine, uint8_t* d_out_inf);
  // sum of `count` XYZZ points laid out `stride_bytes` apart (rank order) -> d_out_pt (XYZZ, may be null) and/or
  // canonical affine d_out_affine + d_out_inf (may be null)
  int (*fold)(zkb_ctx*, cudaStream_t, const void* d_points, uint32_t count, uint32_t stride_bytes, void* d_out_pt,
              void* d_out_affine, uint8_t* d_out_inf);
  // ark-serialize compressed points (affine_bytes / 2 bytes each) -> affine Montgomery + identity flags + per-point status
  int (*decompress)(zkb_ctx*, cudaStream_t, const uint8_t* d_in, size_t n, int check_subgroup, void* d_out_affine,
                    uint8_t* d_out_inf, uint8_t* d_out_status);
  // n XYZZ points -> canonical affine + identity flags
  int (*to_affine)(zkb_ctx*, cudaStream_t, const void* d_points, size_t n, void* d_out_affine, uint8_t* d_out_inf);
  // many tiny MSMs at once (a batch verifier's g_ic, one per proof): one thread per (scalar, base) term, one per sum
  int (*small_msms)(zkb_ctx*, cudaStream_t, const SmallMsmJob* d_jobs, uint32_t n_jobs, const uint32_t* d_term_job,
                    uint32_t n_terms, int scalars_mont, void* d_terms_scratch, void* d_out_pts);
};
const GroupOps* group_ops(int curve, int group);

}  // namespace zkb
