// Variable-base multi-scalar multiplication on the device.
//
// Replaces ark_ec::msm::VariableBaseMSM::multi_scalar_mul (ark-ec 0.2, un-vendored; call
// sites groth16/src/prover.rs:187,190,220, marlin/src/pc/kzg10.rs:109,118,137,146,
// curve/src/lib.rs:44).  Same mathematical function sum_i s_i * P_i; the schedule is
// re-designed for B200:
//
//   * signed c-bit digits (half the buckets of the reference's unsigned windows);
//   * with ZKB_SRS_PRECOMPUTE the bases 2^(c*j) * P_i of every window j are resident in
//     HBM (180 GB makes a 16x copy of the SRS cheap), so ALL windows share ONE bucket
//     array: no per-window reduction and no doubling chain at the end;
//   * counting sort of (bucket, base index) pairs: histogram -> exclusive scan -> scatter;
//   * bucket accumulation is chunked over the SORTED list -- every thread adds exactly K
//     consecutive entries whatever the bucket sizes are (boolean-heavy witnesses put
//     millions of entries into one bucket: a thread-per-bucket schedule would serialise);
//     the first partial bucket of a chunk is a "continuation", folded in log_K2 levels;
//   * bucket reduction sum (b+1) * S_b by per-thread running sums over M buckets plus a
//     short scalar multiplication by the chunk offset, then a tree sum.
//
// Entry layout (uint32): bit 31 = negate, bits 0..30 = index into the base table.
#pragma once
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
};

namespace zkb {

constexpr int kAccK = 16;     // sorted entries per accumulate thread
constexpr int kFoldK = 16;    // continuation points per fold thread
constexpr int kRedM = 16;     // buckets per reduce thread
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
    cursor[i] = v;
  }
  if (i == 0) offsets[n] = tile_sums[n_tiles];
}

// ------------------------------------------------------------------------------------------
// bucket accumulation over the sorted entry list, K entries per thread
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_accumulate(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets, uint32_t n_buckets,
             const Affine<F>* __restrict__ table, XYZZ<F>* __restrict__ bucket_acc,
             XYZZ<F>* __restrict__ cont, int32_t* __restrict__ cont_key, uint32_t n_threads) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_threads) return;
  const uint32_t total = offsets[n_buckets];
  uint32_t start = t * kAccK;
  if (start >= total) { cont_key[t] = -1; return; }
  uint32_t end = min(start + (uint32_t)kAccK, total);
  // largest b with offsets[b] <= start
  uint32_t lo = 0, hi = n_buckets;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= start) lo = mid; else hi = mid;
  }
  uint32_t b = lo;
  bool is_cont = offsets[b] < start;
  bool wrote_cont = false;
  uint32_t next = offsets[b + 1];
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t pos = start; pos < end; pos++) {
    if (pos == next) {
      if (is_cont) { st_vec(&cont[t], acc); cont_key[t] = (int32_t)b; wrote_cont = true; is_cont = false; }
      else st_vec(&bucket_acc[b], acc);
      acc = XYZZ<F>::inf();
      do { b++; next = offsets[b + 1]; } while (next <= pos);
    }
    uint32_t e = entries[pos];
    Affine<F> p = ld_vec(&table[e & 0x7fffffffu]);
    acc.madd_xy(p.x, p.y, (e >> 31) != 0);
  }
  if (is_cont) { st_vec(&cont[t], acc); cont_key[t] = (int32_t)b; wrote_cont = true; }
  else st_vec(&bucket_acc[b], acc);
  if (!wrote_cont) cont_key[t] = -1;
}

// fold one level of continuation points (sorted by key, runs are contiguous)
template <class F>
__global__ void __launch_bounds__(128)
k_fold(const XYZZ<F>* __restrict__ in_pts, const int32_t* __restrict__ in_keys, uint32_t n_in,
       XYZZ<F>* __restrict__ out_pts, int32_t* __restrict__ out_keys, XYZZ<F>* __restrict__ bucket_acc,
       uint32_t n_threads) {
  uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_threads) return;
  uint32_t start = u * kFoldK;
  if (start >= n_in) { out_keys[u] = -1; return; }
  uint32_t end = min(start + (uint32_t)kFoldK, n_in);
  int32_t cur = -1;
  bool is_cont = false, wrote_cont = false;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t i = start; i < end; i++) {
    int32_t k = in_keys[i];
    if (k != cur) {
      if (cur >= 0) {
        if (is_cont) { st_vec(&out_pts[u], acc); out_keys[u] = cur; wrote_cont = true; }
        else { XYZZ<F> t = ld_vec_rw(&bucket_acc[cur]); pt_add(t, acc); st_vec(&bucket_acc[cur], t); }
      }
      cur = k;
      acc = XYZZ<F>::inf();
      is_cont = (i == start) && start > 0 && k >= 0 && in_keys[start - 1] == k;
    }
    if (k >= 0) { XYZZ<F> q = ld_vec_rw(&in_pts[i]); pt_add(acc, q); }
  }
  if (cur >= 0) {
    if (is_cont) { st_vec(&out_pts[u], acc); out_keys[u] = cur; wrote_cont = true; }
    else { XYZZ<F> t = ld_vec_rw(&bucket_acc[cur]); pt_add(t, acc); st_vec(&bucket_acc[cur], t); }
  }
  if (!wrote_cont) out_keys[u] = -1;
}

// ------------------------------------------------------------------------------------------
// bucket reduction: out[u] = sum_{b in chunk u} (b + 1) * bucket[b]   (b = index inside its set)
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_bucket_reduce(const XYZZ<F>* __restrict__ bucket_acc, uint32_t B, uint32_t n_sets, XYZZ<F>* __restrict__ out) {
  uint32_t chunks_per_set = (B + kRedM - 1) / kRedM;
  uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= chunks_per_set * n_sets) return;
  uint32_t set = u / chunks_per_set, ch = u % chunks_per_set;
  uint32_t b0 = ch * kRedM;
  uint32_t b1 = min(b0 + (uint32_t)kRedM, B);
  const XYZZ<F>* src = bucket_acc + (size_t)set * B;
  XYZZ<F> running = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  for (uint32_t b = b1; b-- > b0;) {
    XYZZ<F> q = ld_vec_rw(&src[b]);
    pt_add(running, q);
    pt_add(acc, running);
  }
  if (b0 && !running.is_inf()) { XYZZ<F> m = XYZZ<F>::mul_u32(running, b0); pt_add(acc, m); }
  st_vec(&out[u], acc);
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
inline int msm_pick_c(size_t n, int precomp) {
  int l = (int)ceil_log2(n < 2 ? 2 : n);
  int c = precomp ? l : (l > 4 ? l - 3 : 2);     // few bucket sets vs W bucket sets
  if (precomp) { if (c > 20) c = 20; } else { if (c > 16) c = 16; }
  if (c < 4) c = 4;
  if (const char* e = getenv(precomp ? "ZKB_MSM_C" : "ZKB_MSM_C_NOPRE")) {
    int v = atoi(e);
    if (v >= 2 && v <= 23) c = v;
  }
  return c;
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
    srs->c = msm_pick_c(n, srs->precomp);
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
    Pt* d_result = (Pt*)d_result_v;
    if (base_offset + n > srs->n) return set_err(ctx, ZKB_E_INVALID, "msm: base range out of bounds");
    if (n == 0) {
      ZKB_CUDA(ctx, cudaMemsetAsync(d_result, 0, sizeof(Pt), st));
      return ZKB_OK;
    }
    MsmGeom g;
    g.c = srs->c; g.W = srs->W; g.B = 1u << (g.c - 1); g.precomp = srs->precomp;
    g.n_srs = (uint32_t)srs->n; g.n_sets = g.precomp ? 1u : (uint32_t)g.W;
    const uint32_t n_buckets = g.B * g.n_sets;
    if (n_buckets > (1u << 23)) return set_err(ctx, ZKB_E_INVALID, "msm: too many buckets");
    const size_t max_entries = n * (size_t)g.W;
    if (max_entries >= (size_t(1) << 32)) return set_err(ctx, ZKB_E_INVALID, "msm: too many entries");

    Scratch ws(ctx, st);
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

    // accumulate
    Pt *bucket_acc, *cont_a, *cont_b;
    int32_t *key_a, *key_b;
    const uint32_t n_acc_threads = ceil_div(max_entries, kAccK);
    const uint32_t n_fold1 = ceil_div(n_acc_threads, kFoldK);
    ZKB_TRY(ws.alloc(&bucket_acc, n_buckets));
    ZKB_TRY(ws.alloc(&cont_a, n_acc_threads));
    ZKB_TRY(ws.alloc(&key_a, n_acc_threads));
    ZKB_TRY(ws.alloc(&cont_b, n_fold1));
    ZKB_TRY(ws.alloc(&key_b, n_fold1));
    ZKB_CUDA(ctx, cudaMemsetAsync(bucket_acc, 0, sizeof(Pt) * (size_t)n_buckets, st));
    prof_begin(ctx, st);
    ZKB_LAUNCH(ctx, (k_accumulate<F>), ceil_div(n_acc_threads, 128), 128, 0, st, entries, offsets, n_buckets,
               (const Aff*)srs->table, bucket_acc, cont_a, key_a, n_acc_threads);
    prof_end(ctx, st, (double)n * (32.0 + sizeof(Aff)));   // one read of each (scalar, base) pair
    {
      uint32_t n_in = n_acc_threads;
      Pt *pin = cont_a, *pout = cont_b;
      int32_t *kin = key_a, *kout = key_b;
      while (n_in > 1) {
        uint32_t n_out = ceil_div(n_in, kFoldK);
        ZKB_LAUNCH(ctx, (k_fold<F>), ceil_div(n_out, 128), 128, 0, st, pin, kin, n_in, pout, kout, bucket_acc, n_out);
        std::swap(pin, pout);
        std::swap(kin, kout);
        n_in = n_out;
      }
    }
    // reduce
    const uint32_t chunks_per_set = ceil_div(g.B, kRedM);
    Pt *red_a, *red_b;
    ZKB_TRY(ws.alloc(&red_a, (size_t)chunks_per_set * g.n_sets));
    ZKB_TRY(ws.alloc(&red_b, (size_t)ceil_div(chunks_per_set, 64) * g.n_sets + 1));
    ZKB_LAUNCH(ctx, (k_bucket_reduce<F>), ceil_div((size_t)chunks_per_set * g.n_sets, 128), 128, 0, st, bucket_acc,
               g.B, g.n_sets, red_a);
    uint32_t n_per = chunks_per_set;
    Pt *pin = red_a, *pout = red_b;
    while (n_per > 1) {
      uint32_t per_thread = n_per > 64 * 64 ? 4 : 1;
      uint32_t n_blocks = ceil_div(n_per, 64 * per_thread);
      ZKB_LAUNCH(ctx, (k_sum_points<F>), dim3(n_blocks, g.n_sets), 64, 0, st, pin, n_per, pout, per_thread);
      std::swap(pin, pout);
      n_per = n_blocks;
    }
    if (g.n_sets > 1) {
      ZKB_LAUNCH(ctx, (k_window_combine<F>), 1, 32, 0, st, pin, g.n_sets, g.c, d_result);
    } else {
      ZKB_CUDA(ctx, cudaMemcpyAsync(d_result, pin, sizeof(Pt), cudaMemcpyDeviceToDevice, st));
    }
    return ZKB_OK;
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
struct GroupOps {
  size_t affine_bytes, xyzz_bytes;
  int (*srs_build)(zkb_ctx*, zkb_srs*, const void*, const uint8_t*, unsigned);
  int (*msm_run)(zkb_ctx*, cudaStream_t, const zkb_srs*, size_t, const uint32_t*, size_t, int, void*);
  int (*msm_to_host)(zkb_ctx*, const zkb_srs*, size_t, const uint32_t*, size_t, int, uint64_t*, uint8_t*);
  // out = k * P for `count` (scalar, point) pairs, one thread each (small counts: proof assembly)
  int (*fixed_base_mul)(zkb_ctx*, cudaStream_t, const void* d_base_affine, const uint32_t* d_scalars, size_t n,
                        void* d_out_affine, uint8_t* d_out_inf);
};
const GroupOps* group_ops(int curve, int group);

}  // namespace zkb
