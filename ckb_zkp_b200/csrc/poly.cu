// Polynomial helpers over Fr used by the Marlin prover around its MSM / NTT calls
// (kernel group K7):
//   * division by (x - z) and evaluation at z   -- KZG10::compute_witness_polynomial
//     (marlin/src/pc/kzg10.rs:211-226) and the evaluations of marlin/src/lib.rs:147-156
//   * linear combination of (shifted) polynomials -- PC::open (marlin/src/pc/mod.rs:85-98)
//   * batch inversion                             -- ark_ff::batch_inversion as used in
//     marlin/src/ahp/prover.rs:365-367 and marlin/src/ahp/arithmetic.rs:28-34
//
// Division by a linear factor is the suffix recurrence H_j = p_j + z * H_(j+1) (q_(j-1) = H_j,
// remainder H_0 = p(z)).  It is evaluated as a tree: an upward pass collapses chunks of 64
// coefficients into their local Horner values with the multiplier z^(64^level), a downward pass
// re-expands the suffix values from the carry of the next chunk -- O(n) work, log_64(n) launches.
#include <cstring>

#include "common.cuh"
#include "devutil.cuh"
#include "field.cuh"

namespace zkb {

constexpr int kPolyChunk = 64;
constexpr int kMaxLevels = 8;
constexpr int kMaxLincomb = 64;

// zp[l] = z^(64^l)
template <class FrP>
__global__ void k_poly_zpows(Fp<FrP> z, int levels, Fp<FrP>* zp) {
  using Fr = Fp<FrP>;
  if (threadIdx.x | blockIdx.x) return;
  Fr cur = z;
  for (int l = 0; l < levels; l++) {
    zp[l] = cur;
    for (int i = 0; i < 6; i++) cur = Fr::sqr(cur);      // ^64
  }
}
// up[c] = sum_i in[c*64 + i] * zl^i
template <class FrP>
__global__ void __launch_bounds__(128)
k_poly_up(const Fp<FrP>* __restrict__ in, size_t n, const Fp<FrP>* __restrict__ zl_p, Fp<FrP>* __restrict__ up,
          size_t n_chunks) {
  using Fr = Fp<FrP>;
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const Fr zl = *zl_p;
  size_t lo = c * kPolyChunk, hi = lo + kPolyChunk < n ? lo + kPolyChunk : n;
  Fr acc = Fr::zero();
  for (size_t j = hi; j-- > lo;) acc = Fr::add(Fr::mul(acc, zl), ld_vec_rw(&in[j]));
  st_vec(&up[c], acc);
}
// suffix values of one level: H[j] = in[j] + zl * H[j+1], the carry into a chunk's top element being
// the suffix value of the next chunk one level up (0 past the end).  At level 0 the values are the
// quotient coefficients shifted by one (q[j-1] = H[j]) and H[0] is the remainder p(z).
template <class FrP>
__global__ void __launch_bounds__(128)
k_poly_down(const Fp<FrP>* __restrict__ in, size_t n, const Fp<FrP>* __restrict__ zl_p, const Fp<FrP>* __restrict__ upper_H,
            size_t n_chunks, Fp<FrP>* __restrict__ H, int level0, Fp<FrP>* __restrict__ q, Fp<FrP>* __restrict__ rem) {
  using Fr = Fp<FrP>;
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const Fr zl = *zl_p;
  size_t lo = c * kPolyChunk, hi = lo + kPolyChunk < n ? lo + kPolyChunk : n;
  Fr carry = (upper_H && c + 1 < n_chunks) ? ld_vec_rw(&upper_H[c + 1]) : Fr::zero();
  for (size_t j = hi; j-- > lo;) {
    carry = Fr::add(Fr::mul(carry, zl), ld_vec_rw(&in[j]));
    if (level0) {
      if (j == 0) { if (rem) st_vec(rem, carry); }
      else if (q) st_vec(&q[j - 1], carry);
    } else {
      st_vec(&H[j], carry);
    }
  }
}

// out[i] = sum_j coeff[j] * poly_j[i - shift_j]
template <class FrP>
struct LincombArgs {
  const Fp<FrP>* poly[kMaxLincomb];
  size_t len[kMaxLincomb];
  size_t shift[kMaxLincomb];
  Fp<FrP> coeff[kMaxLincomb];
  int k;
};
template <class FrP>
__global__ void __launch_bounds__(256)
k_poly_lincomb(const LincombArgs<FrP>* __restrict__ a, Fp<FrP>* __restrict__ out, size_t out_len) {
  using Fr = Fp<FrP>;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < out_len; i += (size_t)gridDim.x * blockDim.x) {
    Fr acc = Fr::zero();
    for (int j = 0; j < a->k; j++) {
      size_t sh = a->shift[j];
      if (i >= sh && i - sh < a->len[j]) acc = Fr::add(acc, Fr::mul(a->coeff[j], ld_vec(&a->poly[j][i - sh])));
    }
    st_vec(&out[i], acc);
  }
}

// batch inversion, one chunk of 64 per thread (Montgomery's trick, zeros are left untouched like
// ark_ff::batch_inversion); prefix products go through `scratch`
template <class FrP>
__global__ void __launch_bounds__(128)
k_batch_inverse(const Fp<FrP>* __restrict__ in, Fp<FrP>* __restrict__ out, Fp<FrP>* __restrict__ scratch, size_t n) {
  using Fr = Fp<FrP>;
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = c * kPolyChunk;
  if (lo >= n) return;
  size_t hi = lo + kPolyChunk < n ? lo + kPolyChunk : n;
  Fr acc = Fr::one();
  for (size_t j = lo; j < hi; j++) {
    Fr v = ld_vec(&in[j]);
    if (!v.is_zero()) acc = Fr::mul(acc, v);
    st_vec(&scratch[j], acc);
  }
  Fr inv = Fr::inv(acc);
  for (size_t j = hi; j-- > lo;) {
    Fr v = ld_vec(&in[j]);
    if (v.is_zero()) { st_vec(&out[j], v); continue; }
    Fr prev = j > lo ? ld_vec_rw(&scratch[j - 1]) : Fr::one();
    st_vec(&out[j], Fr::mul(inv, prev));
    inv = Fr::mul(inv, v);
  }
}

// elementwise vector operations used by the AHP rounds (marlin/src/ahp/prover.rs pointwise loops)
//   0: a + b   1: a - b   2: a * b   3: s * a   4: a + s * b   5: s - a   6: a + s
template <class FrP>
__global__ void __launch_bounds__(256)
k_fr_vec_op(int op, const Fp<FrP>* __restrict__ a, const Fp<FrP>* __restrict__ b, Fp<FrP> s, Fp<FrP>* __restrict__ out, size_t n) {
  using Fr = Fp<FrP>;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr x = ld_vec(&a[i]);
    Fr y = (op == 0 || op == 1 || op == 2 || op == 4) ? ld_vec(&b[i]) : Fr::zero();
    Fr r;
    switch (op) {
      case 0: r = Fr::add(x, y); break;
      case 1: r = Fr::sub(x, y); break;
      case 2: r = Fr::mul(x, y); break;
      case 3: r = Fr::mul(x, s); break;
      case 4: r = Fr::add(x, Fr::mul(y, s)); break;
      case 5: r = Fr::sub(s, x); break;
      default: r = Fr::add(x, s); break;
    }
    st_vec(&out[i], r);
  }
}
// out[i] = scale * base^i
template <class FrP>
__global__ void __launch_bounds__(128)
k_fr_powers(Fp<FrP> base, Fp<FrP> scale, Fp<FrP>* __restrict__ out, size_t n) {
  using Fr = Fp<FrP>;
  constexpr int CH = 32;
  size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * CH;
  if (i0 >= n) return;
  Fr cur = Fr::mul(Fr::pow_u64(base, (uint64_t)i0), scale);
  for (int k = 0; k < CH && i0 + k < n; k++) {
    st_vec(&out[i0 + k], cur);
    cur = Fr::mul(cur, base);
  }
}
// y[i] = sum_p coeff[p] * x[col[p]] over row i (generic sparse matrix-vector product)
// y = M x, one thread per row; rows longer than kSpmvLong (Marlin's transposed matrices have one row per
// VARIABLE, and the constant ONE appears in every other constraint) are deferred to k_spmv_long
constexpr uint32_t kSpmvLong = 256;
template <class FrP>
__global__ void __launch_bounds__(256)
k_spmv_generic(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col_idx, const Fp<FrP>* __restrict__ coeff,
               const Fp<FrP>* __restrict__ x, Fp<FrP>* __restrict__ y, uint32_t n_rows, uint32_t* __restrict__ long_rows,
               uint32_t* __restrict__ n_long) {
  using Fr = Fp<FrP>;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const uint32_t p0 = row_ptr[i], p1 = row_ptr[i + 1];
  if (p1 - p0 > kSpmvLong) {
    long_rows[atomicAdd(n_long, 1u)] = i;
    return;
  }
  Fr acc = Fr::zero();
  const Fr one = Fr::one();
  for (uint32_t p = p0; p < p1; p++) {
    Fr c = ld_vec(&coeff[p]);
    Fr v = ld_vec(&x[col_idx[p]]);
    if (c != one) v = Fr::mul(v, c);
    acc = Fr::add(acc, v);
  }
  st_vec(&y[i], acc);
}
// one block per long row: strided partial sums, tree reduction in shared memory
template <class FrP>
__global__ void __launch_bounds__(256)
k_spmv_long(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col_idx, const Fp<FrP>* __restrict__ coeff,
            const Fp<FrP>* __restrict__ x, Fp<FrP>* __restrict__ y, const uint32_t* __restrict__ long_rows,
            const uint32_t* __restrict__ n_long) {
  using Fr = Fp<FrP>;
  __shared__ Fr sh[256];
  const Fr one = Fr::one();
  for (uint32_t j = blockIdx.x; j < *n_long; j += gridDim.x) {
    const uint32_t i = long_rows[j];
    Fr acc = Fr::zero();
    for (uint32_t p = row_ptr[i] + threadIdx.x; p < row_ptr[i + 1]; p += blockDim.x) {
      Fr c = ld_vec(&coeff[p]);
      Fr v = ld_vec(&x[col_idx[p]]);
      if (c != one) v = Fr::mul(v, c);
      acc = Fr::add(acc, v);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s_ = 128; s_ > 0; s_ >>= 1) {
      if ((int)threadIdx.x < s_) sh[threadIdx.x] = Fr::add(sh[threadIdx.x], sh[threadIdx.x + s_]);
      __syncthreads();
    }
    if (threadIdx.x == 0) st_vec(&y[i], sh[0]);
    __syncthreads();
  }
}

// ---- host-or-device buffers ------------------------------------------------------------------------------------
// A caller that keeps its vectors resident in HBM (the Marlin / PLONK host layers) passes device pointers: those are
// used IN PLACE (no staging copy) and a call whose outputs all live on the device returns without synchronising -- the
// work is ordered on the library's stream, which the resident host layers share (marlin.Ops).  Host buffers are staged
// through stream-ordered scratch as before and the call returns when the result has landed.
static bool on_device(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice;
}
// device view of `count` elements at `src`: the pointer itself, or a staged copy
template <class T>
static int stage_in(zkb_ctx* ctx, Scratch& ws, cudaStream_t st, const void* src, size_t count, const T** out) {
  if (count == 0 || on_device(src)) { *out = (const T*)src; return ZKB_OK; }
  T* d;
  ZKB_TRY(ws.alloc(&d, count));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyDefault, st));
  *out = d;
  return ZKB_OK;
}
// device buffer a kernel writes `count` elements to: `dst` itself when it is device memory, else scratch (*staged = true)
template <class T>
static int stage_out(zkb_ctx* ctx, Scratch& ws, void* dst, size_t count, T** out, bool* staged) {
  (void)ctx;
  if (on_device(dst)) { *out = (T*)dst; *staged = false; return ZKB_OK; }
  *staged = true;
  return ws.alloc(out, count ? count : 1);
}
// copy a staged result back and wait; a device-resident result needs neither
template <class T>
static int finish_out(zkb_ctx* ctx, cudaStream_t st, void* dst, const T* d, size_t count, bool staged) {
  if (!staged) return ZKB_OK;
  if (count) ZKB_CUDA(ctx, cudaMemcpyAsync(dst, d, count * sizeof(T), cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

template <class FrP>
static int vec_op_t(zkb_ctx* ctx, cudaStream_t st, int op, const uint64_t* a, const uint64_t* b, const uint64_t* s_host,
                    uint64_t* out, size_t n) {
  using Fr = Fp<FrP>;
  Scratch ws(ctx, st);
  const bool binary = op == 0 || op == 1 || op == 2 || op == 4;
  const Fr *d_a, *d_b = nullptr;
  Fr* d_o;
  bool staged;
  ZKB_TRY(stage_in(ctx, ws, st, a, n, &d_a));
  if (binary) ZKB_TRY(stage_in(ctx, ws, st, b, n, &d_b));
  ZKB_TRY(stage_out(ctx, ws, out, n, &d_o, &staged));
  Fr s = Fr::zero();
  if (s_host) memcpy(s.v, s_host, 32);
  unsigned blocks = ceil_div(n, 256);
  if (blocks > (unsigned)ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  ZKB_LAUNCH(ctx, (k_fr_vec_op<FrP>), blocks, 256, 0, st, op, d_a, d_b, s, d_o, n);
  return finish_out(ctx, st, out, d_o, n, staged);
}
template <class FrP>
static int powers_t(zkb_ctx* ctx, cudaStream_t st, const uint64_t* base, const uint64_t* scale, uint64_t* out, size_t n) {
  using Fr = Fp<FrP>;
  Scratch ws(ctx, st);
  Fr* d_o;
  bool staged;
  ZKB_TRY(stage_out(ctx, ws, out, n, &d_o, &staged));
  Fr b, sc = Fr::one();
  memcpy(b.v, base, 32);
  if (scale) memcpy(sc.v, scale, 32);
  ZKB_LAUNCH(ctx, (k_fr_powers<FrP>), ceil_div(ceil_div(n, 32), 128), 128, 0, st, b, sc, d_o, n);
  return finish_out(ctx, st, out, d_o, n, staged);
}
template <class FrP>
static int spmv_t(zkb_ctx* ctx, cudaStream_t st, const zkb_csr* m, const uint64_t* x, size_t n_cols, uint64_t* y) {
  using Fr = Fp<FrP>;
  Scratch ws(ctx, st);
  const uint32_t *d_ptr, *d_col;
  const Fr *d_coeff, *d_x;
  Fr* d_y;
  bool staged;
  ZKB_TRY(stage_in(ctx, ws, st, m->row_ptr, m->n_rows + 1, &d_ptr));
  ZKB_TRY(stage_in(ctx, ws, st, m->col_idx, m->nnz, &d_col));
  ZKB_TRY(stage_in(ctx, ws, st, m->coeff_mont, m->nnz, &d_coeff));
  ZKB_TRY(stage_in(ctx, ws, st, x, n_cols, &d_x));
  ZKB_TRY(stage_out(ctx, ws, y, m->n_rows, &d_y, &staged));
  const size_t max_long = m->nnz / kSpmvLong + 1;
  uint32_t *d_long, *d_n_long;
  ZKB_TRY(ws.alloc(&d_long, max_long));
  ZKB_TRY(ws.alloc(&d_n_long, 1));
  ZKB_CUDA(ctx, cudaMemsetAsync(d_n_long, 0, 4, st));
  ZKB_LAUNCH(ctx, (k_spmv_generic<FrP>), ceil_div(m->n_rows, 256), 256, 0, st, d_ptr, d_col, d_coeff, d_x, d_y,
             (uint32_t)m->n_rows, d_long, d_n_long);
  unsigned long_blocks = max_long < (size_t)ctx->sm_count * 4 ? (unsigned)max_long : (unsigned)ctx->sm_count * 4;
  ZKB_LAUNCH(ctx, (k_spmv_long<FrP>), long_blocks, 256, 0, st, d_ptr, d_col, d_coeff, d_x, d_y, (const uint32_t*)d_long,
             (const uint32_t*)d_n_long);
  return finish_out(ctx, st, y, d_y, m->n_rows, staged);
}

template <class FrP>
static int div_linear_t(zkb_ctx* ctx, cudaStream_t st, const void* d_p_v, size_t n, const uint64_t* z_host, void* d_q_v,
                        void* d_rem_v) {
  using Fr = Fp<FrP>;
  const Fr* d_p = (const Fr*)d_p_v;
  Fr* d_q = (Fr*)d_q_v;
  Fr* d_rem = (Fr*)d_rem_v;
  if (n == 0) {
    ZKB_CUDA(ctx, cudaMemsetAsync(d_rem, 0, sizeof(Fr), st));
    return ZKB_OK;
  }
  Fr z;
  memcpy(z.v, z_host, 32);
  Scratch ws(ctx, st);
  size_t sizes[kMaxLevels + 1];
  int levels = 0;
  sizes[0] = n;
  while (sizes[levels] > 1 && levels < kMaxLevels) {
    sizes[levels + 1] = (sizes[levels] + kPolyChunk - 1) / kPolyChunk;
    levels++;
  }
  if (levels == 0) levels = 1, sizes[1] = 1;      // n == 1: one chunk
  Fr* zp;
  ZKB_TRY(ws.alloc(&zp, kMaxLevels));
  ZKB_LAUNCH(ctx, (k_poly_zpows<FrP>), 1, 32, 0, st, z, levels, zp);
  // up[l] = collapsed values feeding level l (up[0] = p); H[l] = suffix values of level l (l >= 1)
  const Fr* up[kMaxLevels + 1];
  Fr* Hs[kMaxLevels + 1];
  up[0] = d_p;
  for (int l = 1; l <= levels; l++) {
    Fr* buf;
    ZKB_TRY(ws.alloc(&buf, sizes[l]));
    ZKB_LAUNCH(ctx, (k_poly_up<FrP>), ceil_div(sizes[l], 128), 128, 0, st, up[l - 1], sizes[l - 1], zp + (l - 1), buf, sizes[l]);
    up[l] = buf;
    ZKB_TRY(ws.alloc(&Hs[l], sizes[l]));
  }
  // downward: level `levels - 1` has a single chunk feeding from nothing, ..., level 0 writes q / rem
  for (int l = levels - 1; l >= 0; l--) {
    const Fr* upper = (l + 1 <= levels - 1) ? Hs[l + 1] : nullptr;     // suffix values of the chunks of level l
    ZKB_LAUNCH(ctx, (k_poly_down<FrP>), ceil_div(sizes[l + 1], 128), 128, 0, st, up[l], sizes[l], zp + l, upper, sizes[l + 1],
               l ? Hs[l] : (Fr*)nullptr, l == 0 ? 1 : 0, d_q, d_rem);
  }
  return ZKB_OK;
}

int poly_div_linear_dev(zkb_ctx* ctx, cudaStream_t st, int curve, const void* d_p, size_t n, const uint64_t* z_mont,
                        void* d_q, void* d_rem) {
  return curve == ZKB_BLS12_381 ? div_linear_t<BlsFr>(ctx, st, d_p, n, z_mont, d_q, d_rem)
                                : div_linear_t<BnFr>(ctx, st, d_p, n, z_mont, d_q, d_rem);
}

template <class FrP>
static int lincomb_t(zkb_ctx* ctx, cudaStream_t st, size_t k, const uint64_t* const* polys, const size_t* lens,
                     const size_t* shifts, const uint64_t* coeffs, uint64_t* out, size_t out_len) {
  using Fr = Fp<FrP>;
  Scratch ws(ctx, st);
  LincombArgs<FrP> h;
  memset(&h, 0, sizeof h);
  h.k = (int)k;
  for (size_t j = 0; j < k; j++) {
    const Fr* d;
    ZKB_TRY(stage_in(ctx, ws, st, polys[j], lens[j], &d));
    h.poly[j] = d;
    h.len[j] = lens[j];
    h.shift[j] = shifts ? shifts[j] : 0;
    memcpy(h.coeff[j].v, coeffs + 4 * j, 32);
  }
  LincombArgs<FrP>* d_args;
  Fr* d_out;
  bool staged;
  ZKB_TRY(ws.alloc(&d_args, 1));
  ZKB_TRY(stage_out(ctx, ws, out, out_len, &d_out, &staged));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_args, &h, sizeof h, cudaMemcpyDefault, st));   // pageable source: staged before the call returns
  unsigned blocks = ceil_div(out_len, 256);
  if (blocks > (unsigned)ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  ZKB_LAUNCH(ctx, (k_poly_lincomb<FrP>), blocks, 256, 0, st, (const LincombArgs<FrP>*)d_args, d_out, out_len);
  return finish_out(ctx, st, out, d_out, out_len, staged);
}

template <class FrP>
static int batch_inverse_t(zkb_ctx* ctx, cudaStream_t st, const uint64_t* in, uint64_t* out, size_t n) {
  using Fr = Fp<FrP>;
  Scratch ws(ctx, st);
  const Fr* d_in;
  Fr *d_out, *d_scr;
  bool staged;
  ZKB_TRY(stage_in(ctx, ws, st, in, n, &d_in));
  ZKB_TRY(stage_out(ctx, ws, out, n, &d_out, &staged));
  if (!staged && (const void*)d_out == (const void*)d_in) {      // in place on a device vector: the kernel reads `in` after writing `out`
    staged = false;
    Fr* copy;
    ZKB_TRY(ws.alloc(&copy, n));
    ZKB_CUDA(ctx, cudaMemcpyAsync(copy, d_in, n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    d_in = copy;
  }
  ZKB_TRY(ws.alloc(&d_scr, n));
  ZKB_LAUNCH(ctx, (k_batch_inverse<FrP>), ceil_div(ceil_div(n, kPolyChunk), 128), 128, 0, st, d_in, d_out, d_scr, n);
  return finish_out(ctx, st, out, d_out, n, staged);
}

// ---- exclusive prefix products: out[i] = in[0] * ... * in[i - 1], out[0] = 1 -----------------------------------
// (the grand-product accumulator z of PLONK's permutation argument, plonk/src/ahp/indexer/permutation.rs:111-118)
// Three phases over chunks of kScanChunk consecutive elements: chunk products, one block scans them, chunks re-walked.
constexpr int kScanChunk = 128;
template <class FrP>
__global__ void k_prefix_chunk_products(const Fp<FrP>* __restrict__ in, size_t n, Fp<FrP>* __restrict__ partial) {
  using Fr = Fp<FrP>;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * kScanChunk;
  if (lo >= n) return;
  size_t hi = lo + kScanChunk < n ? lo + kScanChunk : n;
  Fr acc = ld_vec(&in[lo]);
  for (size_t i = lo + 1; i < hi; i++) acc = Fr::mul(acc, ld_vec(&in[i]));
  st_vec(&partial[t], acc);
}
// one block: partial[j] <- product of the chunk products before chunk j (exclusive)
template <class FrP>
__global__ void __launch_bounds__(1024) k_prefix_scan_partials(Fp<FrP>* partial, size_t m) {
  using Fr = Fp<FrP>;
  __shared__ Fr sh[1024];
  const size_t per = (m + blockDim.x - 1) / blockDim.x;
  const size_t lo = threadIdx.x * per, hi = lo + per < m ? lo + per : m;
  Fr total = Fr::one();
  for (size_t i = lo; i < hi; i++) total = Fr::mul(total, ld_vec_rw(&partial[i]));
  sh[threadIdx.x] = total;
  __syncthreads();
  for (unsigned off = 1; off < blockDim.x; off <<= 1) {           // inclusive Hillis-Steele scan of the per-thread totals
    Fr v = sh[threadIdx.x];
    const bool take = threadIdx.x >= off;
    Fr u = take ? sh[threadIdx.x - off] : Fr::one();
    __syncthreads();
    if (take) sh[threadIdx.x] = Fr::mul(u, v);
    __syncthreads();
  }
  Fr run = threadIdx.x ? sh[threadIdx.x - 1] : Fr::one();
  for (size_t i = lo; i < hi; i++) {
    Fr v = ld_vec_rw(&partial[i]);
    st_vec(&partial[i], run);
    run = Fr::mul(run, v);
  }
}
template <class FrP>
__global__ void k_prefix_apply(const Fp<FrP>* __restrict__ in, size_t n, const Fp<FrP>* __restrict__ partial,
                               Fp<FrP>* __restrict__ out) {
  using Fr = Fp<FrP>;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * kScanChunk;
  if (lo >= n) return;
  size_t hi = lo + kScanChunk < n ? lo + kScanChunk : n;
  Fr run = ld_vec(&partial[t]);
  for (size_t i = lo; i < hi; i++) {
    Fr v = ld_vec(&in[i]);
    st_vec(&out[i], run);
    run = Fr::mul(run, v);
  }
}
template <class FrP>
static int prefix_product_t(zkb_ctx* ctx, cudaStream_t st, const uint64_t* in, uint64_t* out, size_t n) {
  using Fr = Fp<FrP>;
  Scratch ws(ctx, st);
  const Fr* d_in;
  Fr *d_out, *d_part;
  bool staged;
  const size_t m = (n + kScanChunk - 1) / kScanChunk;
  ZKB_TRY(stage_in(ctx, ws, st, in, n, &d_in));
  ZKB_TRY(stage_out(ctx, ws, out, n, &d_out, &staged));
  if (!staged && (const void*)d_out == (const void*)d_in) {      // in place: k_prefix_apply reads in[i] before it writes out[i], element by element
    // (safe as is: each thread owns its chunk and reads an element before overwriting it)
  }
  ZKB_TRY(ws.alloc(&d_part, m));
  ZKB_LAUNCH(ctx, (k_prefix_chunk_products<FrP>), ceil_div(m, 128), 128, 0, st, d_in, n, d_part);
  ZKB_LAUNCH(ctx, (k_prefix_scan_partials<FrP>), 1, 1024, 0, st, d_part, m);
  ZKB_LAUNCH(ctx, (k_prefix_apply<FrP>), ceil_div(m, 128), 128, 0, st, d_in, n, (const Fr*)d_part, d_out);
  return finish_out(ctx, st, out, d_out, n, staged);
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_fr_prefix_product(zkb_ctx* ctx, int curve, const uint64_t* in_mont, uint64_t* out_mont, size_t n) {
  if (!ctx || (n && (!in_mont || !out_mont))) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "prefix_product: unknown curve %d", curve);
  if (n == 0) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return curve == ZKB_BLS12_381 ? prefix_product_t<BlsFr>(ctx, ctx->main, in_mont, out_mont, n)
                                : prefix_product_t<BnFr>(ctx, ctx->main, in_mont, out_mont, n);
}

int zkb_poly_div_linear(zkb_ctx* ctx, int curve, const uint64_t* p_mont, size_t n, const uint64_t z_mont[4], uint64_t* q_mont,
                        uint64_t rem_mont[4]) {
  if (!ctx || (n && !p_mont) || !z_mont || !rem_mont) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "poly_div_linear: unknown curve %d", curve);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  const uint32_t* d_p;
  uint32_t *d_q = nullptr, *d_rem;
  bool q_staged = false;
  ZKB_TRY(stage_in(ctx, ws, st, p_mont, n * 8, &d_p));
  if (q_mont) ZKB_TRY(stage_out(ctx, ws, q_mont, n * 8, &d_q, &q_staged));
  ZKB_TRY(ws.alloc(&d_rem, 8));
  ZKB_TRY(poly_div_linear_dev(ctx, st, curve, d_p, n, z_mont, d_q, d_rem));
  if (q_mont && q_staged && n > 1) ZKB_CUDA(ctx, cudaMemcpyAsync(q_mont, d_q, (n - 1) * 32, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(rem_mont, d_rem, 32, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));          // the remainder is a host result
  return ZKB_OK;
}

int zkb_poly_eval_batch(zkb_ctx* ctx, int curve, size_t k, const uint64_t* const* polys_mont, const size_t* lens,
                        const uint64_t* points_mont, uint64_t* out_mont) {
  if (!ctx || (k && (!polys_mont || !lens || !points_mont || !out_mont))) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "poly_eval_batch: unknown curve %d", curve);
  if (k == 0) return ZKB_OK;
  if (k > 4096) return set_err(ctx, ZKB_E_INVALID, "poly_eval_batch: too many polynomials in one call");
  for (size_t j = 0; j < k; j++)
    if (lens[j] && !polys_mont[j]) return set_err(ctx, ZKB_E_INVALID, "poly_eval_batch: null polynomial %zu", j);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  uint32_t* d_rem;
  ZKB_TRY(ws.alloc(&d_rem, 8 * k));
  // the k remainder trees are enqueued back to back; ONE copy and ONE synchronisation fetch all values
  for (size_t j = 0; j < k; j++) {
    const uint32_t* d_p;
    ZKB_TRY(stage_in(ctx, ws, st, polys_mont[j], lens[j] * 8, &d_p));
    ZKB_TRY(poly_div_linear_dev(ctx, st, curve, d_p, lens[j], points_mont + 4 * j, nullptr, d_rem + 8 * j));
  }
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_mont, d_rem, 32 * k, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

int zkb_poly_lincomb(zkb_ctx* ctx, int curve, size_t k, const uint64_t* const* polys_mont, const size_t* lens,
                     const size_t* shifts, const uint64_t* coeffs_mont, uint64_t* out_mont, size_t out_len) {
  if (!ctx || (k && (!polys_mont || !lens || !coeffs_mont)) || (out_len && !out_mont)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "poly_lincomb: unknown curve %d", curve);
  if (k > (size_t)kMaxLincomb) return set_err(ctx, ZKB_E_INVALID, "poly_lincomb: at most %d polynomials", kMaxLincomb);
  if (out_len == 0) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return curve == ZKB_BLS12_381 ? lincomb_t<BlsFr>(ctx, ctx->main, k, polys_mont, lens, shifts, coeffs_mont, out_mont, out_len)
                                : lincomb_t<BnFr>(ctx, ctx->main, k, polys_mont, lens, shifts, coeffs_mont, out_mont, out_len);
}

int zkb_fr_vec_op(zkb_ctx* ctx, int curve, int op, const uint64_t* a_mont, const uint64_t* b_mont, const uint64_t* s_mont,
                  uint64_t* out_mont, size_t n) {
  if (!ctx || op < 0 || op > 6 || (n && (!a_mont || !out_mont))) return ZKB_E_INVALID;
  const bool binary = op == 0 || op == 1 || op == 2 || op == 4;
  const bool scalar = op >= 3;
  if (n && ((binary && !b_mont) || (scalar && !s_mont))) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "fr_vec_op: unknown curve %d", curve);
  if (n == 0) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return curve == ZKB_BLS12_381 ? vec_op_t<BlsFr>(ctx, ctx->main, op, a_mont, b_mont, s_mont, out_mont, n)
                                : vec_op_t<BnFr>(ctx, ctx->main, op, a_mont, b_mont, s_mont, out_mont, n);
}

int zkb_fr_powers(zkb_ctx* ctx, int curve, const uint64_t base_mont[4], const uint64_t* scale_mont, uint64_t* out_mont, size_t n) {
  if (!ctx || !base_mont || (n && !out_mont)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "fr_powers: unknown curve %d", curve);
  if (n == 0) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return curve == ZKB_BLS12_381 ? powers_t<BlsFr>(ctx, ctx->main, base_mont, scale_mont, out_mont, n)
                                : powers_t<BnFr>(ctx, ctx->main, base_mont, scale_mont, out_mont, n);
}

int zkb_spmv(zkb_ctx* ctx, int curve, const zkb_csr* m, const uint64_t* x_mont, size_t n_cols, uint64_t* y_mont) {
  if (!ctx || !m || (m->n_rows && (!m->row_ptr || !y_mont)) || (m->nnz && (!m->col_idx || !m->coeff_mont || !x_mont)))
    return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "spmv: unknown curve %d", curve);
  if (m->n_rows == 0) return ZKB_OK;
  if (m->n_rows >= (size_t(1) << 31) || m->nnz >= (size_t(1) << 32)) return set_err(ctx, ZKB_E_INVALID, "spmv: matrix too large");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return curve == ZKB_BLS12_381 ? spmv_t<BlsFr>(ctx, ctx->main, m, x_mont, n_cols, y_mont)
                                : spmv_t<BnFr>(ctx, ctx->main, m, x_mont, n_cols, y_mont);
}

int zkb_fr_batch_inverse(zkb_ctx* ctx, int curve, const uint64_t* in_mont, uint64_t* out_mont, size_t n) {
  if (!ctx || (n && (!in_mont || !out_mont))) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (curve != ZKB_BN254 && curve != ZKB_BLS12_381) return set_err(ctx, ZKB_E_INVALID, "batch_inverse: unknown curve %d", curve);
  if (n == 0) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return curve == ZKB_BLS12_381 ? batch_inverse_t<BlsFr>(ctx, ctx->main, in_mont, out_mont, n)
                                : batch_inverse_t<BnFr>(ctx, ctx->main, in_mont, out_mont, n);
}

}  // extern "C"
