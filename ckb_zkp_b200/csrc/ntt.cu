// Radix-2 NTT kernels and domain tables -- see ntt.cuh for the design.
#include "ntt.cuh"

#include "devutil.cuh"

namespace zkb {

// Tile geometry (runtime switches, read once): 2^tile_log elements per shared-memory tile (32 bytes each: 2^10 = 32 KiB,
// 2^11 = 64 KiB of the 227 KiB an SM offers), later passes read runs of 2^min_run_log contiguous elements.
// 2^11-element tiles with 64-byte runs turn a 2^21 transform into 11 + 10 stages = TWO passes over HBM instead of three,
// but measured slower on B200 (2^21 BLS12-381: 0.550 ms vs 0.519 ms, 2^24: 4.77 vs 4.28 ms, profiles/r2_ntt_geometry.txt):
// the transform is bound by the multiplier pipe, not by HBM passes, and the larger tile costs occupancy.  Default 2^10 / 2^2.
struct NttGeom { unsigned tile_log, min_run_log; };
static NttGeom ntt_geom() {
  static const NttGeom g = []() {
    NttGeom r{10u, 2u};
    if (const char* e = getenv("ZKB_NTT_TILE")) { int v = atoi(e); if (v >= 8 && v <= 12) r.tile_log = (unsigned)v; }
    if (const char* e = getenv("ZKB_NTT_RUN")) { int v = atoi(e); if (v >= 0 && v <= 4) r.min_run_log = (unsigned)v; }
    return r;
  }();
  return g;
}

template <class FrP>
__global__ void k_domain_consts(unsigned log_n, Fp<FrP>* c) {
  using Fr = Fp<FrP>;
  if (threadIdx.x | blockIdx.x) return;
  Fr w, g, gi;
#pragma unroll
  for (int i = 0; i < Fr::N; i++) { w.v[i] = FrP::root(i); g.v[i] = FrP::gen(i); gi.v[i] = FrP::gen_inv(i); }
  for (unsigned i = log_n; i < (unsigned)FrP::TWO_ADICITY; i++) w = Fr::sqr(w);
  Fr two = Fr::add(Fr::one(), Fr::one());
  Fr n = Fr::one(), gn = g;
  for (unsigned i = 0; i < log_n; i++) { n = Fr::mul(n, two); gn = Fr::sqr(gn); }
  c[kConstOmega] = w;
  c[kConstOmegaInv] = Fr::inv(w);
  c[kConstNInv] = Fr::inv(n);
  c[kConstG] = g;
  c[kConstGInv] = gi;
  c[kConstZInv] = Fr::inv(Fr::sub(gn, Fr::one()));   // 1 / Z(g), Z = x^n - 1 (r1cs_to_qap.rs:168)
  c[kConstGInvScaled] = Fr::zero();
}

// out[i] = scale * base^i
template <class FrP>
__global__ void k_powers(Fp<FrP>* out, size_t count, const Fp<FrP>* base_p, const Fp<FrP>* scale_p) {
  using Fr = Fp<FrP>;
  constexpr int CH = 32;
  size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * CH;
  if (i0 >= count) return;
  Fr base = *base_p;
  Fr cur = Fr::pow_u64(base, (uint64_t)i0);
  if (scale_p) cur = Fr::mul(cur, *scale_p);
  for (int k = 0; k < CH && i0 + k < count; k++) {
    st_vec(&out[i0 + k], cur);
    cur = Fr::mul(cur, base);
  }
}

// One DIT pass: stages s0 .. s0+k-1 on tiles of 2^(k+t) elements.
//   element e of a tile: lo = e & (2^t - 1), mid = e >> t
//   global index        = (hi << (s0 + k)) | (mid << s0) | (lo_base + lo)
template <class FrP, int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
k_ntt_pass(const Fp<FrP>* src, Fp<FrP>* dst, const Fp<FrP>* __restrict__ tw, unsigned log_n, unsigned s0, unsigned k,
           unsigned t, int first, const Fp<FrP>* __restrict__ pre_scale, const Fp<FrP>* __restrict__ post_scale,
           const Fp<FrP>* __restrict__ post_const, int radix4) {
  using Fr = Fp<FrP>;
  extern __shared__ uint32_t sm[];
  const unsigned E = 1u << (k + t);
  const unsigned lo_groups_log = s0 - t;                       // tiles per hi value = 2^(s0 - t)
  const size_t tile = blockIdx.x;
  const size_t hi = tile >> lo_groups_log;
  const size_t lo_base = (tile & ((size_t(1) << lo_groups_log) - 1)) << t;
  const unsigned lo_mask = (1u << t) - 1;

  for (unsigned e = threadIdx.x; e < E; e += blockDim.x) {
    size_t idx = (hi << (s0 + k)) | ((size_t)(e >> t) << s0) | (lo_base + (e & lo_mask));
    size_t sidx = idx;
    if (first) sidx = (size_t)(__brevll((unsigned long long)idx) >> (64 - log_n));
    Fr v = ld_vec_rw(&src[sidx]);
    if (pre_scale) v = Fr::mul(v, ld_vec(&pre_scale[sidx]));
#pragma unroll
    for (int w = 0; w < Fr::N; w++) sm[w * E + e] = v.v[w];
  }
  __syncthreads();

  // Two stages per shared-memory round trip (radix-4 step: four elements in registers, three twiddles, four
  // multiplications, one barrier) while at least two stages remain; an odd stage count ends with one radix-2 stage.
  unsigned q = 0;
  if (radix4) {
    for (; q + 1 < k; q += 2) {
      for (unsigned i = threadIdx.x; i < E / 4; i += blockDim.x) {
        const unsigned lo = i & lo_mask, m = i >> t;
        const unsigned mid0 = ((m >> q) << (q + 2)) | (m & ((1u << q) - 1));
        const unsigned h = 1u << (q + t);
        const unsigned e0 = (mid0 << t) | lo, e1 = e0 + h, e2 = e0 + 2 * h, e3 = e0 + 3 * h;
        const size_t j = ((size_t)(mid0 & ((1u << q) - 1)) << s0) | (lo_base + lo);
        const unsigned sh = log_n - 1 - (s0 + q);
        const Fr w1 = ld_vec(&tw[j << sh]);                                       // stage q, both pairs
        const Fr w2a = ld_vec(&tw[j << (sh - 1)]);                                // stage q + 1, pair (e0, e2)
        const Fr w2b = ld_vec(&tw[(j + (size_t(1) << (q + s0))) << (sh - 1)]);    // stage q + 1, pair (e1, e3)
        Fr a0, a1, a2, a3;
#pragma unroll
        for (int x = 0; x < Fr::N; x++) { a0.v[x] = sm[x * E + e0]; a1.v[x] = sm[x * E + e1]; a2.v[x] = sm[x * E + e2]; a3.v[x] = sm[x * E + e3]; }
        a1 = Fr::mul(a1, w1);
        a3 = Fr::mul(a3, w1);
        Fr b0 = Fr::add(a0, a1), b1 = Fr::sub(a0, a1), b2 = Fr::add(a2, a3), b3 = Fr::sub(a2, a3);
        b2 = Fr::mul(b2, w2a);
        b3 = Fr::mul(b3, w2b);
        a0 = Fr::add(b0, b2); a2 = Fr::sub(b0, b2); a1 = Fr::add(b1, b3); a3 = Fr::sub(b1, b3);
#pragma unroll
        for (int x = 0; x < Fr::N; x++) { sm[x * E + e0] = a0.v[x]; sm[x * E + e1] = a1.v[x]; sm[x * E + e2] = a2.v[x]; sm[x * E + e3] = a3.v[x]; }
      }
      __syncthreads();
    }
  }
  for (; q < k; q++) {
    for (unsigned i = threadIdx.x; i < E / 2; i += blockDim.x) {
      unsigned lo = i & lo_mask, m = i >> t;
      unsigned mid0 = ((m >> q) << (q + 1)) | (m & ((1u << q) - 1));
      unsigned e0 = (mid0 << t) | lo, e1 = e0 + (1u << (q + t));
      size_t j = ((size_t)(mid0 & ((1u << q) - 1)) << s0) | (lo_base + lo);   // index inside the 2^(s0+q) half-group
      Fr w = ld_vec(&tw[j << (log_n - 1 - (s0 + q))]);
      Fr a, b;
#pragma unroll
      for (int x = 0; x < Fr::N; x++) { a.v[x] = sm[x * E + e0]; b.v[x] = sm[x * E + e1]; }
      b = Fr::mul(b, w);
      Fr s = Fr::add(a, b), d = Fr::sub(a, b);
#pragma unroll
      for (int x = 0; x < Fr::N; x++) { sm[x * E + e0] = s.v[x]; sm[x * E + e1] = d.v[x]; }
    }
    __syncthreads();
  }

  Fr pc;
  if (post_const) pc = *post_const;
  for (unsigned e = threadIdx.x; e < E; e += blockDim.x) {
    size_t idx = (hi << (s0 + k)) | ((size_t)(e >> t) << s0) | (lo_base + (e & lo_mask));
    Fr v;
#pragma unroll
    for (int w = 0; w < Fr::N; w++) v.v[w] = sm[w * E + e];
    if (post_scale) v = Fr::mul(v, ld_vec(&post_scale[idx]));
    if (post_const) v = Fr::mul(v, pc);
    st_vec(&dst[idx], v);
  }
}

template <class FrP>
static int build_domain(zkb_ctx* ctx, NttDomain* d) {
  using Fr = Fp<FrP>;
  cudaStream_t st = ctx->main;
  size_t n = d->n, half = n > 1 ? n / 2 : 1;
  ZKB_CUDA(ctx, cudaMalloc(&d->consts, sizeof(Fr) * kNumConsts));
  ZKB_CUDA(ctx, cudaMalloc(&d->tw, sizeof(Fr) * half));
  ZKB_CUDA(ctx, cudaMalloc(&d->tw_inv, sizeof(Fr) * half));
  ZKB_CUDA(ctx, cudaMalloc(&d->coset, sizeof(Fr) * n));
  ZKB_CUDA(ctx, cudaMalloc(&d->coset_inv, sizeof(Fr) * n));
  Fr* c = (Fr*)d->consts;
  ZKB_LAUNCH(ctx, (k_domain_consts<FrP>), 1, 32, 0, st, d->log_n, c);
  ZKB_LAUNCH(ctx, (k_powers<FrP>), ceil_div(ceil_div(half, 32), 128), 128, 0, st, (Fr*)d->tw, half, c + kConstOmega, (const Fr*)nullptr);
  ZKB_LAUNCH(ctx, (k_powers<FrP>), ceil_div(ceil_div(half, 32), 128), 128, 0, st, (Fr*)d->tw_inv, half, c + kConstOmegaInv, (const Fr*)nullptr);
  ZKB_LAUNCH(ctx, (k_powers<FrP>), ceil_div(ceil_div(n, 32), 128), 128, 0, st, (Fr*)d->coset, n, c + kConstG, (const Fr*)nullptr);
  ZKB_LAUNCH(ctx, (k_powers<FrP>), ceil_div(ceil_div(n, 32), 128), 128, 0, st, (Fr*)d->coset_inv, n, c + kConstGInv, c + kConstNInv);
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

int ntt_get_domain(zkb_ctx* ctx, int curve, unsigned log_n, NttDomain** out) {
  int two_adicity = curve == ZKB_BLS12_381 ? BlsFr::TWO_ADICITY : BnFr::TWO_ADICITY;
  if ((int)log_n > two_adicity) return set_err(ctx, ZKB_E_TOO_LARGE, "domain 2^%u exceeds the field's 2-adicity %d", log_n, two_adicity);
  if (log_n > 30) return set_err(ctx, ZKB_E_TOO_LARGE, "domain 2^%u not supported", log_n);
  int key = curve * 64 + (int)log_n;
  auto it = ctx->domains.find(key);
  if (it != ctx->domains.end()) { *out = it->second; return ZKB_OK; }
  NttDomain* d = new NttDomain();
  d->curve = curve; d->log_n = log_n; d->n = size_t(1) << log_n;
  d->tw = d->tw_inv = d->coset = d->coset_inv = d->consts = nullptr;
  int rc = curve == ZKB_BLS12_381 ? build_domain<BlsFr>(ctx, d) : build_domain<BnFr>(ctx, d);
  if (rc != ZKB_OK) { delete d; return rc; }
  ctx->domains[key] = d;
  *out = d;
  return ZKB_OK;
}

void ntt_free_domains(zkb_ctx* ctx) {
  for (auto& kv : ctx->domains) {
    NttDomain* d = kv.second;
    cudaFree(d->tw); cudaFree(d->tw_inv); cudaFree(d->coset); cudaFree(d->coset_inv); cudaFree(d->consts);
    delete d;
  }
  ctx->domains.clear();
}

template <class FrP>
static int ntt_run_t(zkb_ctx* ctx, cudaStream_t st, NttDomain* dom, void* d_data, void* d_scratch, unsigned flags) {
  using Fr = Fp<FrP>;
  const unsigned log_n = dom->log_n;
  const bool inverse = flags & ZKB_NTT_INVERSE, coset = flags & ZKB_NTT_COSET;
  Fr* data = (Fr*)d_data;
  Fr* scratch = (Fr*)d_scratch;
  const Fr* tw = (const Fr*)(inverse ? dom->tw_inv : dom->tw);
  const Fr* pre = (coset && !inverse) ? (const Fr*)dom->coset : nullptr;
  const Fr* post = (coset && inverse) ? (const Fr*)dom->coset_inv : nullptr;
  const Fr* pconst = (inverse && !coset) ? (const Fr*)dom->consts + kConstNInv : nullptr;
  if (log_n == 0) return ZKB_OK;      // size-1 transform is the identity (g^0 = 1, 1/1 = 1)

  // plan the passes
  const unsigned kTileLog = ntt_geom().tile_log, kMinRunLog = ntt_geom().min_run_log;
  unsigned ks[8], np = 0;
  unsigned k1 = log_n < kTileLog ? log_n : kTileLog;
  ks[np++] = k1;
  unsigned rem = log_n - k1;
  if (rem) {
    unsigned kmax = kTileLog - kMinRunLog;
    unsigned cnt = (rem + kmax - 1) / kmax;
    for (unsigned i = 0; i < cnt; i++) ks[np++] = rem / cnt + (i < rem % cnt ? 1 : 0);
  }
  unsigned s0 = 0;
  for (unsigned p = 0; p < np; p++) {
    unsigned k = ks[p];
    unsigned t = p == 0 ? 0 : kTileLog - k;
    unsigned E = 1u << (k + t);
    const Fr* src = p == 0 ? data : scratch;
    Fr* dst = (p == np - 1 && np > 1) ? data : scratch;
    static const int radix4 = []() { const char* e = getenv("ZKB_NTT_RADIX4"); return e ? atoi(e) : 1; }();
    static const unsigned div = []() { const char* e = getenv("ZKB_NTT_DIV"); unsigned v = e ? (unsigned)atoi(e) : 4u; return v < 2 ? 2u : v; }();
    unsigned threads = E / div < 32 ? 32 : E / div;      // div / 2 butterflies per thread and stage; measured at 2^21: 0.578 / 0.520 / 0.538 ms for div = 2 / 4 / 8
    size_t tiles = dom->n >> (k + t);
    if (threads > 256) threads = 256;                    // __launch_bounds__ of the pass kernel; the loops stride by blockDim
    const size_t smem = sizeof(uint32_t) * Fr::N * E;
    // resident blocks per SM the kernel is compiled for: the radix-4 step wants ~100 registers (2 blocks of 256 threads),
    // capped at 80 / 64 it spills 4 / 40 bytes and 3 / 4 blocks fit.  Measured (fft, 2^21 / 2^24 BLS12-381, ms;
    // profiles/r2q_ntt_radix4_occ.txt): radix-2 only 0.614 / 5.26, 0.551 / 4.58, 0.527 / 4.30 for 2 / 3 / 4 blocks;
    // with radix-4 steps 0.528 / 4.60, 0.505 / 4.21, 0.502 / 4.11 -> default radix-4, 4 blocks
    static const int occ = []() { const char* e = getenv("ZKB_NTT_OCC"); int v = e ? atoi(e) : 4; return v < 2 ? 2 : v > 4 ? 4 : v; }();
    auto launch = [&](auto kernel) -> int {
      if (smem > 48 * 1024)                              // opt-in above 48 KiB is a per-device function attribute: set per call
        ZKB_CUDA(ctx, cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ZKB_LAUNCH(ctx, kernel, (unsigned)tiles, threads, smem, st, src, dst, tw, log_n, s0, k, t, p == 0 ? 1 : 0,
                 p == 0 ? pre : (const Fr*)nullptr, p == np - 1 ? post : (const Fr*)nullptr,
                 p == np - 1 ? pconst : (const Fr*)nullptr, radix4);
      return ZKB_OK;
    };
    if (occ == 2) ZKB_TRY(launch(k_ntt_pass<FrP, 2>));
    else if (occ == 3) ZKB_TRY(launch(k_ntt_pass<FrP, 3>));
    else ZKB_TRY(launch(k_ntt_pass<FrP, 4>));
    s0 += k;
  }
  if (np == 1) ZKB_CUDA(ctx, cudaMemcpyAsync(data, scratch, sizeof(Fr) * dom->n, cudaMemcpyDeviceToDevice, st));
  return ZKB_OK;
}

int ntt_run(zkb_ctx* ctx, cudaStream_t st, NttDomain* dom, void* d_data, void* d_scratch, unsigned flags) {
  return dom->curve == ZKB_BLS12_381 ? ntt_run_t<BlsFr>(ctx, st, dom, d_data, d_scratch, flags)
                                     : ntt_run_t<BnFr>(ctx, st, dom, d_data, d_scratch, flags);
}

}  // namespace zkb
