// zkref -- CPU restatement of the reference's prove path (TEST INFRASTRUCTURE / CPU BASELINE ONLY).
//
// Nothing in the product (ckb_zkp_b200/, libzkb.so) links or calls this file.  It is used by
// tests/ as the mid-size parity checker and by bench.py's `cpu_baseline` / `--impl reference`
// legs as the "restated reference CPU path (arkworks-0.2 algorithm)".
//
// The reference (sec-bit/ckb-zkp @ 8f2141a) is pure Rust; its arithmetic lives in the
// un-vendored crates ark-ff / ark-ec / ark-poly 0.2 (groth16/Cargo.toml:20-24), and there is no
// Rust toolchain in the build image, so the real prover cannot run here: PARITY UNPINNED at the
// byte level by reference artefacts.  This file restates, with the reference's own schedule:
//   * ark-ff 0.2 Fp256/Fp384: Montgomery form, R = 2^(64*limbs), 64-bit limbs, CIOS multiplication
//   * ark-ec 0.2 short-Weierstrass Jacobian: add_assign_mixed (madd-2007-bl), double_in_place
//     (dbl-2009-l), add_assign (add-2007-bl), into_affine
//   * ark-ec 0.2 VariableBaseMSM::multi_scalar_mul: c = 3 if n < 32 else ln_without_floats(n) + 2,
//     windows processed in parallel, zero scalars skipped, unit scalars added once in window 0,
//     running-sum bucket reduction, high-to-low fold with c doublings
//   * ark-poly 0.2 Radix2EvaluationDomain: bit-reversal + radix-2 DIT butterflies, ifft scaling by
//     size_inv, coset variants through distribute_powers(g)
//   * groth16/src/r1cs_to_qap.rs:15-52,113-172 (evaluate_constraint, witness_map)
//   * groth16/src/prover.rs:124-228 (create_proof, calculate_coeff)
// and is validated against the first-principles Python oracle (oracle/pyref) in tests/test_oracle_c.py.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

typedef uint64_t u64;
typedef unsigned __int128 u128;

// --------------------------------------------------------------------------------------------
// parallel_for: the stand-in for rayon's cfg_iter!/cfg_into_iter! (feature "parallel")
// --------------------------------------------------------------------------------------------
static void parallel_for(size_t n_tasks, int threads, const std::function<void(size_t)>& fn) {
  if (threads <= 1 || n_tasks <= 1) {
    for (size_t i = 0; i < n_tasks; i++) fn(i);
    return;
  }
  std::atomic<size_t> next(0);
  size_t nt = std::min<size_t>(threads, n_tasks);
  std::vector<std::thread> pool;
  for (size_t t = 0; t < nt; t++)
    pool.emplace_back([&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= n_tasks) break;
        fn(i);
      }
    });
  for (auto& th : pool) th.join();
}
// split [0, n) into contiguous chunks, one task per chunk
static void parallel_chunks(size_t n, int threads, const std::function<void(size_t, size_t)>& fn) {
  size_t chunks = threads <= 1 ? 1 : std::min<size_t>((size_t)threads * 4, std::max<size_t>(n / 256, 1));
  size_t per = (n + chunks - 1) / chunks;
  parallel_for(chunks, threads, [&](size_t c) {
    size_t lo = c * per, hi = std::min(n, lo + per);
    if (lo < hi) fn(lo, hi);
  });
}

// --------------------------------------------------------------------------------------------
// prime fields
// --------------------------------------------------------------------------------------------
template <int N>
struct FieldConsts {
  u64 mod[N], one[N], r2[N], pm2[N];
  u64 inv;   // -p^-1 mod 2^64
  int bits;
};

template <int N>
static bool geq(const u64* a, const u64* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
template <int N>
static u64 add_n(u64* r, const u64* a, const u64* b) {
  u128 c = 0;
  for (int i = 0; i < N; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
template <int N>
static u64 sub_n(u64* r, const u64* a, const u64* b) {
  u64 borrow = 0;
  for (int i = 0; i < N; i++) {
    u128 d = (u128)a[i] - b[i] - borrow;
    r[i] = (u64)d;
    borrow = (u64)(d >> 64) & 1;
  }
  return borrow;
}

template <int N>
static void init_consts(FieldConsts<N>& c, const u64* modulus) {
  memcpy(c.mod, modulus, sizeof(c.mod));
  u64 inv = 1;
  for (int i = 0; i < 63; i++) { inv *= inv; inv *= modulus[0]; }   // p^(2^63 - 1) = p^-1 mod 2^64
  c.inv = (u64)0 - inv;
  int bits = 64 * N;
  while (bits > 0 && !((modulus[(bits - 1) / 64] >> ((bits - 1) % 64)) & 1)) bits--;
  c.bits = bits;
  // R mod p and R^2 mod p by repeated doubling
  u64 x[N] = {1};
  for (int i = 0; i < 128 * N; i++) {
    u64 t[N];
    u64 carry = add_n<N>(t, x, x);
    if (carry || geq<N>(t, modulus)) sub_n<N>(t, t, modulus);
    memcpy(x, t, sizeof(x));
    if (i == 64 * N - 1) memcpy(c.one, x, sizeof(x));
  }
  memcpy(c.r2, x, sizeof(x));
  u64 two[N] = {2};
  sub_n<N>(c.pm2, modulus, two);
}

template <int N, int ID>
struct Fp {
  u64 v[N];
  static FieldConsts<N> C;
  static constexpr int LIMBS = N;

  static Fp zero() { Fp r; memset(r.v, 0, sizeof(r.v)); return r; }
  static Fp one() { Fp r; memcpy(r.v, C.one, sizeof(r.v)); return r; }
  bool is_zero() const { u64 t = 0; for (int i = 0; i < N; i++) t |= v[i]; return t == 0; }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  Fp operator+(const Fp& o) const {
    Fp r;
    u64 carry = add_n<N>(r.v, v, o.v);
    if (carry || geq<N>(r.v, C.mod)) sub_n<N>(r.v, r.v, C.mod);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r;
    if (sub_n<N>(r.v, v, o.v)) add_n<N>(r.v, r.v, C.mod);
    return r;
  }
  Fp neg() const { return is_zero() ? *this : zero() - *this; }
  Fp dbl() const { return *this + *this; }
  // CIOS Montgomery multiplication (ark-ff 0.2 `mul_assign` without the no-carry shortcut; same value)
  Fp operator*(const Fp& o) const {
    u64 t[N + 2];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < N; i++) {
      u128 carry = 0;
      for (int j = 0; j < N; j++) {
        u128 cur = (u128)v[j] * o.v[i] + t[j] + carry;
        t[j] = (u64)cur;
        carry = cur >> 64;
      }
      u128 cur = (u128)t[N] + carry;
      t[N] = (u64)cur;
      t[N + 1] = (u64)(cur >> 64);
      u64 m = t[0] * C.inv;
      carry = ((u128)m * C.mod[0] + t[0]) >> 64;
      for (int j = 1; j < N; j++) {
        u128 c2 = (u128)m * C.mod[j] + t[j] + carry;
        t[j - 1] = (u64)c2;
        carry = c2 >> 64;
      }
      cur = (u128)t[N] + carry;
      t[N - 1] = (u64)cur;
      t[N] = t[N + 1] + (u64)(cur >> 64);
    }
    Fp r;
    memcpy(r.v, t, sizeof(r.v));
    if (t[N] || geq<N>(r.v, C.mod)) sub_n<N>(r.v, r.v, C.mod);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp pow(const u64* e, int limbs) const {
    Fp r = one();
    for (int i = limbs - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        r = r.sqr();
        if ((e[i] >> b) & 1) r = r * *this;
      }
    return r;
  }
  Fp pow_u64(u64 e) const { return pow(&e, 1); }
  Fp inv() const { return pow(C.pm2, N); }
  static Fp from_canonical(const u64* x) { Fp r; memcpy(r.v, x, sizeof(r.v)); Fp r2; memcpy(r2.v, C.r2, sizeof(r2.v)); return r * r2; }
  static Fp from_u64(u64 x) { u64 t[N] = {x}; return from_canonical(t); }
  void to_canonical(u64* out) const {   // into_repr
    Fp o = zero();
    o.v[0] = 1;
    Fp r = *this * o;
    memcpy(out, r.v, sizeof(r.v));
  }
};
template <int N, int ID> FieldConsts<N> Fp<N, ID>::C;

template <class B>
struct Fp2 {   // Fq[u]/(u^2 + 1)
  B c0, c1;
  static constexpr int LIMBS = 2 * B::LIMBS;
  static Fp2 zero() { return {B::zero(), B::zero()}; }
  static Fp2 one() { return {B::one(), B::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fp2& o) const { return !(*this == o); }
  Fp2 operator+(const Fp2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fp2 operator-(const Fp2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fp2 neg() const { return {c0.neg(), c1.neg()}; }
  Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fp2 operator*(const Fp2& o) const {   // Karatsuba (ark-ff QuadExtField::mul_assign)
    B v0 = c0 * o.c0, v1 = c1 * o.c1;
    B s = (c0 + c1) * (o.c0 + o.c1);
    return {v0 - v1, s - v0 - v1};
  }
  Fp2 sqr() const {
    B t = c0 * c1;
    return {(c0 + c1) * (c0 - c1), t.dbl()};
  }
  Fp2 inv() const {
    B n = (c0.sqr() + c1.sqr()).inv();
    return {c0 * n, (c1 * n).neg()};
  }
};

typedef Fp<4, 0> BnFr;
typedef Fp<4, 1> BlsFr;
typedef Fp<4, 2> BnFq;
typedef Fp<6, 3> BlsFq;

static const u64 kBlsFr[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const u64 kBnFr[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 kBnFq[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 kBlsFq[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                              0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};

template <class Fr> struct FrInfo;
template <> struct FrInfo<BlsFr> { static constexpr u64 GEN = 7; static constexpr int TWO_ADICITY = 32; };
template <> struct FrInfo<BnFr> { static constexpr u64 GEN = 5; static constexpr int TWO_ADICITY = 28; };

static void init_all() {
  static std::once_flag once;
  std::call_once(once, []() {
    init_consts<4>(BlsFr::C, kBlsFr);
    init_consts<4>(BnFr::C, kBnFr);
    init_consts<4>(BnFq::C, kBnFq);
    init_consts<6>(BlsFq::C, kBlsFq);
  });
}

// --------------------------------------------------------------------------------------------
// short-Weierstrass a = 0 Jacobian arithmetic (ark-ec 0.2 GroupProjective / GroupAffine)
// --------------------------------------------------------------------------------------------
template <class F>
struct Aff {
  F x, y;
  bool inf;
};

template <class F>
struct Jac {
  F X, Y, Z;
  static Jac identity() { return {F::one(), F::one(), F::zero()}; }   // ark: (1, 1, 0)
  bool is_zero() const { return Z.is_zero(); }

  void dbl() {   // double_in_place, a = 0 (dbl-2009-l)
    if (is_zero()) return;
    F A = X.sqr(), B = Y.sqr(), C = B.sqr();
    F D = ((X + B).sqr() - A - C).dbl();
    F E = A + A.dbl();
    F Fv = E.sqr();
    Z = (Z * Y).dbl();
    X = Fv - D.dbl();
    Y = (D - X) * E - C.dbl().dbl().dbl();
  }
  void add_mixed(const Aff<F>& o) {   // add_assign_mixed (madd-2007-bl)
    if (o.inf) return;
    if (is_zero()) { X = o.x; Y = o.y; Z = F::one(); return; }
    F Z1Z1 = Z.sqr();
    F U2 = o.x * Z1Z1;
    F S2 = (o.y * Z) * Z1Z1;
    if (X == U2 && Y == S2) { dbl(); return; }
    F H = U2 - X;
    F HH = H.sqr();
    F I = HH.dbl().dbl();
    F J = H * I;
    F r = (S2 - Y).dbl();
    F V = X * I;
    F X3 = r.sqr() - J - V.dbl();
    F Y3 = r * (V - X3) - (Y * J).dbl();
    Z = (Z + H).sqr() - Z1Z1 - HH;
    X = X3;
    Y = Y3;
  }
  void add(const Jac& o) {   // add_assign (add-2007-bl)
    if (is_zero()) { *this = o; return; }
    if (o.is_zero()) return;
    F Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
    F U1 = X * Z2Z2, U2 = o.X * Z1Z1;
    F S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
    if (U1 == U2 && S1 == S2) { dbl(); return; }
    F H = U2 - U1;
    F I = H.dbl().sqr();
    F J = H * I;
    F r = (S2 - S1).dbl();
    F V = U1 * I;
    F X3 = r.sqr() - J - V.dbl();
    F Y3 = r * (V - X3) - (S1 * J).dbl();
    Z = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
    X = X3;
    Y = Y3;
  }
  void neg() { Y = Y.neg(); }
  Aff<F> to_affine() const {   // into_affine
    if (is_zero()) return {F::zero(), F::one(), true};
    F zi = Z.inv();
    F zi2 = zi.sqr();
    return {X * zi2, Y * zi2 * zi, false};
  }
  // self * k, k canonical little-endian limbs (ark `mul`: double-and-add from the top bit)
  Jac mul(const u64* k, int limbs) const {
    Jac r = identity();
    bool started = false;
    for (int i = limbs - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        if (started) r.dbl();
        if ((k[i] >> b) & 1) { r.add(*this); started = true; }
      }
    return r;
  }
};

template <class F>
static Aff<F> load_affine(const u64* xy, uint8_t inf) {
  Aff<F> p;
  memcpy(&p.x, xy, sizeof(F));
  memcpy(&p.y, xy + F::LIMBS, sizeof(F));
  p.inf = inf != 0;
  return p;
}
template <class F>
static void store_affine(const Aff<F>& p, u64* xy, uint8_t* inf) {
  memcpy(xy, &p.x, sizeof(F));
  memcpy(xy + F::LIMBS, &p.y, sizeof(F));
  *inf = p.inf ? 1 : 0;
}

// ark_std::log2 (ceil) and ark-ec 0.2 ln_without_floats
static unsigned ark_log2(size_t x) {
  if (x <= 1) return 0;
  unsigned l = 0;
  while ((size_t(1) << l) < x) l++;
  return l;
}
static unsigned ark_window(size_t n) { return n < 32 ? 3 : ark_log2(n) * 69 / 100 + 2; }

// VariableBaseMSM::multi_scalar_mul(bases, scalars) -- scalars canonical, 4 limbs
template <class F>
static Jac<F> msm_ark(const u64* bases_xy, const uint8_t* inf, const u64* scalars, size_t n, int num_bits, int threads,
                      int* threads_used) {
  const unsigned c = ark_window(n);
  std::vector<unsigned> window_starts;
  for (unsigned w = 0; w < (unsigned)num_bits; w += c) window_starts.push_back(w);
  std::vector<Jac<F>> window_sums(window_starts.size());
  if (threads_used) *threads_used = (int)std::min<size_t>(std::max(threads, 1), window_starts.size());
  constexpr int W = 2 * F::LIMBS;   // u64 words per affine point
  parallel_for(window_starts.size(), threads, [&](size_t wi) {
    const unsigned w_start = window_starts[wi];
    Jac<F> res = Jac<F>::identity();
    std::vector<Jac<F>> buckets((size_t(1) << c) - 1, Jac<F>::identity());
    for (size_t i = 0; i < n; i++) {
      const u64* s = scalars + 4 * i;
      if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;                  // .filter(|(s, _)| !s.is_zero())
      if (s[0] == 1 && (s[1] | s[2] | s[3]) == 0) {                    // scalar == fr_one
        if (w_start == 0) res.add_mixed(load_affine<F>(bases_xy + W * i, inf[i]));
      } else {
        // scalar.divn(w_start); scalar.as_ref()[0] % (1 << c)
        unsigned limb = w_start / 64, sh = w_start % 64;
        u64 lo = s[limb] >> sh;
        if (sh && limb + 1 < 4) lo |= s[limb + 1] << (64 - sh);
        u64 d = lo & ((u64(1) << c) - 1);
        if (d != 0) buckets[d - 1].add_mixed(load_affine<F>(bases_xy + W * i, inf[i]));
      }
    }
    Jac<F> running = Jac<F>::identity();
    for (size_t b = buckets.size(); b-- > 0;) {
      running.add(buckets[b]);
      res.add(running);
    }
    window_sums[wi] = res;
  });
  // lowest + fold(rev(rest)): total += w; c doublings
  Jac<F> total = Jac<F>::identity();
  for (size_t wi = window_sums.size(); wi-- > 1;) {
    total.add(window_sums[wi]);
    for (unsigned k = 0; k < c; k++) total.dbl();
  }
  Jac<F> lowest = window_sums[0];
  lowest.add(total);
  return lowest;
}

// FixedBaseMSM-equivalent: out[i] = scalars[i] * base via one window table, then batch_normalization
template <class F>
static void fixed_base_mul(const u64* base_xy, const u64* scalars, size_t n, int scalar_bits, int threads, u64* out_xy,
                           uint8_t* out_inf) {
  constexpr int W = 2 * F::LIMBS;
  const unsigned w = n < 256 ? 4 : 8;
  const unsigned n_win = (scalar_bits + w - 1) / w;
  Aff<F> base = load_affine<F>(base_xy, 0);
  // table[j][d] = d * 2^(w*j) * base, affine
  std::vector<std::vector<Aff<F>>> table(n_win);
  Jac<F> cur{base.x, base.y, F::one()};
  for (unsigned j = 0; j < n_win; j++) {
    std::vector<Jac<F>> row(size_t(1) << w, Jac<F>::identity());
    for (size_t d = 1; d < row.size(); d++) { row[d] = row[d - 1]; row[d].add(cur); }
    table[j].resize(row.size());
    for (size_t d = 0; d < row.size(); d++) table[j][d] = row[d].to_affine();
    for (unsigned k = 0; k < w; k++) cur.dbl();
  }
  std::vector<Jac<F>> res(n);
  parallel_chunks(n, threads, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      const u64* s = scalars + 4 * i;
      Jac<F> acc = Jac<F>::identity();
      for (unsigned j = 0; j < n_win; j++) {
        unsigned bit = j * w, limb = bit / 64, sh = bit % 64;
        if (limb >= 4) break;
        u64 lo64 = s[limb] >> sh;
        if (sh && limb + 1 < 4) lo64 |= s[limb + 1] << (64 - sh);
        u64 d = lo64 & ((u64(1) << w) - 1);
        if (d) acc.add_mixed(table[j][d]);
      }
      res[i] = acc;
    }
  });
  // batch normalisation (Montgomery's trick) per chunk
  parallel_chunks(n, threads, [&](size_t lo, size_t hi) {
    std::vector<F> prod(hi - lo);
    F acc = F::one();
    for (size_t i = lo; i < hi; i++) {
      if (!res[i].is_zero()) acc = acc * res[i].Z;
      prod[i - lo] = acc;
    }
    F inv = acc.inv();
    for (size_t i = hi; i-- > lo;) {
      Aff<F> a{F::zero(), F::one(), true};
      if (!res[i].is_zero()) {
        F prev = i > lo ? prod[i - lo - 1] : F::one();
        F zi = inv * prev;
        inv = inv * res[i].Z;
        F zi2 = zi.sqr();
        a = {res[i].X * zi2, res[i].Y * zi2 * zi, false};
      }
      store_affine<F>(a, out_xy + W * i, out_inf + i);
    }
  });
}

// --------------------------------------------------------------------------------------------
// Radix2EvaluationDomain (ark-poly 0.2)
// --------------------------------------------------------------------------------------------
template <class Fr>
struct Domain {
  unsigned log_n;
  size_t n;
  Fr group_gen, group_gen_inv, size_inv, g, g_inv;

  static bool make(size_t min_size, Domain* d) {
    unsigned l = ark_log2(min_size);
    if ((int)l > FrInfo<Fr>::TWO_ADICITY) return false;   // -> SynthesisError::PolynomialDegreeTooLarge
    d->log_n = l;
    d->n = size_t(1) << l;
    Fr gen = Fr::from_u64(FrInfo<Fr>::GEN);
    // two-adic root = g^((p-1)/2^s); then square down to the domain size
    u64 e[4];
    u64 onev[4] = {1, 0, 0, 0};
    sub_n<4>(e, Fr::C.mod, onev);
    for (int i = 0; i < FrInfo<Fr>::TWO_ADICITY; i++) {
      for (int j = 0; j < 3; j++) e[j] = (e[j] >> 1) | (e[j + 1] << 63);
      e[3] >>= 1;
    }
    Fr w = gen.pow(e, 4);
    for (int i = (int)l; i < FrInfo<Fr>::TWO_ADICITY; i++) w = w.sqr();
    d->group_gen = w;
    d->group_gen_inv = w.inv();
    d->size_inv = Fr::from_u64((u64)d->n).inv();
    d->g = gen;
    d->g_inv = gen.inv();
    return true;
  }

  // serial_fft restated; the butterfly loops are spread over `threads` (rayon's parallel_fft
  // computes the same values)
  void fft_core(Fr* a, const Fr& omega, int threads) const {
    for (size_t k = 0; k < n; k++) {
      size_t rk = 0;
      for (unsigned b = 0; b < log_n; b++) rk |= ((k >> b) & 1) << (log_n - 1 - b);
      if (k < rk) std::swap(a[k], a[rk]);
    }
    size_t m = 1;
    for (unsigned s = 0; s < log_n; s++) {
      Fr w_m = omega.pow_u64((u64)(n / (2 * m)));
      size_t blocks = n / (2 * m);
      if (blocks >= (size_t)threads * 4 || threads <= 1) {
        parallel_chunks(blocks, threads, [&](size_t lo, size_t hi) {
          for (size_t blk = lo; blk < hi; blk++) {
            size_t k = blk * 2 * m;
            Fr w = Fr::one();
            for (size_t j = 0; j < m; j++) {
              Fr t = a[k + j + m] * w;
              a[k + j + m] = a[k + j] - t;
              a[k + j] = a[k + j] + t;
              w = w * w_m;
            }
          }
        });
      } else {
        for (size_t blk = 0; blk < blocks; blk++) {
          size_t k = blk * 2 * m;
          parallel_chunks(m, threads, [&](size_t lo, size_t hi) {
            Fr w = w_m.pow_u64((u64)lo);
            for (size_t j = lo; j < hi; j++) {
              Fr t = a[k + j + m] * w;
              a[k + j + m] = a[k + j] - t;
              a[k + j] = a[k + j] + t;
              w = w * w_m;
            }
          });
        }
      }
      m *= 2;
    }
  }
  void distribute_powers(Fr* a, const Fr& gg, int threads) const {
    parallel_chunks(n, threads, [&](size_t lo, size_t hi) {
      Fr p = gg.pow_u64((u64)lo);
      for (size_t i = lo; i < hi; i++) { a[i] = a[i] * p; p = p * gg; }
    });
  }
  void fft(Fr* a, int threads) const { fft_core(a, group_gen, threads); }
  void ifft(Fr* a, int threads) const {
    fft_core(a, group_gen_inv, threads);
    parallel_chunks(n, threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) a[i] = a[i] * size_inv; });
  }
  void coset_fft(Fr* a, int threads) const { distribute_powers(a, g, threads); fft(a, threads); }
  void coset_ifft(Fr* a, int threads) const { ifft(a, threads); distribute_powers(a, g_inv, threads); }
};

// --------------------------------------------------------------------------------------------
// R1CS -> QAP witness map (groth16/src/r1cs_to_qap.rs:15-52,113-172)
// --------------------------------------------------------------------------------------------
struct Csr {
  size_t n_rows, nnz;
  const uint32_t* row_ptr;
  const uint32_t* col_idx;
  const u64* coeff_mont;
};

template <class Fr>
static Fr evaluate_constraint(const Csr& m, size_t row, const Fr* z) {
  Fr acc = Fr::zero();
  const Fr one = Fr::one();
  for (uint32_t p = m.row_ptr[row]; p < m.row_ptr[row + 1]; p++) {
    Fr c;
    memcpy(&c, m.coeff_mont + 4 * (size_t)p, sizeof(Fr));
    Fr v = z[m.col_idx[p]];
    if (c != one) v = v * c;
    acc = acc + v;
  }
  return acc;
}

// returns h (Montgomery) of length domain size; 0 on success, -3 when the domain is too large
template <class Fr>
static int witness_map(const Csr& A, const Csr& B, const Csr& C, const Fr* z, size_t n_inputs, int threads,
                       std::vector<Fr>& h) {
  Domain<Fr> d;
  if (!Domain<Fr>::make(A.n_rows + n_inputs, &d)) return -3;
  const size_t n = d.n, nc = A.n_rows;
  std::vector<Fr> a(n, Fr::zero()), b(n, Fr::zero()), c(n, Fr::zero());
  parallel_chunks(nc, threads, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      a[i] = evaluate_constraint<Fr>(A, i, z);
      b[i] = evaluate_constraint<Fr>(B, i, z);
    }
  });
  for (size_t i = 0; i < n_inputs; i++) a[nc + i] = z[i];
  d.ifft(a.data(), threads);
  d.ifft(b.data(), threads);
  d.coset_fft(a.data(), threads);
  d.coset_fft(b.data(), threads);
  parallel_chunks(n, threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) a[i] = a[i] * b[i]; });
  parallel_chunks(nc, threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) c[i] = evaluate_constraint<Fr>(C, i, z); });
  d.ifft(c.data(), threads);
  d.coset_fft(c.data(), threads);
  Fr zinv = (d.g.pow_u64((u64)n) - Fr::one()).inv();   // evaluate_vanishing_polynomial(g)^-1
  parallel_chunks(n, threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) a[i] = (a[i] - c[i]) * zinv; });
  d.coset_ifft(a.data(), threads);
  h.swap(a);
  return 0;
}

// --------------------------------------------------------------------------------------------
// Groth16 create_proof (groth16/src/prover.rs:124-228)
// --------------------------------------------------------------------------------------------
struct G16Key {
  const u64 *a_xy, *b1_xy, *b2_xy, *h_xy, *l_xy;
  const uint8_t *a_inf, *b1_inf, *b2_inf, *h_inf, *l_inf;
  size_t a_len, b1_len, b2_len, h_len, l_len;
  const u64* g1_singles;   // alpha, beta, delta
  const u64* g2_singles;   // beta, delta
};

template <class G>
static Jac<G> calculate_coeff(const Jac<G>& initial, const u64* q_xy, const uint8_t* q_inf, size_t q_len, const u64* vk_param,
                              const u64* assignment, size_t n_assign, int bits, int threads) {
  constexpr int W = 2 * G::LIMBS;
  size_t n = std::min(q_len ? q_len - 1 : 0, n_assign);
  Jac<G> acc = msm_ark<G>(q_xy + W, q_inf + 1, assignment, n, bits, threads, nullptr);
  Jac<G> res = initial;
  res.add_mixed(load_affine<G>(q_xy, q_inf[0]));
  res.add(acc);
  res.add_mixed(load_affine<G>(vk_param, 0));
  return res;
}

template <class Fr, class Fq>
static int groth16_prove(const G16Key& pk, const Csr& A, const Csr& B, const Csr& C, const u64* z_mont, size_t n_inputs,
                         size_t n_aux, const u64* r, const u64* s, int threads, u64* proof_xy, uint8_t* proof_inf) {
  typedef Fp2<Fq> Fq2;
  constexpr int W1 = 2 * Fq::LIMBS, W2 = 4 * Fq::LIMBS;
  const int bits = Fr::C.bits;
  const Fr* z = reinterpret_cast<const Fr*>(z_mont);
  std::vector<Fr> h;
  int rc = witness_map<Fr>(A, B, C, z, n_inputs, threads, h);
  if (rc) return rc;
  // into_repr sweeps (prover.rs:150-161)
  const size_t n_assign = n_inputs - 1 + n_aux;
  std::vector<u64> assign(4 * n_assign), h_repr(4 * h.size());
  parallel_chunks(n_assign, threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) z[1 + i].to_canonical(&assign[4 * i]); });
  parallel_chunks(h.size(), threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) h[i].to_canonical(&h_repr[4 * i]); });
  const u64* aux_assign = assign.data() + 4 * (n_inputs - 1);

  Aff<Fq> delta_g1 = load_affine<Fq>(pk.g1_singles + 2 * W1, 0);
  Aff<Fq2> delta_g2 = load_affine<Fq2>(pk.g2_singles + W2, 0);
  Jac<Fq> d1{delta_g1.x, delta_g1.y, Fq::one()};
  Jac<Fq2> d2{delta_g2.x, delta_g2.y, Fq2::one()};
  bool r_zero = (r[0] | r[1] | r[2] | r[3]) == 0;

  Jac<Fq> g_a = calculate_coeff<Fq>(d1.mul(r, 4), pk.a_xy, pk.a_inf, pk.a_len, pk.g1_singles, assign.data(), n_assign, bits, threads);
  Jac<Fq> g1_b = Jac<Fq>::identity();
  if (!r_zero)
    g1_b = calculate_coeff<Fq>(d1.mul(s, 4), pk.b1_xy, pk.b1_inf, pk.b1_len, pk.g1_singles + W1, assign.data(), n_assign, bits, threads);
  Jac<Fq2> g2_b = calculate_coeff<Fq2>(d2.mul(s, 4), pk.b2_xy, pk.b2_inf, pk.b2_len, pk.g2_singles, assign.data(), n_assign, bits, threads);
  Jac<Fq> h_acc = msm_ark<Fq>(pk.h_xy, pk.h_inf, h_repr.data(), std::min(pk.h_len, h.size()), bits, threads, nullptr);
  Jac<Fq> l_acc = msm_ark<Fq>(pk.l_xy, pk.l_inf, aux_assign, std::min(pk.l_len, n_aux), bits, threads, nullptr);

  Jac<Fq> s_g_a = g_a.mul(s, 4);
  Jac<Fq> r_g1_b = g1_b.mul(r, 4);
  Jac<Fq> r_s_delta = d1.mul(r, 4).mul(s, 4);
  Jac<Fq> g_c = s_g_a;
  g_c.add(r_g1_b);
  r_s_delta.neg();
  g_c.add(r_s_delta);
  g_c.add(l_acc);
  g_c.add(h_acc);
  store_affine<Fq>(g_a.to_affine(), proof_xy, proof_inf);
  store_affine<Fq2>(g2_b.to_affine(), proof_xy + W1, proof_inf + 1);
  store_affine<Fq>(g_c.to_affine(), proof_xy + W1 + W2, proof_inf + 2);
  return 0;
}

// --------------------------------------------------------------------------------------------
// C entry points (ctypes)
// --------------------------------------------------------------------------------------------
#define BN254 0
#define BLS12_381 1

// flags: 1 = inverse, 2 = coset (same as ZKB_NTT_*); data Montgomery, in place
template <class Fr>
static int ntt_t(u64* data, unsigned log_n, unsigned flags, int threads) {
  Domain<Fr> d;
  if (log_n > 40 || !Domain<Fr>::make(size_t(1) << log_n, &d)) return -3;
  Fr* a = reinterpret_cast<Fr*>(data);
  if (flags == 0) d.fft(a, threads);
  else if (flags == 1) d.ifft(a, threads);
  else if (flags == 2) d.coset_fft(a, threads);
  else d.coset_ifft(a, threads);
  return 0;
}
extern "C" {

int zkref_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

// VariableBaseMSM::multi_scalar_mul; out = canonical affine.  *threads_used <- min(threads, windows)
int zkref_msm(int curve, int group, const u64* bases_xy, const uint8_t* inf, const u64* scalars, size_t n, int threads,
              u64* out_xy, uint8_t* out_inf, int* threads_used) {
  init_all();
  if (curve == BN254 && group == 1) store_affine<BnFq>(msm_ark<BnFq>(bases_xy, inf, scalars, n, BnFr::C.bits, threads, threads_used).to_affine(), out_xy, out_inf);
  else if (curve == BN254 && group == 2) store_affine<Fp2<BnFq>>(msm_ark<Fp2<BnFq>>(bases_xy, inf, scalars, n, BnFr::C.bits, threads, threads_used).to_affine(), out_xy, out_inf);
  else if (curve == BLS12_381 && group == 1) store_affine<BlsFq>(msm_ark<BlsFq>(bases_xy, inf, scalars, n, BlsFr::C.bits, threads, threads_used).to_affine(), out_xy, out_inf);
  else if (curve == BLS12_381 && group == 2) store_affine<Fp2<BlsFq>>(msm_ark<Fp2<BlsFq>>(bases_xy, inf, scalars, n, BlsFr::C.bits, threads, threads_used).to_affine(), out_xy, out_inf);
  else return -1;
  return 0;
}

int zkref_fixed_base_mul(int curve, int group, const u64* base_xy, const u64* scalars, size_t n, int threads, u64* out_xy,
                         uint8_t* out_inf) {
  init_all();
  if (curve == BN254 && group == 1) fixed_base_mul<BnFq>(base_xy, scalars, n, BnFr::C.bits, threads, out_xy, out_inf);
  else if (curve == BN254 && group == 2) fixed_base_mul<Fp2<BnFq>>(base_xy, scalars, n, BnFr::C.bits, threads, out_xy, out_inf);
  else if (curve == BLS12_381 && group == 1) fixed_base_mul<BlsFq>(base_xy, scalars, n, BlsFr::C.bits, threads, out_xy, out_inf);
  else if (curve == BLS12_381 && group == 2) fixed_base_mul<Fp2<BlsFq>>(base_xy, scalars, n, BlsFr::C.bits, threads, out_xy, out_inf);
  else return -1;
  return 0;
}

int zkref_ntt(int curve, u64* data, unsigned log_n, unsigned flags, int threads) {
  init_all();
  return curve == BLS12_381 ? ntt_t<BlsFr>(data, log_n, flags, threads) : ntt_t<BnFr>(data, log_n, flags, threads);
}

// mode 0: into_repr, 1: from_repr
int zkref_fr_convert(int curve, const u64* in, u64* out, size_t n, int mode) {
  init_all();
  for (size_t i = 0; i < n; i++) {
    if (curve == BLS12_381) {
      BlsFr x;
      if (mode) { x = BlsFr::from_canonical(in + 4 * i); memcpy(out + 4 * i, x.v, 32); }
      else { memcpy(x.v, in + 4 * i, 32); x.to_canonical(out + 4 * i); }
    } else {
      BnFr x;
      if (mode) { x = BnFr::from_canonical(in + 4 * i); memcpy(out + 4 * i, x.v, 32); }
      else { memcpy(x.v, in + 4 * i, 32); x.to_canonical(out + 4 * i); }
    }
  }
  return 0;
}

// witness_map + into_repr: h_canonical gets next_pow2(n_rows + n_inputs) * 4 limbs
int zkref_witness_map(int curve, const Csr* A, const Csr* B, const Csr* C, const u64* z_mont, size_t n_inputs, int threads,
                      u64* h_canonical) {
  init_all();
  if (curve == BLS12_381) {
    std::vector<BlsFr> h;
    int rc = witness_map<BlsFr>(*A, *B, *C, reinterpret_cast<const BlsFr*>(z_mont), n_inputs, threads, h);
    if (rc) return rc;
    for (size_t i = 0; i < h.size(); i++) h[i].to_canonical(h_canonical + 4 * i);
  } else {
    std::vector<BnFr> h;
    int rc = witness_map<BnFr>(*A, *B, *C, reinterpret_cast<const BnFr*>(z_mont), n_inputs, threads, h);
    if (rc) return rc;
    for (size_t i = 0; i < h.size(); i++) h[i].to_canonical(h_canonical + 4 * i);
  }
  return 0;
}

int zkref_groth16_prove(int curve, const G16Key* pk, const Csr* A, const Csr* B, const Csr* C, const u64* z_mont,
                        size_t n_inputs, size_t n_aux, const u64* r, const u64* s, int threads, u64* proof_xy,
                        uint8_t* proof_inf) {
  init_all();
  if (curve == BLS12_381) return groth16_prove<BlsFr, BlsFq>(*pk, *A, *B, *C, z_mont, n_inputs, n_aux, r, s, threads, proof_xy, proof_inf);
  if (curve == BN254) return groth16_prove<BnFr, BnFq>(*pk, *A, *B, *C, z_mont, n_inputs, n_aux, r, s, threads, proof_xy, proof_inf);
  return -1;
}

}  // extern "C"
