// Prime-field arithmetic in Montgomery form on 32-bit limbs (R = 2^(32*N)), and the
// quadratic extension Fq2 = Fq[u]/(u^2 + 1) used by both G2 groups.
//
// Replaces what the reference obtains from ark-ff 0.2 `Fp256/Fp384` (Montgomery,
// R = 2^(64*limbs): identical R, identical bytes in memory -- u64 LE limbs are two
// u32 LE limbs), i.e. the arithmetic underneath groth16/src/prover.rs:187-228 and
// groth16/src/r1cs_to_qap.rs:131-169.
//
// Multiplication is a word-serial CIOS: for every limb b[i] the products a[j]*b[i]
// are added in two carry chains (even j / odd j) of mad.lo.cc + madc.hi.cc pairs,
// which ptxas fuses into IMAD.WIDE.U32 with carry; then one Montgomery step
// (m = T[0] * -p^-1, T += m*p, shift one limb).  All values stay fully reduced
// (< p) so equality tests are plain limb compares.
#pragma once
#include "field_params.cuh"

namespace zkb {

template <class P>
struct alignas(16) Fp {
  static constexpr int N = P::N;
  using Params = P;
  uint32_t v[N];

  ZKB_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  ZKB_HD static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::one(i);
    return r;
  }
  ZKB_HD static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
    return r;
  }
  ZKB_HD bool is_zero() const {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= v[i];
    return t == 0;
  }
  ZKB_HD bool operator==(const Fp& o) const {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= v[i] ^ o.v[i];
    return t == 0;
  }
  ZKB_HD bool operator!=(const Fp& o) const { return !(*this == o); }

  // r = a + b mod p
  ZKB_HD static Fp add(const Fp& a, const Fp& b) {
    Fp s, t;
    s.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) s.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    s.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);   // 2p < 2^(32N): no carry out
    t.v[0] = ptx::sub_cc(s.v[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t.v[i] = ptx::subc_cc(s.v[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);                 // 0xffffffff when s < p
#pragma unroll
    for (int i = 0; i < N; i++) s.v[i] = borrow ? s.v[i] : t.v[i];
    return s;
  }
  // r = a - b mod p
  ZKB_HD static Fp sub(const Fp& a, const Fp& b) {
    Fp d;
    d.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) d.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = ptx::subc(0, 0);
    d.v[0] = ptx::add_cc(d.v[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) d.v[i] = ptx::addc_cc(d.v[i], P::mod(i) & borrow);
    d.v[N - 1] = ptx::addc(d.v[N - 1], P::mod(N - 1) & borrow);
    return d;
  }
  ZKB_HD static Fp neg(const Fp& a) { return sub(zero(), a); }
  ZKB_HD static Fp dbl(const Fp& a) { return add(a, a); }

  // ---- Montgomery multiplication -------------------------------------------------
  // The running value is kept as T = X + 2^32 * Y in two N-limb accumulators so that
  // every 64-bit partial product lands on an even-aligned register pair of one of them
  // (IMAD.WIDE needs aligned pairs; a single array would need a realigning MOV per limb
  // per row).  Products of even columns a[j] go to X at limbs (j, j+1); products of odd
  // columns go to Y at limbs (j-1, j).  One row = multiply-accumulate by b_i, then one
  // Montgomery step m = X[0] * (-p^-1), T += m * p, which zeroes X[0].  Dividing by 2^32
  // swaps the roles: T/2^32 = Y + X[1] + 2^32 * (X >> 64), i.e. Y becomes the even
  // accumulator of the next row (plus the straggler X[1], whose carry has exactly the
  // weight of limb 0 of the new odd accumulator) and X >> 64 the new odd accumulator.

  // Y chain for odd columns: (Y[j-1], Y[j]) += c[j] * s
  template <class Acc>
  ZKB_HD static void odd_chain(uint32_t* Y, Acc c, uint32_t s) {
    ptx::mad_wide_cc(Y[0], Y[1], c(1), s);
#pragma unroll
    for (int j = 3; j < N; j += 2) ptx::madc_wide_cc(Y[j - 1], Y[j], c(j), s);
  }
  // X chain for even columns: (X[j], X[j+1]) += c[j] * s; leaves the carry out in CC
  template <class Acc>
  ZKB_HD static void even_chain(uint32_t* X, Acc c, uint32_t s) {
    ptx::mad_wide_cc(X[0], X[1], c(0), s);
#pragma unroll
    for (int j = 2; j < N; j += 2) ptx::madc_wide_cc(X[j], X[j + 1], c(j), s);
  }
  struct ModAcc { ZKB_HD uint32_t operator()(int j) const { return P::mod(j); } };
  struct ArrAcc { const uint32_t* a; ZKB_HD uint32_t operator()(int j) const { return a[j]; } };

  // One row.  On entry (not first): X = clean even accumulator, Y = previous even
  // accumulator with Y[0] == 0 and straggler Y[1].  On exit: X[0] == 0, roles swapped.
  ZKB_HD static void mad_redc_row(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi, bool first) {
    if (first) {
#pragma unroll
      for (int j = 0; j < N; j += 2) ptx::mul_wide(X[j], X[j + 1], a[j], bi);
#pragma unroll
      for (int j = 1; j < N; j += 2) ptx::mul_wide(Y[j - 1], Y[j], a[j], bi);
    } else {
      X[0] = ptx::add_cc(X[0], Y[1]);
      // shifted odd chain: new Y = (old Y >> 64) + odd products + carry of the straggler add
#pragma unroll
      for (int j = 1; j < N - 1; j += 2) ptx::madc_wide_cc_3(Y[j - 1], Y[j], a[j], bi, Y[j + 1], Y[j + 2]);
      ptx::madc_wide_cc_3(Y[N - 2], Y[N - 1], a[N - 1], bi, 0, 0);
      even_chain(X, ArrAcc{a}, bi);
      Y[N - 1] = ptx::addc(Y[N - 1], 0);          // carry of the even chain has weight 2^(32N)
    }
    uint32_t m = ptx::mul_lo(X[0], P::INV);
    odd_chain(Y, ModAcc{}, m);
    even_chain(X, ModAcc{}, m);
    Y[N - 1] = ptx::addc(Y[N - 1], 0);
  }
  // result = Y + (X >> 32) after an even number of rows (last row had X = odd array)
  ZKB_HD static Fp merge_final(const uint32_t* Y, const uint32_t* X) {
    uint32_t T[N];
    T[0] = ptx::add_cc(Y[0], X[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) T[k] = ptx::addc_cc(Y[k], X[k + 1]);
    T[N - 1] = ptx::addc(Y[N - 1], 0);
    return final_sub(T);
  }
  // final conditional subtraction: T (N limbs, < 2p) -> [0, p)
  ZKB_HD static Fp final_sub(const uint32_t* T) {
    Fp s, t;
    t.v[0] = ptx::sub_cc(T[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t.v[i] = ptx::subc_cc(T[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
    for (int i = 0; i < N; i++) s.v[i] = borrow ? T[i] : t.v[i];
    return s;
  }

  // Montgomery product a * b * R^-1 mod p
  ZKB_HD static Fp mul(const Fp& a, const Fp& b) {
    static_assert(N % 2 == 0, "even limb count expected");
    uint32_t even[N], odd[N];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      mad_redc_row(even, odd, a.v, b.v[i], i == 0);
      mad_redc_row(odd, even, a.v, b.v[i + 1], false);
    }
    return merge_final(even, odd);
  }
  ZKB_HD static Fp sqr(const Fp& a) { return mul(a, a); }

  // ---- alternative: Karatsuba product + separated reduction ---------------------------------------
  // out[0 .. 2H) = a[0 .. H) * b[0 .. H).  Same even / odd accumulator idea as mul(): the product of
  // limbs (i, j) lands at position i + j, even positions accumulate in X, odd ones in Y (stored one limb
  // lower), so every IMAD.WIDE hits an aligned register pair; each row is two carry chains whose carry
  // out goes to the next limb, which at that point holds at most a carry bit of the previous row.
  template <int H>
  ZKB_HD static void wide_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
    static_assert(H % 2 == 0, "even half size expected");
    uint32_t X[2 * H], Y[2 * H];
#pragma unroll
    for (int k = 0; k < 2 * H; k++) X[k] = Y[k] = 0;
#pragma unroll
    for (int j = 0; j < H; j += 2) ptx::mul_wide(X[j], X[j + 1], a[j], b[0]);
#pragma unroll
    for (int j = 1; j < H; j += 2) ptx::mul_wide(Y[j - 1], Y[j], a[j], b[0]);
#pragma unroll
    for (int i = 1; i < H; i++) {
      if (i % 2 == 0) {
        ptx::mad_wide_cc(X[i], X[i + 1], a[0], b[i]);
#pragma unroll
        for (int j = 2; j < H; j += 2) ptx::madc_wide_cc(X[i + j], X[i + j + 1], a[j], b[i]);
        X[i + H] = ptx::addc(X[i + H], 0);
        ptx::mad_wide_cc(Y[i], Y[i + 1], a[1], b[i]);
#pragma unroll
        for (int j = 3; j < H; j += 2) ptx::madc_wide_cc(Y[i + j - 1], Y[i + j], a[j], b[i]);
        Y[i + H] = ptx::addc(Y[i + H], 0);
      } else {
        ptx::mad_wide_cc(Y[i - 1], Y[i], a[0], b[i]);
#pragma unroll
        for (int j = 2; j < H; j += 2) ptx::madc_wide_cc(Y[i + j - 1], Y[i + j], a[j], b[i]);
        Y[i + H - 1] = ptx::addc(Y[i + H - 1], 0);
        ptx::mad_wide_cc(X[i + 1], X[i + 2], a[1], b[i]);
#pragma unroll
        for (int j = 3; j < H; j += 2) ptx::madc_wide_cc(X[i + j], X[i + j + 1], a[j], b[i]);
        if (i + H + 1 < 2 * H) X[i + H + 1] = ptx::addc(X[i + H + 1], 0);     // top row: X < 2^(64 H), no carry out
      }
    }
    out[0] = X[0];
    out[1] = ptx::add_cc(X[1], Y[0]);
#pragma unroll
    for (int k = 2; k < 2 * H - 1; k++) out[k] = ptx::addc_cc(X[k], Y[k - 1]);
    out[2 * H - 1] = ptx::addc(X[2 * H - 1], Y[2 * H - 2]);
  }
  // T[0 .. 2N) = a * b with one Karatsuba level: 3 half-size products instead of 4
  ZKB_HD static void wide_mul_karatsuba(const uint32_t* a, const uint32_t* b, uint32_t* T) {
    constexpr int H = N / 2;
    uint32_t sa[H], sb[H], Pm[N + 1];
    sa[0] = ptx::add_cc(a[0], a[H]);
#pragma unroll
    for (int k = 1; k < H; k++) sa[k] = ptx::addc_cc(a[k], a[H + k]);
    const uint32_t ca = ptx::addc(0, 0);
    sb[0] = ptx::add_cc(b[0], b[H]);
#pragma unroll
    for (int k = 1; k < H; k++) sb[k] = ptx::addc_cc(b[k], b[H + k]);
    const uint32_t cb = ptx::addc(0, 0);
    wide_mul<H>(a, b, T);
    wide_mul<H>(a + H, b + H, T + N);
    wide_mul<H>(sa, sb, Pm);
    // (sa + ca 2^(32H)) (sb + cb 2^(32H)) = sa sb + 2^(32H) (ca sb + cb sa) + 2^(64H) ca cb
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    Pm[H] = ptx::add_cc(Pm[H], sb[0] & ma);
#pragma unroll
    for (int k = 1; k < H; k++) Pm[H + k] = ptx::addc_cc(Pm[H + k], sb[k] & ma);
    Pm[N] = ptx::addc(0, 0);
    Pm[H] = ptx::add_cc(Pm[H], sa[0] & mb);
#pragma unroll
    for (int k = 1; k < H; k++) Pm[H + k] = ptx::addc_cc(Pm[H + k], sa[k] & mb);
    Pm[N] = ptx::addc(Pm[N], ca & cb);
    // middle term = Pm - low product - high product
    Pm[0] = ptx::sub_cc(Pm[0], T[0]);
#pragma unroll
    for (int k = 1; k < N; k++) Pm[k] = ptx::subc_cc(Pm[k], T[k]);
    Pm[N] = ptx::subc(Pm[N], 0);
    Pm[0] = ptx::sub_cc(Pm[0], T[N]);
#pragma unroll
    for (int k = 1; k < N; k++) Pm[k] = ptx::subc_cc(Pm[k], T[N + k]);
    Pm[N] = ptx::subc(Pm[N], 0);
    T[H] = ptx::add_cc(T[H], Pm[0]);
#pragma unroll
    for (int k = 1; k <= N; k++) T[H + k] = ptx::addc_cc(T[H + k], Pm[k]);
#pragma unroll
    for (int k = H + N + 1; k < 2 * N - 1; k++) T[k] = ptx::addc_cc(T[k], 0);
    T[2 * N - 1] = ptx::addc(T[2 * N - 1], 0);
  }
  // Montgomery reduction of a 2N-limb value T < p R: N rows of the m * p step of mul() on the low half
  // (the high half does not influence any m), then + high half, then one conditional subtraction.
  ZKB_HD static void redc_row(uint32_t* X, uint32_t* Y, bool first) {
    if (!first) {
      X[0] = ptx::add_cc(X[0], Y[1]);                 // straggler of the previous row
#pragma unroll
      for (int k = 0; k < N - 2; k++) Y[k] = ptx::addc_cc(Y[k + 2], 0);     // Y >>= 64, carrying the straggler's carry
      Y[N - 2] = ptx::addc(0, 0);
      Y[N - 1] = 0;
    }
    uint32_t m = ptx::mul_lo(X[0], P::INV);
    odd_chain(Y, ModAcc{}, m);
    even_chain(X, ModAcc{}, m);
    Y[N - 1] = ptx::addc(Y[N - 1], 0);
  }
  ZKB_HD static Fp redc_wide(const uint32_t* T) {
    uint32_t even[N], odd[N];
#pragma unroll
    for (int k = 0; k < N; k++) { even[k] = T[k]; odd[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      redc_row(even, odd, i == 0);
      redc_row(odd, even, false);
    }
    // (T_low + sum m_k p 2^(32k)) / R = even + (odd >> 32)   (<= p), then the high half (< p^2 / R)
    uint32_t V[N];
    V[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) V[k] = ptx::addc_cc(even[k], odd[k + 1]);
    V[N - 1] = ptx::addc(even[N - 1], 0);
    V[0] = ptx::add_cc(V[0], T[N]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) V[k] = ptx::addc_cc(V[k], T[N + k]);
    V[N - 1] = ptx::addc(V[N - 1], T[2 * N - 1]);
    return final_sub(V);
  }
  ZKB_HD static Fp mul_sos(const Fp& a, const Fp& b) {
    uint32_t T[2 * N];
    wide_mul_karatsuba(a.v, b.v, T);
    return redc_wide(T);
  }

  ZKB_HD static Fp to_mont(const Fp& a) { return mul(a, r2()); }
  ZKB_HD static Fp from_mont(const Fp& a) {
    Fp o = zero();
    o.v[0] = 1;
    return mul(a, o);
  }

  // a^(p-2) (Fermat); kept as the independent cross-check of inv()
  ZKB_HD static Fp inv_fermat(const Fp& a) {
    Fp r = one();
    for (int i = N - 1; i >= 0; i--) {
      uint32_t e = P::pm2(i);
      for (int b = 31; b >= 0; b--) {
        r = sqr(r);
        if ((e >> b) & 1) r = mul(r, a);
      }
    }
    return r;
  }
  // a^-1 (Montgomery in, Montgomery out; 0 -> 0) by a branch-free binary extended Euclid.
  // Invariants: x1 * a == u and x2 * a == v (mod p), v odd.  One round: if u is odd subtract the
  // smaller of (u, v) from the larger into u (which makes it even; v takes the old u when u < v),
  // same on (x1, x2) mod p, then halve u and x1.  len(u) + len(v) drops every round, so at most
  // 2 * BITS rounds run until u == 0, v == gcd == 1, x2 == a^-1.  Only add / logic / shift
  // instructions: on the GPU this is ~10x shorter than the 1.5 * BITS dependent multiplications of
  // the Fermat ladder and leaves the multiplier pipe to other warps.
  ZKB_HD static Fp inv(const Fp& a) {
    uint32_t u[N], v[N], x1[N], x2[N];
#pragma unroll
    for (int i = 0; i < N; i++) { u[i] = a.v[i]; v[i] = P::mod(i); x1[i] = i == 0 ? 1u : 0u; x2[i] = 0; }
    for (int round = 0; round < 2 * P::BITS + 2; round++) {
      uint32_t nz = 0;
#pragma unroll
      for (int i = 0; i < N; i++) nz |= u[i];
      if (nz == 0) break;
      const uint32_t odd = 0u - (u[0] & 1u);
      // borrow of u - v
      ptx::sub_cc(u[0], v[0]);
#pragma unroll
      for (int i = 1; i < N; i++) ptx::subc_cc(u[i], v[i]);
      const uint32_t sw = odd & ptx::subc(0, 0);        // all ones when u is odd and u < v
      // (u, v) <- (A - (B & odd), B) with (A, B) = sw ? (v, u) : (u, v)
      uint32_t A[N], B[N];
#pragma unroll
      for (int i = 0; i < N; i++) { A[i] = (v[i] & sw) | (u[i] & ~sw); B[i] = (u[i] & sw) | (v[i] & ~sw); }
      u[0] = ptx::sub_cc(A[0], B[0] & odd);
#pragma unroll
      for (int i = 1; i < N - 1; i++) u[i] = ptx::subc_cc(A[i], B[i] & odd);
      u[N - 1] = ptx::subc(A[N - 1], B[N - 1] & odd);
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = B[i];
      // the same on (x1, x2), modulo p
#pragma unroll
      for (int i = 0; i < N; i++) { A[i] = (x2[i] & sw) | (x1[i] & ~sw); B[i] = (x1[i] & sw) | (x2[i] & ~sw); }
      x1[0] = ptx::sub_cc(A[0], B[0] & odd);
#pragma unroll
      for (int i = 1; i < N; i++) x1[i] = ptx::subc_cc(A[i], B[i] & odd);
      const uint32_t borrow = ptx::subc(0, 0);
      x1[0] = ptx::add_cc(x1[0], P::mod(0) & borrow);
#pragma unroll
      for (int i = 1; i < N - 1; i++) x1[i] = ptx::addc_cc(x1[i], P::mod(i) & borrow);
      x1[N - 1] = ptx::addc(x1[N - 1], P::mod(N - 1) & borrow);
#pragma unroll
      for (int i = 0; i < N; i++) x2[i] = B[i];
      // halve u (even now) and x1 (add p first when odd; x1 + p < 2^(32N))
#pragma unroll
      for (int i = 0; i < N - 1; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
      u[N - 1] >>= 1;
      const uint32_t xo = 0u - (x1[0] & 1u);
      x1[0] = ptx::add_cc(x1[0], P::mod(0) & xo);
#pragma unroll
      for (int i = 1; i < N - 1; i++) x1[i] = ptx::addc_cc(x1[i], P::mod(i) & xo);
      x1[N - 1] = ptx::addc(x1[N - 1], P::mod(N - 1) & xo);
#pragma unroll
      for (int i = 0; i < N - 1; i++) x1[i] = (x1[i] >> 1) | (x1[i + 1] << 31);
      x1[N - 1] >>= 1;
    }
    // x2 = (aR)^-1 = a^-1 R^-1; two Montgomery products with R^2 bring it to a^-1 R
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = x2[i];
    return mul(mul(r, r2()), r2());
  }
  // ---- a^-1 by Bernstein-Yang division steps in batches of 30 ("safegcd", the half-delta variant) --------------
  // Same contract as inv() (Montgomery in, Montgomery out, 0 -> 0), ~5x fewer instructions: the 30 division steps of a
  // batch look only at the low 30 bits of (f, g) and produce a 2x2 transition matrix t with |entries| <= 2^30, which is
  // then applied to the full-width (f, g) and, modulo p, to (d, e) -- 10 signed 32x32->64 multiply-adds per 30-bit limb
  // and batch instead of ~190 add/logic instructions per bit.  Values are signed, L = ceil((BITS + 2) / 30) limbs of
  // 30 bits (top limb carries the sign); invariants d * x == f, e * x == g (mod p) up to the common power of two that
  // the modular update divides out.  Terminates when g == 0 (f = +-1); the step bound of the half-delta rule for a
  // BITS-bit modulus is below 49 * BITS / 17 + 4 steps, kMaxBatches covers it.  The batched-affine bucket accumulation
  // (msm_batch.cuh) runs one of these per thread and round, so its length decides how many buckets a thread must own.
  static constexpr int L30 = (P::BITS + 2 + 29) / 30;
  static constexpr int kMaxBatches = (49 * P::BITS / 17 + 4 + 29) / 30;
  ZKB_HD static constexpr int32_t mod30(int i) {            // limb i of p in base 2^30
    const int bit = 30 * i, w = bit >> 5, s = bit & 31;
    uint64_t x = w < N ? (uint64_t)P::mod(w) : 0;
    if (w + 1 < N) x |= (uint64_t)P::mod(w + 1) << 32;
    return (int32_t)((x >> s) & 0x3fffffffu);
  }
  ZKB_HD static Fp inv_safegcd(const Fp& a) {
    constexpr int32_t M30 = 0x3fffffff;
    constexpr uint32_t pinv30 = (0u - P::INV) & 0x3fffffffu;          // p^-1 mod 2^30
    int32_t d[L30], e[L30], f[L30], g[L30];
#pragma unroll
    for (int i = 0; i < L30; i++) {
      const int bit = 30 * i, w = bit >> 5, s = bit & 31;
      uint64_t x = w < N ? (uint64_t)a.v[w] : 0;
      if (w + 1 < N) x |= (uint64_t)a.v[w + 1] << 32;
      g[i] = (int32_t)((x >> s) & 0x3fffffffu);
      f[i] = mod30(i);
      d[i] = 0;
      e[i] = i == 0 ? 1 : 0;
    }
    int32_t zeta = -1;                                                // -(delta + 1/2), delta = 1/2
    for (int batch = 0; batch < kMaxBatches; batch++) {
      int32_t gz = 0;
#pragma unroll
      for (int i = 0; i < L30; i++) gz |= g[i];
      if (gz == 0) break;
      // 30 division steps on the low limbs -> t = (u v; q r)
      uint32_t u = 1, v = 0, q = 0, r = 1, fl = (uint32_t)f[0], gl = (uint32_t)g[0];
#pragma unroll 6
      for (int i = 0; i < 30; i++) {
        uint32_t c1 = (uint32_t)(zeta >> 31);                         // delta > 0
        const uint32_t c2 = 0u - (gl & 1u);                           // g odd
        const uint32_t x = (fl ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
        gl += x & c2; q += y & c2; r += z & c2;
        c1 &= c2;
        zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;
        fl += gl & c1; u += q & c1; v += r & c1;
        gl >>= 1; u <<= 1; v <<= 1;
      }
      const int32_t tu = (int32_t)u, tv = (int32_t)v, tq = (int32_t)q, tr = (int32_t)r;
      // (d, e) <- t * (d, e) / 2^30 mod p: a multiple of p is added first so the division is exact
      {
        const int32_t sd = d[L30 - 1] >> 31, se = e[L30 - 1] >> 31;
        int32_t md = (tu & sd) + (tv & se), me = (tq & sd) + (tr & se);
        int64_t cd = (int64_t)tu * d[0] + (int64_t)tv * e[0];
        int64_t ce = (int64_t)tq * d[0] + (int64_t)tr * e[0];
        md -= (int32_t)((pinv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
        me -= (int32_t)((pinv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
        cd += (int64_t)mod30(0) * md;
        ce += (int64_t)mod30(0) * me;
        cd >>= 30;
        ce >>= 30;
#pragma unroll
        for (int i = 1; i < L30; i++) {
          const int32_t di = d[i], ei = e[i];
          cd += (int64_t)tu * di + (int64_t)tv * ei + (int64_t)mod30(i) * md;
          ce += (int64_t)tq * di + (int64_t)tr * ei + (int64_t)mod30(i) * me;
          d[i - 1] = (int32_t)cd & M30;
          e[i - 1] = (int32_t)ce & M30;
          cd >>= 30;
          ce >>= 30;
        }
        d[L30 - 1] = (int32_t)cd;
        e[L30 - 1] = (int32_t)ce;
      }
      // (f, g) <- t * (f, g) / 2^30 (exact)
      {
        int64_t cf = (int64_t)tu * f[0] + (int64_t)tv * g[0];
        int64_t cg = (int64_t)tq * f[0] + (int64_t)tr * g[0];
        cf >>= 30;
        cg >>= 30;
#pragma unroll
        for (int i = 1; i < L30; i++) {
          const int32_t fi = f[i], gi = g[i];
          cf += (int64_t)tu * fi + (int64_t)tv * gi;
          cg += (int64_t)tq * fi + (int64_t)tr * gi;
          f[i - 1] = (int32_t)cf & M30;
          g[i - 1] = (int32_t)cg & M30;
          cf >>= 30;
          cg >>= 30;
        }
        f[L30 - 1] = (int32_t)cf;
        g[L30 - 1] = (int32_t)cg;
      }
    }
    // f = +-1 (or +-p when a == 0, then d == 0 mod p): x^-1 = d * sign(f), brought into [0, p)
    {
      const int32_t sign = f[L30 - 1] >> 31;                          // all ones when f < 0
      int32_t cond_add = d[L30 - 1] >> 31;
#pragma unroll
      for (int i = 0; i < L30; i++) d[i] = ((d[i] + (mod30(i) & cond_add)) ^ sign) - sign;
#pragma unroll
      for (int i = 0; i < L30 - 1; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
      cond_add = d[L30 - 1] >> 31;
#pragma unroll
      for (int i = 0; i < L30; i++) d[i] += mod30(i) & cond_add;
#pragma unroll
      for (int i = 0; i < L30 - 1; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
    }
    Fp rr;
#pragma unroll
    for (int j = 0; j < N; j++) {
      const int bit = 32 * j, k = bit / 30, o = bit - 30 * k;
      uint64_t x = (uint64_t)(uint32_t)d[k] >> o;
      if (k + 1 < L30) x |= (uint64_t)(uint32_t)d[k + 1] << (30 - o);
      if (k + 2 < L30) x |= (uint64_t)(uint32_t)d[k + 2] << (60 - o);
      rr.v[j] = (uint32_t)x;
    }
    // rr = (a R)^-1 = a^-1 R^-1 -> a^-1 R
    return mul(mul(rr, r2()), r2());
  }

  ZKB_HD static Fp inv_fast(const Fp& a) { return inv_safegcd(a); }

  // a^e for a runtime 64-bit exponent
  ZKB_HD static Fp pow_u64(const Fp& a, uint64_t e) {
    Fp r = one();
    for (int b = 63; b >= 0; b--) {
      r = sqr(r);
      if ((e >> b) & 1) r = mul(r, a);
    }
    return r;
  }
  ZKB_HD static Fp small(uint32_t k) {  // k as a field element (Montgomery)
    Fp t = zero();
    t.v[0] = k;
    return to_mont(t);
  }
};

// ---------------------------------------------------------------------------------
// Fq2 = Fq[u]/(u^2+1)   (ark-ff `Fp2` with NONRESIDUE = -1 for both curves)
// ---------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------
// FpC: the same field element (same bytes) whose multiplication is an out-of-line call.
// The bucket-accumulation loop is instruction-fetch bound when ten fully unrolled 381-bit
// multiplications are inlined (46 KB of SASS per iteration); with the multiplication as one
// shared 4.6 KB function the loop body fits the instruction cache.  Operands travel in
// registers (by-value ABI), the extra MOVs issue in the slots the 4-cycle IMAD.WIDE leaves free.
// ---------------------------------------------------------------------------------
#ifdef __CUDACC__
template <class P>
__device__ __noinline__ Fp<P> fp_mul_call(Fp<P> a, Fp<P> b) { return Fp<P>::mul(a, b); }
#else
template <class P>
inline Fp<P> fp_mul_call(Fp<P> a, Fp<P> b) { return Fp<P>::mul(a, b); }
#endif

#ifdef __CUDACC__
template <class P>
__device__ __noinline__ Fp<P> fp_inv_safegcd_call(Fp<P> a) { return Fp<P>::inv_safegcd(a); }
#else
template <class P>
inline Fp<P> fp_inv_safegcd_call(Fp<P> a) { return Fp<P>::inv_safegcd(a); }
#endif

template <class P>
struct alignas(16) FpC {
  static constexpr int N = P::N;
  using Params = P;
  Fp<P> f;
  ZKB_HD static FpC zero() { return {Fp<P>::zero()}; }
  ZKB_HD static FpC one() { return {Fp<P>::one()}; }
  ZKB_HD bool is_zero() const { return f.is_zero(); }
  ZKB_HD bool operator==(const FpC& o) const { return f == o.f; }
  ZKB_HD bool operator!=(const FpC& o) const { return f != o.f; }
  ZKB_HD static FpC add(const FpC& a, const FpC& b) { return {Fp<P>::add(a.f, b.f)}; }
  ZKB_HD static FpC sub(const FpC& a, const FpC& b) { return {Fp<P>::sub(a.f, b.f)}; }
  ZKB_HD static FpC neg(const FpC& a) { return {Fp<P>::neg(a.f)}; }
  ZKB_HD static FpC dbl(const FpC& a) { return {Fp<P>::dbl(a.f)}; }
  ZKB_HD static FpC mul(const FpC& a, const FpC& b) { return {fp_mul_call<P>(a.f, b.f)}; }
  ZKB_HD static FpC sqr(const FpC& a) { return {fp_mul_call<P>(a.f, a.f)}; }
  ZKB_HD static FpC inv(const FpC& a) { return {Fp<P>::inv(a.f)}; }
  ZKB_HD static FpC inv_fast(const FpC& a) { return {fp_inv_safegcd_call<P>(a.f)}; }
};

template <class P, class BaseT = Fp<P>>
struct Fp2 {
  using Base = BaseT;
  using Params = P;
  Base c0, c1;

  ZKB_HD static Fp2 zero() { return {Base::zero(), Base::zero()}; }
  ZKB_HD static Fp2 one() { return {Base::one(), Base::zero()}; }
  ZKB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  ZKB_HD bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  ZKB_HD bool operator!=(const Fp2& o) const { return !(*this == o); }
  ZKB_HD static Fp2 add(const Fp2& a, const Fp2& b) { return {Base::add(a.c0, b.c0), Base::add(a.c1, b.c1)}; }
  ZKB_HD static Fp2 sub(const Fp2& a, const Fp2& b) { return {Base::sub(a.c0, b.c0), Base::sub(a.c1, b.c1)}; }
  ZKB_HD static Fp2 neg(const Fp2& a) { return {Base::neg(a.c0), Base::neg(a.c1)}; }
  ZKB_HD static Fp2 dbl(const Fp2& a) { return {Base::dbl(a.c0), Base::dbl(a.c1)}; }
  ZKB_HD static Fp2 mul(const Fp2& a, const Fp2& b) {
    // Karatsuba: 3 base multiplications
    Base t0 = Base::mul(a.c0, b.c0);
    Base t1 = Base::mul(a.c1, b.c1);
    Base s = Base::mul(Base::add(a.c0, a.c1), Base::add(b.c0, b.c1));
    return {Base::sub(t0, t1), Base::sub(Base::sub(s, t0), t1)};
  }
  ZKB_HD static Fp2 sqr(const Fp2& a) {
    // (c0+c1)(c0-c1) + 2 c0 c1 u
    Base t = Base::mul(a.c0, a.c1);
    Base r0 = Base::mul(Base::add(a.c0, a.c1), Base::sub(a.c0, a.c1));
    return {r0, Base::dbl(t)};
  }
  ZKB_HD static Fp2 inv(const Fp2& a) {
    Base n = Base::inv(Base::add(Base::sqr(a.c0), Base::sqr(a.c1)));
    return {Base::mul(a.c0, n), Base::neg(Base::mul(a.c1, n))};
  }
  ZKB_HD static Fp2 inv_fast(const Fp2& a) {
    Base n = Base::inv_fast(Base::add(Base::sqr(a.c0), Base::sqr(a.c1)));
    return {Base::mul(a.c0, n), Base::neg(Base::mul(a.c1, n))};
  }
};

}  // namespace zkb
