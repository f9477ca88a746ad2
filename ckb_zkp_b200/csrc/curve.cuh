// Short-Weierstrass a = 0 group arithmetic in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2), generic over the coordinate field (Fp -> G1, Fp2 -> G2).
//
// Replaces ark-ec 0.2 `GroupProjective::{add_assign_mixed, add_assign, double_in_place,
// into_affine}` as used from groth16/src/prover.rs:164-210.  The reference works in
// Jacobian coordinates; since every result leaves the device as a canonical affine
// point, the coordinate system is free -- XYZZ saves a squaring per mixed addition
// and needs no inversion-free "Z=1" special case.
//
// Device representation of the identity: affine (0, 0) (never on y^2 = x^3 + b, b != 0);
// XYZZ with ZZ = 0.
#pragma once
#include "field.cuh"

namespace zkb {

template <class F>
struct Affine {
  F x, y;
  ZKB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  ZKB_HD static Affine inf() { return {F::zero(), F::zero()}; }
};

// Out-of-line (never inlined) versions of the group operations, defined after XYZZ.  Cold kernels
// and the rare branches of the hot ones call these so that every translation unit carries each
// unrolled formula once instead of once per call site (ptxas time and instruction-cache footprint).
template <class F> struct XYZZ;
template <class F> ZKB_NOINLINE void pt_add(XYZZ<F>& a, const XYZZ<F>& b);
template <class F> ZKB_NOINLINE void pt_dbl(XYZZ<F>& a);
template <class F> ZKB_NOINLINE XYZZ<F> pt_dbl_affine(F x, F y);
template <class F> ZKB_NOINLINE void pt_madd(XYZZ<F>& a, const Affine<F>& p, bool negate);
template <class F> ZKB_NOINLINE void pt_to_affine(Affine<F>& out, const XYZZ<F>& a);

template <class F>
struct XYZZ {
  F X, Y, ZZ, ZZZ;

  ZKB_HD bool is_inf() const { return ZZ.is_zero(); }
  ZKB_HD static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  ZKB_HD static XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return {p.x, p.y, F::one(), F::one()};
  }

  // dbl-2008-s-1
  ZKB_HD static XYZZ dbl(const XYZZ& p) {
    if (p.is_inf()) return p;
    F U = F::dbl(p.Y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(p.X, V);
    F M = F::sqr(p.X);
    M = F::add(F::dbl(M), M);
    XYZZ r;
    r.X = F::sub(F::sqr(M), F::dbl(S));
    r.Y = F::sub(F::mul(M, F::sub(S, r.X)), F::mul(W, p.Y));
    r.ZZ = F::mul(V, p.ZZ);
    r.ZZZ = F::mul(W, p.ZZZ);
    return r;
  }
  // doubling of an affine point (mdbl-2008-s-1)
  ZKB_HD static XYZZ dbl_affine(const F& x, const F& y) {
    F U = F::dbl(y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(x, V);
    F M = F::sqr(x);
    M = F::add(F::dbl(M), M);
    XYZZ r;
    r.X = F::sub(F::sqr(M), F::dbl(S));
    r.Y = F::sub(F::mul(M, F::sub(S, r.X)), F::mul(W, y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
  }

  // acc += (x, +-y)   (madd-2008-s); the affine point must not be the identity
  ZKB_HD void madd_xy(const F& x2, const F& y2in, bool negate) {
    F y2 = negate ? F::neg(y2in) : y2in;
    if (is_inf()) {
      X = x2; Y = y2; ZZ = F::one(); ZZZ = F::one();
      return;
    }
    F U2 = F::mul(x2, ZZ);
    F S2 = F::mul(y2, ZZZ);
    F Pp = F::sub(U2, X);
    F R = F::sub(S2, Y);
    if (Pp.is_zero()) {
      if (R.is_zero()) *this = pt_dbl_affine<F>(x2, y2);
      else *this = inf();
      return;
    }
    F PP = F::sqr(Pp);
    F PPP = F::mul(Pp, PP);
    F Q = F::mul(X, PP);
    F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    Y = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(Y, PPP));
    X = X3;
    ZZ = F::mul(ZZ, PP);
    ZZZ = F::mul(ZZZ, PPP);
  }
  ZKB_HD void madd(const Affine<F>& p, bool negate = false) {
    if (p.is_inf()) return;
    madd_xy(p.x, p.y, negate);
  }

  // acc += q   (add-2008-s)
  ZKB_HD void add(const XYZZ& q) {
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    F U1 = F::mul(X, q.ZZ);
    F U2 = F::mul(q.X, ZZ);
    F S1 = F::mul(Y, q.ZZZ);
    F S2 = F::mul(q.Y, ZZZ);
    F Pp = F::sub(U2, U1);
    F R = F::sub(S2, S1);
    if (Pp.is_zero()) {
      if (R.is_zero()) pt_dbl(*this);
      else *this = inf();
      return;
    }
    F PP = F::sqr(Pp);
    F PPP = F::mul(Pp, PP);
    F Q = F::mul(U1, PP);
    F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    Y = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(S1, PPP));
    X = X3;
    ZZ = F::mul(F::mul(ZZ, q.ZZ), PP);
    ZZZ = F::mul(F::mul(ZZZ, q.ZZZ), PPP);
  }
  ZKB_HD void neg_in_place() { Y = F::neg(Y); }

  // canonical affine (identity -> (0,0))
  ZKB_HD Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::inf();
    // 1/ZZZ gives both: 1/ZZ = ZZ^2 / ZZZ^2 ... cheaper: one inversion of ZZ*ZZZ
    F zi = F::inv_fast(F::mul(ZZ, ZZZ));   // 1/(ZZ*ZZZ): division-step inversion (field.cuh), ~5x shorter than the Euclid one
    F zz_inv = F::mul(zi, ZZZ);            // 1/ZZ
    F zzz_inv = F::mul(zi, ZZ);            // 1/ZZZ
    return {F::mul(X, zz_inv), F::mul(Y, zzz_inv)};
  }

  // k * p for a scalar given as little-endian 32-bit limbs (canonical integer)
  ZKB_HD static XYZZ mul_limbs(const XYZZ& p, const uint32_t* k, int nlimbs) {
    XYZZ r = inf();
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; i--) {
      for (int b = 31; b >= 0; b--) {
        if (started) pt_dbl(r);
        if ((k[i] >> b) & 1) { pt_add(r, p); started = true; }
      }
    }
    return r;
  }
  ZKB_HD static XYZZ mul_u32(const XYZZ& p, uint32_t k) { return mul_limbs(p, &k, 1); }
};

template <class F> ZKB_NOINLINE void pt_add(XYZZ<F>& a, const XYZZ<F>& b) { a.add(b); }
template <class F> ZKB_NOINLINE void pt_dbl(XYZZ<F>& a) { a = XYZZ<F>::dbl(a); }
template <class F> ZKB_NOINLINE XYZZ<F> pt_dbl_affine(F x, F y) { return XYZZ<F>::dbl_affine(x, y); }
template <class F> ZKB_NOINLINE void pt_madd(XYZZ<F>& a, const Affine<F>& p, bool negate) { a.madd(p, negate); }
template <class F> ZKB_NOINLINE void pt_to_affine(Affine<F>& out, const XYZZ<F>& a) { out = a.to_affine(); }

// the call-multiplication twin of a coordinate field (same memory layout)
template <class F> struct CallVariant;
template <class P> struct CallVariant<Fp<P>> { using type = FpC<P>; };
template <class P> struct CallVariant<Fp2<P, Fp<P>>> { using type = Fp2<P, FpC<P>>; };

}  // namespace zkb
