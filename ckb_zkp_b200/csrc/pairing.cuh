// Ate pairing on BLS12-381 / BN254 for batched verification: the step after the proving path.
//
// Replaces what groth16/src/verifier.rs:18-44 (`verify_proof`: three Miller loops, one final exponentiation) and
// marlin/src/pc/kzg10.rs `check` / `batch_check` obtain from ark-ec 0.2's `PairingEngine` (un-vendored).  Those callers
// only compare GT elements, so any non-degenerate bilinear pairing on (G1, G2) gives the same accept / reject decision;
// this file computes
//
//     BLS12-381:  f_{|x|, Q}(P) ^ (3 (q^12 - 1) / r)        (plain ate, loop |t - 1| = |x|; the x-chain of the hard part
//                                                            yields the cube, gcd(3, r) = 1)
//     BN254:      (f_{6x+2, Q}(P) l_{[6x+2]Q, pi(Q)}(P) l_{[6x+2]Q + pi(Q), -pi^2(Q)}(P)) ^ ((q^12 - 1) / r)    (optimal ate)
//
// Work distribution: a verifier checks MANY proofs, so the unit of parallelism is the pair -- one thread runs one Miller
// loop (homogeneous doubling / addition steps on the twist, sparse products by the lines; the affine form with one Fq2
// inversion per step is kept as miller_loop_affine, whose value is the oracle's bit for bit), a second kernel multiplies
// each group's loop values and runs one final exponentiation per group.
// Every Fq multiplication is the out-of-line call of FpC (field.cuh), and the Fq6 / Fq12 products are out-of-line
// functions working on thread-local operands, so the whole pairing is a few tens of KB of SASS.
//
// Tower: Fq2 = Fq[u]/(u^2 + 1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = c + u (c = 1 / 9).
// In powers of w an element (c0, c1) = ((a0, a1, a2), (b0, b1, b2)) is a0 + b0 w + a1 w^2 + b1 w^3 + a2 w^4 + b2 w^5.
#pragma once
#include "curve.cuh"
#include "pairing_params.cuh"

namespace zkb {

template <class PP>
struct PairingT {
  using FqP = typename PP::FqP;
  using FC = FpC<FqP>;
  using F2 = Fp2<FqP, FC>;

  struct F6 { F2 a0, a1, a2; };
  struct F12 { F6 c0, c1; };

  // ---- Fq2 helpers ------------------------------------------------------------------------------------------------
  ZKB_HD static F2 mul_xi(const F2& z) {          // (c + u)(z0 + z1 u) = (c z0 - z1) + (z0 + c z1) u
    if (PP::XI_C == 1) return {FC::sub(z.c0, z.c1), FC::add(z.c0, z.c1)};
    FC t0 = FC::dbl(FC::dbl(FC::dbl(z.c0))), t1 = FC::dbl(FC::dbl(FC::dbl(z.c1)));   // 8 z
    return {FC::sub(FC::add(t0, z.c0), z.c1), FC::add(FC::add(t1, z.c1), z.c0)};     // c = 9
  }
  ZKB_HD static F2 conj(const F2& z) { return {z.c0, FC::neg(z.c1)}; }
  ZKB_HD static F2 mul_fq(const F2& z, const FC& k) { return {FC::mul(z.c0, k), FC::mul(z.c1, k)}; }

  // ---- Fq6 --------------------------------------------------------------------------------------------------------
  ZKB_HD static F6 f6_zero() { return {F2::zero(), F2::zero(), F2::zero()}; }
  ZKB_HD static F6 f6_one() { return {F2::one(), F2::zero(), F2::zero()}; }
  ZKB_HD static F6 f6_add(const F6& a, const F6& b) { return {F2::add(a.a0, b.a0), F2::add(a.a1, b.a1), F2::add(a.a2, b.a2)}; }
  ZKB_HD static F6 f6_sub(const F6& a, const F6& b) { return {F2::sub(a.a0, b.a0), F2::sub(a.a1, b.a1), F2::sub(a.a2, b.a2)}; }
  ZKB_HD static F6 f6_neg(const F6& a) { return {F2::neg(a.a0), F2::neg(a.a1), F2::neg(a.a2)}; }
  ZKB_HD static F6 f6_mul_v(const F6& a) { return {mul_xi(a.a2), a.a0, a.a1}; }
  static ZKB_NOINLINE void f6_mul(F6& r, const F6& a, const F6& b) {      // Karatsuba, 6 Fq2 products
    F2 v0 = F2::mul(a.a0, b.a0), v1 = F2::mul(a.a1, b.a1), v2 = F2::mul(a.a2, b.a2);
    F2 t0 = F2::sub(F2::sub(F2::mul(F2::add(a.a1, a.a2), F2::add(b.a1, b.a2)), v1), v2);
    F2 t1 = F2::sub(F2::sub(F2::mul(F2::add(a.a0, a.a1), F2::add(b.a0, b.a1)), v0), v1);
    F2 t2 = F2::sub(F2::sub(F2::mul(F2::add(a.a0, a.a2), F2::add(b.a0, b.a2)), v0), v2);
    r.a0 = F2::add(v0, mul_xi(t0));
    r.a1 = F2::add(t1, mul_xi(v2));
    r.a2 = F2::add(t2, v1);
  }
  static ZKB_NOINLINE void f6_inv(F6& r, const F6& a) {
    F2 t0 = F2::sub(F2::sqr(a.a0), mul_xi(F2::mul(a.a1, a.a2)));
    F2 t1 = F2::sub(mul_xi(F2::sqr(a.a2)), F2::mul(a.a0, a.a1));
    F2 t2 = F2::sub(F2::sqr(a.a1), F2::mul(a.a0, a.a2));
    F2 d = F2::add(F2::mul(a.a0, t0), mul_xi(F2::add(F2::mul(a.a2, t1), F2::mul(a.a1, t2))));
    F2 di = F2::inv_fast(d);
    r.a0 = F2::mul(t0, di); r.a1 = F2::mul(t1, di); r.a2 = F2::mul(t2, di);
  }

  // ---- Fq12 -------------------------------------------------------------------------------------------------------
  ZKB_HD static F12 f12_one() { return {f6_one(), f6_zero()}; }
  ZKB_HD static F12 f12_conj(const F12& a) { return {a.c0, f6_neg(a.c1)}; }
  static ZKB_NOINLINE void f12_mul(F12& r, const F12& a, const F12& b) {  // 3 Fq6 products
    F6 v0, v1, s;
    f6_mul(v0, a.c0, b.c0);
    f6_mul(v1, a.c1, b.c1);
    f6_mul(s, f6_add(a.c0, a.c1), f6_add(b.c0, b.c1));
    r.c1 = f6_sub(f6_sub(s, v0), v1);
    r.c0 = f6_add(v0, f6_mul_v(v1));
  }
  static ZKB_NOINLINE void f12_sqr(F12& r, const F12& a) {                // complex method, 2 Fq6 products
    F6 ab, s;
    f6_mul(ab, a.c0, a.c1);
    f6_mul(s, f6_add(a.c0, a.c1), f6_add(a.c0, f6_mul_v(a.c1)));
    r.c0 = f6_sub(f6_sub(s, ab), f6_mul_v(ab));
    r.c1 = f6_add(ab, ab);
  }
  static ZKB_NOINLINE void f12_inv(F12& r, const F12& a) {
    F6 t0, t1, d;
    f6_mul(t0, a.c0, a.c0);
    f6_mul(t1, a.c1, a.c1);
    f6_inv(d, f6_sub(t0, f6_mul_v(t1)));
    f6_mul(r.c0, a.c0, d);
    f6_mul(t0, a.c1, d);
    r.c1 = f6_neg(t0);
  }
  ZKB_HD static F2 frob_const(int k) {
    F2 g;
#pragma unroll
    for (int i = 0; i < FC::N; i++) { g.c0.f.v[i] = PP::frob(k, 0, i); g.c1.f.v[i] = PP::frob(k, 1, i); }
    return g;
  }
  static ZKB_NOINLINE void f12_frob(F12& r, const F12& a) {               // a^q
    r.c0.a0 = conj(a.c0.a0);
    r.c1.a0 = F2::mul(conj(a.c1.a0), frob_const(1));
    r.c0.a1 = F2::mul(conj(a.c0.a1), frob_const(2));
    r.c1.a1 = F2::mul(conj(a.c1.a1), frob_const(3));
    r.c0.a2 = F2::mul(conj(a.c0.a2), frob_const(4));
    r.c1.a2 = F2::mul(conj(a.c1.a2), frob_const(5));
  }
  // Squaring in the cyclotomic subgroup (Granger-Scott, "Faster squaring in the cyclotomic subgroup of sixth degree
  // extensions"): three Fq4 squarings of two Fq2 products each -- 18 Fq multiplications instead of the 36 of f12_sqr.
  // Valid after the easy part of the final exponentiation only.
  ZKB_HD static void fq4_sqr(F2& t0, F2& t1, const F2& a, const F2& b) {   // (a + b y)^2, y^2 = xi
    F2 ab = F2::mul(a, b);
    t0 = F2::sub(F2::sub(F2::mul(F2::add(a, b), F2::add(mul_xi(b), a)), ab), mul_xi(ab));
    t1 = F2::dbl(ab);
  }
  ZKB_HD static F2 three_t_minus_2z(const F2& t, const F2& z) { return F2::add(F2::dbl(F2::sub(t, z)), t); }
  ZKB_HD static F2 three_t_plus_2z(const F2& t, const F2& z) { return F2::add(F2::dbl(F2::add(t, z)), t); }
  static ZKB_NOINLINE void f12_cyclotomic_sqr(F12& r, const F12& a) {
    F2 t0, t1, t2, t3, t4, t5;
    fq4_sqr(t0, t1, a.c0.a0, a.c1.a1);
    fq4_sqr(t2, t3, a.c1.a0, a.c0.a2);
    fq4_sqr(t4, t5, a.c0.a1, a.c1.a2);
    F12 o;
    o.c0.a0 = three_t_minus_2z(t0, a.c0.a0);
    o.c1.a1 = three_t_plus_2z(t1, a.c1.a1);
    o.c1.a0 = three_t_plus_2z(mul_xi(t5), a.c1.a0);
    o.c0.a2 = three_t_minus_2z(t4, a.c0.a2);
    o.c0.a1 = three_t_minus_2z(t2, a.c0.a1);
    o.c1.a2 = three_t_plus_2z(t3, a.c1.a2);
    r = o;
  }

  // a^|x| by square-and-multiply (64-bit parameter, top bit first); a in the cyclotomic subgroup
  static ZKB_NOINLINE void f12_pow_x(F12& r, const F12& a) {
    F12 acc = a, t;
    int top = 63;
    while (!((PP::X_ABS >> top) & 1)) top--;
    for (int b = top - 1; b >= 0; b--) {
      f12_cyclotomic_sqr(t, acc);
      if ((PP::X_ABS >> b) & 1) f12_mul(acc, t, a); else acc = t;
    }
    r = acc;
  }

  // ---- Miller loop ------------------------------------------------------------------------------------------------
  // The line through psi(T) with twist slope lam evaluated at P = (xP, yP), up to a factor in Fq2:
  //   D-type:  yP - lam xP w + (lam x_T - y_T) w^3             M-type:  xi yP + (lam x_T - y_T) w^3 - lam xP w^5
  ZKB_HD static void line(F12& l, const F2& lam, const F2& tx, const F2& ty, const FC& xP, const FC& yP) {
    F2 t3 = F2::sub(F2::mul(lam, tx), ty);
    F2 lx = F2::neg(mul_fq(lam, xP));
    l.c0 = f6_zero(); l.c1 = f6_zero();
    if (PP::TWIST_D) {
      l.c0.a0 = {yP, FC::zero()};
      l.c1.a0 = lx;
      l.c1.a1 = t3;
    } else {
      l.c0.a0 = mul_xi(F2{yP, FC::zero()});
      l.c1.a1 = t3;
      l.c1.a2 = lx;
    }
  }

  // pi_q on the twist (psi^-1 o Frobenius o psi): (x', y') -> (conj(x') gamma_2, conj(y') gamma_3)
  ZKB_HD static void frob_twist(F2& xo, F2& yo, const F2& x, const F2& y) {
    xo = F2::mul(conj(x), frob_const(2));
    yo = F2::mul(conj(y), frob_const(3));
  }
  // f <- f * l_{T,R}(P), T <- T + R  (R affine on the twist)
  ZKB_HD static void add_step_affine(F12& f, F2& tx, F2& ty, const F2& xR, const F2& yR, const FC& xP, const FC& yP) {
    F12 l, t;
    F2 lam = F2::mul(F2::sub(yR, ty), F2::inv_fast(F2::sub(xR, tx)));
    line(l, lam, tx, ty, xP, yP);
    f12_mul(t, f, l);
    f = t;
    F2 x3 = F2::sub(F2::sub(F2::sqr(lam), tx), xR);
    ty = F2::sub(F2::mul(lam, F2::sub(tx, x3)), ty);
    tx = x3;
  }

  // The Miller function: f_{|x|, Q}(P) on BLS12-381 (plain ate); on BN254 the optimal ate
  // f_{6x+2, Q}(P) l_{[6x+2]Q, pi(Q)}(P) l_{[6x+2]Q + pi(Q), -pi^2(Q)}(P).
  // P in G1 (affine over Fq), Q in G2 (affine on the twist over Fq2); identity on either side -> 1.
  // Q must lie in the order-r subgroup (no vertical line can then occur before the loop ends).
  static ZKB_NOINLINE void miller_loop_affine(F12& f, const FC& xP, const FC& yP, bool p_inf, const F2& xQ, const F2& yQ, bool q_inf) {
    f = f12_one();
    if (p_inf || q_inf) return;
    F2 tx = xQ, ty = yQ;
    F12 l, t;
    for (int i = PP::LOOP_BITS - 2; i >= 0; i--) {
      F2 x2 = F2::sqr(tx);
      F2 lam = F2::mul(F2::add(F2::dbl(x2), x2), F2::inv_fast(F2::dbl(ty)));
      line(l, lam, tx, ty, xP, yP);
      f12_sqr(t, f);
      f12_mul(f, t, l);
      F2 x3 = F2::sub(F2::sqr(lam), F2::dbl(tx));
      ty = F2::sub(F2::mul(lam, F2::sub(tx, x3)), ty);
      tx = x3;
      if ((PP::loop(i >> 5) >> (i & 31)) & 1) add_step_affine(f, tx, ty, xQ, yQ, xP, yP);
    }
    if (PP::OPT_ATE_BN) {                    // lines through pi(Q) and -pi^2(Q)
      F2 x1, y1, x2, y2;
      frob_twist(x1, y1, xQ, yQ);
      frob_twist(x2, y2, x1, y1);
      add_step_affine(f, tx, ty, x1, y1, xP, yP);
      add_step_affine(f, tx, ty, x2, F2::neg(y2), xP, yP);
    }
  }

  // ---- the same loop without inversions ----------------------------------------------------------------------------
  // T is kept in homogeneous coordinates (X : Y : Z) on the twist and every line is scaled by an element of Fq2 (2 Y Z^2 in
  // a doubling step, x_Q Z - X in an addition step), which the final exponentiation removes: the value differs from
  // miller_loop_affine's by a factor in Fq2*, the pairing does not.  A step costs ~38 Fq products instead of ~80 (the
  // division-step inversion alone is ~60), and the product by the line uses its sparsity (13 Fq2 products instead of 18).
  //   doubling (dbl-2007-bl, a = 0):  w = 3 X^2, s = 2 Y Z, R = Y s, B = 2 X R, h = w^2 - 2 B,
  //                                   (X, Y, Z) <- (h s, w (B - h) - 2 R^2, s^3);   line: (s Z) yP, -(w Z) xP, w X - R
  //   addition (madd-1998-cmo):       u = y_Q Z - Y, v = x_Q Z - X, R = v^2 X, A = u^2 Z - v^3 - 2 R,
  //                                   (X, Y, Z) <- (v A, u (R - A) - v^3 Y, v^3 Z); line: v yP, -u xP, u x_Q - v y_Q
  ZKB_HD static F6 f6_scale(const F6& a, const F2& k) { return {F2::mul(a.a0, k), F2::mul(a.a1, k), F2::mul(a.a2, k)}; }
  // a * (x0 + x1 v)  /  a * (x1 v + x2 v^2): five Fq2 products each (Karatsuba on the two non-zero coefficients)
  ZKB_HD static F6 f6_mul_01(const F6& a, const F2& x0, const F2& x1) {
    F2 p00 = F2::mul(a.a0, x0), p11 = F2::mul(a.a1, x1);
    F2 mid = F2::sub(F2::sub(F2::mul(F2::add(a.a0, a.a1), F2::add(x0, x1)), p00), p11);       // a0 x1 + a1 x0
    return {F2::add(p00, mul_xi(F2::mul(a.a2, x1))), mid, F2::add(p11, F2::mul(a.a2, x0))};
  }
  ZKB_HD static F6 f6_mul_12(const F6& a, const F2& x1, const F2& x2) {
    F2 p11 = F2::mul(a.a1, x1), p22 = F2::mul(a.a2, x2);
    F2 mid = F2::sub(F2::sub(F2::mul(F2::add(a.a1, a.a2), F2::add(x1, x2)), p11), p22);       // a1 x2 + a2 x1
    return {mul_xi(mid), F2::add(F2::mul(a.a0, x1), mul_xi(p22)), F2::add(F2::mul(a.a0, x2), p11)};
  }
  // r = a * line, line = l0 + lw w + l3 w^3 (D-type twist)  or  l0 + l3 w^3 + lw w^5 (M-type), l0 already carrying xi there
  static ZKB_NOINLINE void f12_mul_by_line(F12& r, const F12& a, const F2& l0, const F2& lw, const F2& l3) {
    F6 v0 = f6_scale(a.c0, l0), v1, s;
    F6 sum = f6_add(a.c0, a.c1);
    if (PP::TWIST_D) {                       // c1 of the line = (lw, l3, 0)
      v1 = f6_mul_01(a.c1, lw, l3);
      s = f6_mul_01(sum, F2::add(l0, lw), l3);
    } else {                                 // c1 of the line = (0, l3, lw)
      v1 = f6_mul_12(a.c1, l3, lw);
      f6_mul(s, sum, F6{l0, l3, lw});
    }
    r.c1 = f6_sub(f6_sub(s, v0), v1);
    r.c0 = f6_add(v0, f6_mul_v(v1));
  }
  ZKB_HD static void line_coeffs(F2& l0, F2& lw, const F2& ky, const F2& kx, const FC& xP, const FC& yP) {
    l0 = mul_fq(ky, yP);                     // (scale) yP, times xi on an M-type twist
    if (!PP::TWIST_D) l0 = mul_xi(l0);
    lw = F2::neg(mul_fq(kx, xP));            // -(scale * slope) xP
  }
  static ZKB_NOINLINE void add_step(F12& f, F2& X, F2& Y, F2& Z, const F2& xR, const F2& yR, const FC& xP, const FC& yP) {
    F12 t;
    F2 l0, lw, l3;
    F2 u = F2::sub(F2::mul(yR, Z), Y), v = F2::sub(F2::mul(xR, Z), X);
    line_coeffs(l0, lw, v, u, xP, yP);
    l3 = F2::sub(F2::mul(u, xR), F2::mul(v, yR));
    F2 vv = F2::sqr(v);
    F2 vvv = F2::mul(v, vv);
    F2 R = F2::mul(vv, X);
    F2 A = F2::sub(F2::sub(F2::mul(F2::sqr(u), Z), vvv), F2::dbl(R));
    X = F2::mul(v, A);
    Y = F2::sub(F2::mul(u, F2::sub(R, A)), F2::mul(vvv, Y));
    Z = F2::mul(vvv, Z);
    f12_mul_by_line(t, f, l0, lw, l3);
    f = t;
  }
  static ZKB_NOINLINE void miller_loop(F12& f, const FC& xP, const FC& yP, bool p_inf, const F2& xQ, const F2& yQ, bool q_inf) {
    f = f12_one();
    if (p_inf || q_inf) return;
    F2 X = xQ, Y = yQ, Z = F2::one();
    F12 t;
    F2 l0, lw, l3;
    for (int i = PP::LOOP_BITS - 2; i >= 0; i--) {
      {
        F2 XX = F2::sqr(X);
        F2 w = F2::add(F2::dbl(XX), XX);
        F2 s = F2::mul(F2::dbl(Y), Z);
        F2 R = F2::mul(Y, s);
        line_coeffs(l0, lw, F2::mul(s, Z), F2::mul(w, Z), xP, yP);
        l3 = F2::sub(F2::mul(w, X), R);
        F2 RR = F2::sqr(R);
        F2 B = F2::sub(F2::sub(F2::sqr(F2::add(X, R)), XX), RR);
        F2 h = F2::sub(F2::sqr(w), F2::dbl(B));
        X = F2::mul(h, s);
        Y = F2::sub(F2::mul(w, F2::sub(B, h)), F2::dbl(RR));
        Z = F2::mul(s, F2::sqr(s));
      }
      f12_sqr(t, f);
      f12_mul_by_line(f, t, l0, lw, l3);
      if ((PP::loop(i >> 5) >> (i & 31)) & 1) add_step(f, X, Y, Z, xQ, yQ, xP, yP);
    }
    if (PP::OPT_ATE_BN) {
      F2 x1, y1, x2, y2;
      frob_twist(x1, y1, xQ, yQ);
      frob_twist(x2, y2, x1, y1);
      add_step(f, X, Y, Z, x1, y1, xP, yP);
      add_step(f, X, Y, Z, x2, F2::neg(y2), xP, yP);
    }
  }

  // ---- final exponentiation ---------------------------------------------------------------------------------------
  // f^(m (q^12 - 1) / r): easy part (q^6 - 1)(q^2 + 1), then the hard part (q^4 - q^2 + 1) / r by the x-chains
  //   BLS12 (m = 3):  3 h = (x - 1)^2 (x + q)(x^2 + q^2 - 1) + 3
  //   BN (m = 1):     h = q^3 + (6x^2 + 1) q^2 + (1 - 36x^3 - 18x^2 - 12x) q + (-36x^3 - 30x^2 - 18x - 2), evaluated with
  //                   the vectorial addition chain of Scott et al. (y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36)
  // Inverses in the cyclotomic subgroup are conjugates.
  static ZKB_NOINLINE void pow_x_signed(F12& r, const F12& a) {           // a^x for cyclotomic a
    f12_pow_x(r, a);
    if (PP::X_NEG) r = f12_conj(r);
  }
  static ZKB_NOINLINE void final_exponentiation(F12& out, const F12& fin) {
    F12 f, t, u;
    f12_inv(t, fin);
    f12_mul(u, f12_conj(fin), t);              // ^(q^6 - 1)
    f12_frob(t, u); f12_frob(f, t);
    f12_mul(t, f, u);                          // ^(q^2 + 1)
    f = t;
    if (PP::IS_BLS12) {
      F12 a, b, c;
      pow_x_signed(t, f); f12_mul(a, t, f12_conj(f));       // f^(x - 1)
      pow_x_signed(t, a); f12_mul(u, t, f12_conj(a));       // f^((x - 1)^2)
      a = u;
      pow_x_signed(t, a); f12_frob(u, a); f12_mul(b, t, u); // ^(x + q)
      pow_x_signed(t, b); pow_x_signed(u, t);               // b^(x^2)
      f12_frob(t, b); f12_frob(c, t);                       // b^(q^2)
      f12_mul(t, u, c); f12_mul(c, t, f12_conj(b));         // ^(x^2 + q^2 - 1)
      f12_cyclotomic_sqr(t, f); f12_mul(u, t, f);                      // f^3
      f12_mul(out, c, u);
    } else {
      F12 fp, fp2, fp3, fu, fu2, fu3, y0, y2, y3, y4, y6, t0, t1;
      f12_frob(fp, f); f12_frob(fp2, fp); f12_frob(fp3, fp2);
      pow_x_signed(fu, f); pow_x_signed(fu2, fu); pow_x_signed(fu3, fu2);
      f12_mul(t, fp, fp2); f12_mul(y0, t, fp3);
      f12_frob(t, fu2); f12_frob(y2, t);                     // fu2^(q^2)
      f12_frob(y3, fu); y3 = f12_conj(y3);
      f12_frob(t, fu2); f12_mul(y4, fu, t); y4 = f12_conj(y4);
      f12_frob(t, fu3); f12_mul(y6, fu3, t); y6 = f12_conj(y6);
      const F12 y1 = f12_conj(f), y5 = f12_conj(fu2);
      f12_cyclotomic_sqr(t0, y6); f12_mul(t, t0, y4); f12_mul(t0, t, y5);
      f12_mul(t, y3, y5); f12_mul(t1, t, t0);
      f12_mul(t, t0, y2); t0 = t;
      f12_cyclotomic_sqr(t, t1); f12_mul(t1, t, t0);
      f12_cyclotomic_sqr(t, t1); t1 = t;
      f12_mul(t0, t1, y1);
      f12_mul(t, t1, y0); t1 = t;
      f12_cyclotomic_sqr(t, t0);
      f12_mul(out, t, t1);
    }
  }

  ZKB_HD static void f12_from_mont(F12& a) {
    F2* g[6] = {&a.c0.a0, &a.c0.a1, &a.c0.a2, &a.c1.a0, &a.c1.a1, &a.c1.a2};
    for (int i = 0; i < 6; i++) {
      g[i]->c0.f = Fp<FqP>::from_mont(g[i]->c0.f);
      g[i]->c1.f = Fp<FqP>::from_mont(g[i]->c1.f);
    }
  }
};

#ifdef __CUDACC__
// one thread per pair: ml[i] = f(P_i, Q_i) (Montgomery form)
template <class PP>
__global__ void __launch_bounds__(64) k_miller_loops(const Affine<Fp<typename PP::FqP>>* g1, const uint8_t* g1_inf,
                                                     const Affine<Fp2<typename PP::FqP>>* g2, const uint8_t* g2_inf,
                                                     size_t n_pairs, typename PairingT<PP>::F12* ml) {
  using PT = PairingT<PP>;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  typename PT::FC xP, yP;
  typename PT::F2 xQ, yQ;
  xP.f = g1[i].x; yP.f = g1[i].y;
  xQ.c0.f = g2[i].x.c0; xQ.c1.f = g2[i].x.c1; yQ.c0.f = g2[i].y.c0; yQ.c1.f = g2[i].y.c1;
  bool pi = (g1_inf && g1_inf[i]) || g1[i].is_inf(), qi = (g2_inf && g2_inf[i]) || g2[i].is_inf();
  typename PT::F12 f;
  PT::miller_loop(f, xP, yP, pi, xQ, yQ, qi);
  ml[i] = f;
}

// large groups (a random-linear-combination batch check is ONE product over thousands of pairs): partial products of
// `chunk` consecutive loop values per thread, out[g * n_chunks + c]; applied repeatedly until a group is short
template <class PP>
__global__ void __launch_bounds__(64) k_gt_chunk_products(const typename PairingT<PP>::F12* in, size_t n_groups, size_t group,
                                                          size_t chunk, size_t n_chunks, typename PairingT<PP>::F12* out) {
  using PT = PairingT<PP>;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n_groups * n_chunks) return;
  const size_t g = t / n_chunks, c = t % n_chunks;
  const size_t lo = c * chunk, hi = lo + chunk < group ? lo + chunk : group;
  typename PT::F12 f = in[g * group + lo], u;
  for (size_t j = lo + 1; j < hi; j++) { PT::f12_mul(u, f, in[g * group + j]); f = u; }
  out[t] = f;
}

// one thread per group of `group` consecutive pairs: out[g] = final_exponentiation(prod ml[g * group + j]) (Montgomery form)
template <class PP>
__global__ void __launch_bounds__(64) k_pairing_finish(const typename PairingT<PP>::F12* ml, size_t n_groups, size_t group,
                                                       typename PairingT<PP>::F12* out) {
  using PT = PairingT<PP>;
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  typename PT::F12 f = ml[g * group], t;
  for (size_t j = 1; j < group; j++) { PT::f12_mul(t, f, ml[g * group + j]); f = t; }
  PT::final_exponentiation(t, f);
  out[g] = t;
}
#endif

}  // namespace zkb
