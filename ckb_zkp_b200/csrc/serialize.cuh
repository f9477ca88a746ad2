// Point decompression: ark-serialize 0.2 compressed short-Weierstrass points -> the ABI's x || y Montgomery layout.
//
// Replaces, for the key files the CLI reads (cli/src/zkp_prove.rs:36-50,117-124: `Parameters::<E>::deserialize`,
// groth16/src/lib.rs:81-91; marlin `IndexProverKey` / `UniversalParams`), the per-point work of ark-ec 0.2
// `GroupAffine::deserialize` (un-vendored; layout recalled):
//   * x as canonical little-endian bytes (Fq2: c0 then c1), flags in the two top bits of the LAST byte:
//     bit 7 = "y is the larger of (y, -y)" (Fq2 ordered by c1 then c0), bit 6 = point at infinity;
//   * y = sqrt(x^3 + b) -- both base fields are 3 mod 4, so sqrt(a) = a^((p+1)/4); Fq2 = Fq[u]/(u^2 + 1) by the
//     complex method (two Fq square roots and one inversion);
//   * optional subgroup check r * P == 0 (`is_in_correct_subgroup_assuming_on_curve`).
// Everything here is ZKB_HD so the host emulation (tests/host_emu) runs the same code on the CPU against the oracle.
#pragma once
#include "curve.cuh"

namespace zkb {

enum : uint8_t { kDecompOk = 0, kDecompNotCanonical = 1, kDecompNotOnCurve = 2, kDecompNotInSubgroup = 3 };

// a^e for e given as little-endian 32-bit limbs
template <class P>
ZKB_HD Fp<P> fp_pow_limbs(const Fp<P>& a, const uint32_t* e, int nlimbs) {
  Fp<P> r = Fp<P>::one();
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      if (started) r = Fp<P>::sqr(r);
      if ((e[i] >> b) & 1) { r = started ? Fp<P>::mul(r, a) : a; started = true; }
    }
  }
  return r;
}

// square root in a prime field with p = 3 (mod 4); false when a is a non-residue
template <class P>
ZKB_HD bool fp_sqrt(const Fp<P>& a, Fp<P>& out) {
  constexpr int N = P::N;
  uint32_t e[N];
  uint64_t carry = 1;                                   // e = (p + 1) >> 2
  for (int i = 0; i < N; i++) { uint64_t t = (uint64_t)P::mod(i) + carry; e[i] = (uint32_t)t; carry = t >> 32; }
  for (int i = 0; i < N - 1; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 30);
  e[N - 1] = (e[N - 1] >> 2) | ((uint32_t)carry << 30);
  out = fp_pow_limbs<P>(a, e, N);
  return Fp<P>::sqr(out) == a;
}

template <class P>
ZKB_HD bool fp2_sqrt(const Fp2<P>& a, Fp2<P>& out) {
  using F = Fp<P>;
  if (a.is_zero()) { out = a; return true; }
  if (a.c1.is_zero()) {                                 // a in Fq: sqrt(a) or sqrt(-a) * u (exactly one exists: -1 is a non-residue)
    F s;
    if (fp_sqrt<P>(a.c0, s)) { out = {s, F::zero()}; return true; }
    if (fp_sqrt<P>(F::neg(a.c0), s)) { out = {F::zero(), s}; return true; }
    return false;
  }
  F s;
  if (!fp_sqrt<P>(F::add(F::sqr(a.c0), F::sqr(a.c1)), s)) return false;      // the norm must be a square in Fq
  const F half = F::inv(F::add(F::one(), F::one()));
  F t = F::mul(F::add(a.c0, s), half), x0;
  if (!fp_sqrt<P>(t, x0)) {
    t = F::mul(F::sub(a.c0, s), half);
    if (!fp_sqrt<P>(t, x0)) return false;
  }
  F x1 = F::mul(a.c1, F::inv(F::add(x0, x0)));
  out = {x0, x1};
  return Fp2<P>::sqr(out) == a;
}

// canonical-integer order (ark `Ord for Fp`): is a > b ?
template <class P>
ZKB_HD bool fp_greater(const Fp<P>& a_mont, const Fp<P>& b_mont) {
  Fp<P> a = Fp<P>::from_mont(a_mont), b = Fp<P>::from_mont(b_mont);
  for (int i = P::N - 1; i >= 0; i--) {
    if (a.v[i] != b.v[i]) return a.v[i] > b.v[i];
  }
  return false;
}
template <class P>
ZKB_HD bool fp2_greater(const Fp2<P>& a, const Fp2<P>& b) {     // ark `Ord for QuadExtField`: c1 first, then c0
  if (a.c1 != b.c1) return fp_greater<P>(a.c1, b.c1);
  return fp_greater<P>(a.c0, b.c0);
}

// curve coefficient b of y^2 = x^3 + b per (base field, group)
template <class P> struct CurveB;
template <> struct CurveB<BlsFq> {
  ZKB_HD static Fp<BlsFq> g1() { return Fp<BlsFq>::small(4); }
  ZKB_HD static Fp2<BlsFq> g2() { return {Fp<BlsFq>::small(4), Fp<BlsFq>::small(4)}; }          // 4 (1 + u)
};
template <> struct CurveB<BnFq> {
  ZKB_HD static Fp<BnFq> g1() { return Fp<BnFq>::small(3); }
  ZKB_HD static Fp2<BnFq> g2() {                                                                   // 3 / (9 + u) = (27 - 3 u) / 82
    using F = Fp<BnFq>;
    F i82 = F::inv(F::small(82));
    return {F::mul(F::small(27), i82), F::neg(F::mul(F::small(3), i82))};
  }
};

template <class P> ZKB_HD Fp<P> curve_b(const Fp<P>*) { return CurveB<P>::g1(); }
template <class P> ZKB_HD Fp2<P> curve_b(const Fp2<P>*) { return CurveB<P>::g2(); }

// canonical little-endian bytes -> Montgomery element; false when the value is >= p
template <class P>
ZKB_HD bool fp_from_le_bytes(const uint8_t* bytes, uint8_t last_byte_mask, Fp<P>& out) {
  Fp<P> c;
  for (int i = 0; i < P::N; i++) {
    uint32_t b3 = bytes[4 * i + 3];
    if (i == P::N - 1) b3 &= last_byte_mask;
    c.v[i] = (uint32_t)bytes[4 * i] | ((uint32_t)bytes[4 * i + 1] << 8) | ((uint32_t)bytes[4 * i + 2] << 16) | (b3 << 24);
  }
  bool below = false;                                  // c < p ?
  for (int i = P::N - 1; i >= 0; i--) {
    if (c.v[i] != P::mod(i)) { below = c.v[i] < P::mod(i); break; }
  }
  out = Fp<P>::to_mont(c);
  return below;
}

// One G1 point: 4 * N bytes in, affine out (identity = the device's (0, 0)); returns a kDecomp* status
template <class P>
ZKB_HD uint8_t decompress_point(const uint8_t* bytes, const Fp<P>& b, Affine<Fp<P>>& out, bool& infinity) {
  using F = Fp<P>;
  const uint8_t flags = bytes[4 * P::N - 1];
  infinity = (flags & 0x40) != 0;
  out = Affine<F>::inf();
  if (infinity) return kDecompOk;
  F x, y;
  if (!fp_from_le_bytes<P>(bytes, 0x3f, x)) return kDecompNotCanonical;
  if (!fp_sqrt<P>(F::add(F::mul(F::sqr(x), x), b), y)) return kDecompNotOnCurve;
  F ny = F::neg(y);
  const bool greatest = (flags & 0x80) != 0;            // get_point_from_x(x, greatest)
  if (fp_greater<P>(ny, y) == greatest) y = ny;         // keep the larger root iff the flag is set
  out = {x, y};
  return kDecompOk;
}
// One G2 point: 8 * N bytes (x.c0 then x.c1, flags on the last byte of c1)
template <class P>
ZKB_HD uint8_t decompress_point(const uint8_t* bytes, const Fp2<P>& b, Affine<Fp2<P>>& out, bool& infinity) {
  using F2 = Fp2<P>;
  const uint8_t flags = bytes[8 * P::N - 1];
  infinity = (flags & 0x40) != 0;
  out = Affine<F2>::inf();
  if (infinity) return kDecompOk;
  F2 x, y;
  const bool ok0 = fp_from_le_bytes<P>(bytes, 0xff, x.c0);
  const bool ok1 = fp_from_le_bytes<P>(bytes + 4 * P::N, 0x3f, x.c1);
  if (!ok0 || !ok1) return kDecompNotCanonical;
  if (!fp2_sqrt<P>(F2::add(F2::mul(F2::sqr(x), x), b), y)) return kDecompNotOnCurve;
  F2 ny = F2::neg(y);
  const bool greatest = (flags & 0x80) != 0;
  if (fp2_greater<P>(ny, y) == greatest) y = ny;
  out = {x, y};
  return kDecompOk;
}

}  // namespace zkb
