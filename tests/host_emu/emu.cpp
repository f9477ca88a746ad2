// Host-emulation harness: compiles the device headers (field.cuh / curve.cuh) with a
// host compiler -- ptx.cuh then emulates every carry-chain instruction bit-faithfully --
// so the exact instruction sequences can be checked against the Python oracle without
// a GPU.  Test infrastructure only; not linked into the product library.
#include <cstring>
#include "../../ckb_zkp_b200/csrc/curve.cuh"
#include "../../ckb_zkp_b200/csrc/serialize.cuh"
#include "../../ckb_zkp_b200/csrc/pairing.cuh"

using namespace zkb;

template <class F> static void fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  F x, y, r;
  memcpy(&x, a, sizeof(F));
  memcpy(&y, b, sizeof(F));
  switch (op) {
    case 0: r = F::mul(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::inv(x); break;
    case 4: r = F::to_mont(x); break;
    case 5: r = F::from_mont(x); break;
    case 6: r = F::sqr(x); break;
    case 7: r = F::neg(x); break;
    case 8: r = F::mul_sos(x, y); break;
    case 9: r = F::inv_safegcd(x); break;
    default: r = F::zero();
  }
  memcpy(out, &r, sizeof(F));
}

// op: 0 = acc.madd(q_affine, neg), 1 = acc.add(q_xyzz), 2 = dbl(acc), 3 = to_affine(acc) (out = x,y),
//     4 = mul_limbs(acc, k[8])
template <class F> static void pt_op(int op, const uint32_t* acc, const uint32_t* q, int neg, uint32_t* out) {
  XYZZ<F> a;
  memcpy(&a, acc, sizeof(a));
  if (op == 0) {
    Affine<F> p;
    memcpy(&p, q, sizeof(p));
    a.madd(p, neg != 0);
    memcpy(out, &a, sizeof(a));
  } else if (op == 1) {
    XYZZ<F> b;
    memcpy(&b, q, sizeof(b));
    a.add(b);
    memcpy(out, &a, sizeof(a));
  } else if (op == 2) {
    a = XYZZ<F>::dbl(a);
    memcpy(out, &a, sizeof(a));
  } else if (op == 3) {
    Affine<F> r = a.to_affine();
    memcpy(out, &r, sizeof(r));
  } else if (op == 4) {
    a = XYZZ<F>::mul_limbs(a, q, 8);
    memcpy(out, &a, sizeof(a));
  }
}

// point decompression (serialize.cuh): bytes -> affine Montgomery limbs; returns the kDecomp* status, *inf = identity flag.
// check_subgroup: r * P == 0 with the scalar-field modulus limbs, like the device kernel.
template <class F, class FrP> static int decompress_one(const uint8_t* bytes, int check_subgroup, uint32_t* out, uint8_t* inf) {
  Affine<F> p;
  bool is_inf = false;
  F b = curve_b((const F*)nullptr);
  uint8_t st = decompress_point(bytes, b, p, is_inf);
  if (st == kDecompOk && !is_inf && check_subgroup) {
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = FrP::mod(i);
    XYZZ<F> q = XYZZ<F>::mul_limbs(XYZZ<F>::from_affine(p), r, 8);
    if (!q.is_inf()) st = kDecompNotInSubgroup;
  }
  memcpy(out, &p, sizeof(p));
  *inf = is_inf ? 1 : 0;
  return st;
}


// pairing.cuh: stage 0 = affine Miller loop only, 1 = final exponentiation of `f_in`, 2 = projective loop + final
// exponentiation (what the kernels run), 3 = affine loop + final exponentiation.  Points are affine Montgomery
// limbs (G1: x,y; G2: x.c0,x.c1,y.c0,y.c1); f_in / out are 12 Fq in tower order, canonical (non-Montgomery) limbs.
template <class PP> static void pairing_stage(int stage, const uint32_t* p, const uint32_t* q, const uint32_t* f_in, uint32_t* out) {
  using PT = PairingT<PP>;
  using Fq = Fp<typename PP::FqP>;
  typename PT::F12 f, r;
  if (stage == 1) {
    memcpy(&f, f_in, sizeof(f));
    Fq* e = (Fq*)&f;
    for (int i = 0; i < 12; i++) e[i] = Fq::to_mont(e[i]);
  } else {
    typename PT::FC xP, yP;
    typename PT::F2 xQ, yQ;
    memcpy(&xP, p, sizeof(xP)); memcpy(&yP, p + Fq::N, sizeof(yP));
    memcpy(&xQ, q, sizeof(xQ)); memcpy(&yQ, q + 2 * Fq::N, sizeof(yQ));
    bool pi = xP.is_zero() && yP.is_zero(), qi = xQ.is_zero() && yQ.is_zero();
    if (stage == 0 || stage == 3) PT::miller_loop_affine(f, xP, yP, pi, xQ, yQ, qi);
    else PT::miller_loop(f, xP, yP, pi, xQ, yQ, qi);        // stage 2: the inversion-free loop the kernels run
  }
  if (stage >= 1) { PT::final_exponentiation(r, f); f = r; }
  PT::f12_from_mont(f);
  memcpy(out, &f, sizeof(f));
}

extern "C" {
// field: 0 BnFr, 1 BlsFr, 2 BnFq, 3 BlsFq, 4 BnFq2, 5 BlsFq2
void emu_fp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: fp_op<Fp<BnFr>>(op, a, b, out); break;
    case 1: fp_op<Fp<BlsFr>>(op, a, b, out); break;
    case 2: fp_op<Fp<BnFq>>(op, a, b, out); break;
    case 3: fp_op<Fp<BlsFq>>(op, a, b, out); break;
  }
}
void emu_fp2_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (curve == 0) {
    using F = Fp2<BnFq>; F x, y, r; memcpy(&x, a, sizeof(F)); memcpy(&y, b, sizeof(F));
    r = op == 0 ? F::mul(x, y) : op == 6 ? F::sqr(x) : op == 3 ? F::inv(x) : F::zero();
    memcpy(out, &r, sizeof(F));
  } else {
    using F = Fp2<BlsFq>; F x, y, r; memcpy(&x, a, sizeof(F)); memcpy(&y, b, sizeof(F));
    r = op == 0 ? F::mul(x, y) : op == 6 ? F::sqr(x) : op == 3 ? F::inv(x) : F::zero();
    memcpy(out, &r, sizeof(F));
  }
}
// curve: 0 BN254, 1 BLS12-381; group: 1 or 2
void emu_pt_op(int curve, int group, int op, const uint32_t* acc, const uint32_t* q, int neg, uint32_t* out) {
  if (curve == 0 && group == 1) pt_op<Fp<BnFq>>(op, acc, q, neg, out);
  if (curve == 0 && group == 2) pt_op<Fp2<BnFq>>(op, acc, q, neg, out);
  if (curve == 1 && group == 1) pt_op<Fp<BlsFq>>(op, acc, q, neg, out);
  if (curve == 1 && group == 2) pt_op<Fp2<BlsFq>>(op, acc, q, neg, out);
}

int emu_decompress(int curve, int group, const uint8_t* bytes, int check_subgroup, uint32_t* out, uint8_t* inf) {
  if (curve == 0 && group == 1) return decompress_one<Fp<BnFq>, BnFr>(bytes, check_subgroup, out, inf);
  if (curve == 0 && group == 2) return decompress_one<Fp2<BnFq>, BnFr>(bytes, check_subgroup, out, inf);
  if (curve == 1 && group == 1) return decompress_one<Fp<BlsFq>, BlsFr>(bytes, check_subgroup, out, inf);
  if (curve == 1 && group == 2) return decompress_one<Fp2<BlsFq>, BlsFr>(bytes, check_subgroup, out, inf);
  return -1;
}

void emu_pairing(int curve, int stage, const uint32_t* p, const uint32_t* q, const uint32_t* f_in, uint32_t* out) {
  if (curve == 0) pairing_stage<BnPairing>(stage, p, q, f_in, out);
  else pairing_stage<BlsPairing>(stage, p, q, f_in, out);
}
}
