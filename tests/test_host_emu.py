"""Device field / curve code (exact PTX carry-chain sequences, emulated on the host)
against the Python big-int oracle.  No GPU needed."""
import random

import numpy as np
import pytest

from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, BLS_FQ, BLS_FR, BN_FQ, BN_FR
from tests import emu

FIELDS = {0: (BN_FR, 8), 1: (BLS_FR, 8), 2: (BN_FQ, 8), 3: (BLS_FQ, 12)}


@pytest.fixture(scope="module")
def lib():
    return emu.build()


def edge_values(p):
    return [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (1 << (p.bit_length() - 1)), 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_field_ops(lib, fid):
    fp, n = FIELDS[fid]
    p, R = fp.p, (1 << (32 * n)) % fp.p
    assert R == fp.R
    Rinv = pow(R, -1, p)
    rng = random.Random(fid)
    vals = edge_values(p) + [rng.randrange(p) for _ in range(200)]
    out = np.zeros(n, dtype=np.uint32)

    def run(op, a, b=0):
        A, B = emu.to_u32(a, n), emu.to_u32(b, n)
        lib.emu_fp_op(fid, op, emu.ptr(A), emu.ptr(B), emu.ptr(out))
        return emu.from_u32(out)

    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert run(0, a, b) == a * b * Rinv % p
        assert run(6, a) == a * a * Rinv % p
        assert run(1, a, b) == (a + b) % p
        assert run(2, a, b) == (a - b) % p
        assert run(7, a) == (-a) % p
        assert run(4, a) == a * R % p
        assert run(5, a) == a * Rinv % p
    # inv in Montgomery domain: inv(aR) = a^-1 R (binary extended Euclid), 0 -> 0; powers of two and
    # p - 2^k walk the round count from its minimum to its maximum
    bits = p.bit_length()
    for a in vals + [1 << k for k in range(0, bits - 1, 37)] + [p - (1 << k) for k in range(0, bits - 1, 41)]:
        a %= p
        assert run(3, a * R % p) == (pow(a, -1, p) * R % p if a else 0)
        # the division-step inversion (safegcd, 30 steps per batch) has the same contract
        assert run(9, a * R % p) == (pow(a, -1, p) * R % p if a else 0)
    for _ in range(3000):
        a = rng.randrange(1, p)
        assert run(9, a) == pow(a * Rinv % p, -1, p) * R % p


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_fp2_ops(lib, cid):
    fq = {BN254: BN_FQ, BLS12_381: BLS_FQ}[cid]
    n = fq.limbs * 2
    p, R = fq.p, fq.R
    F2 = CURVES[(cid, 2)].F
    rng = random.Random(cid)
    out = np.zeros(2 * n, dtype=np.uint32)

    def enc(a):
        return np.concatenate([emu.to_u32(a[0] * R % p, n), emu.to_u32(a[1] * R % p, n)])

    def dec(o):
        return (emu.from_u32(o[:n]) * fq.Rinv % p, emu.from_u32(o[n:]) * fq.Rinv % p)

    for _ in range(50):
        a = (rng.randrange(p), rng.randrange(p))
        b = (rng.randrange(p), rng.randrange(p))
        A, B = enc(a), enc(b)
        lib.emu_fp2_op(cid, 0, emu.ptr(A), emu.ptr(B), emu.ptr(out))
        assert dec(out) == F2.mul(a, b)
        lib.emu_fp2_op(cid, 6, emu.ptr(A), emu.ptr(B), emu.ptr(out))
        assert dec(out) == F2.sqr(a)
    lib.emu_fp2_op(cid, 3, emu.ptr(A), emu.ptr(B), emu.ptr(out))
    assert dec(out) == F2.inv(a)


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)])
def test_point_ops(lib, cid, group):
    cv = CURVES[(cid, group)]
    fq = {BN254: BN_FQ, BLS12_381: BLS_FQ}[cid]
    n = fq.limbs * 2
    p, R = fq.p, fq.R
    fe = n * group                         # u32 words per coordinate
    rng = random.Random(cid * 2 + group)

    def enc_f(a):
        if group == 1:
            return emu.to_u32(a * R % p, n)
        return np.concatenate([emu.to_u32(a[0] * R % p, n), emu.to_u32(a[1] * R % p, n)])

    def dec_f(o):
        if group == 1:
            return emu.from_u32(o) * fq.Rinv % p
        return (emu.from_u32(o[:n]) * fq.Rinv % p, emu.from_u32(o[n:]) * fq.Rinv % p)

    def enc_aff(P):
        if P is None:
            return np.zeros(2 * fe, dtype=np.uint32)
        return np.concatenate([enc_f(P[0]), enc_f(P[1])])

    def dec_aff(o):
        if not o.any():
            return None
        return (dec_f(o[:fe]), dec_f(o[fe:]))

    def affine_of(acc):
        o = np.zeros(2 * fe, dtype=np.uint32)
        lib.emu_pt_op(cid, group, 3, emu.ptr(acc), emu.ptr(acc), 0, emu.ptr(o))
        return dec_aff(o)

    G = cv.gen
    pts = [cv.mul_affine(G, rng.randrange(1, cv.r)) for _ in range(6)]
    acc = np.zeros(4 * fe, dtype=np.uint32)            # identity
    ref = cv.identity()
    assert affine_of(acc) is None
    # sequence covering: inf+P, P+Q, P+P (doubling through madd), P+(-P), adding identity
    seq = [(pts[0], 0), (pts[1], 0), (pts[2], 1), (None, 0), (pts[3], 0)]
    for P, neg in seq:
        q = enc_aff(P)
        out = np.zeros(4 * fe, dtype=np.uint32)
        lib.emu_pt_op(cid, group, 0, emu.ptr(acc), emu.ptr(q), neg, emu.ptr(out))
        acc = out
        ref = cv.add_mixed(ref, cv.neg_affine(P) if neg else P)
        assert affine_of(acc) == cv.to_affine(ref)
    # doubling via madd: acc = P then += P
    one = np.zeros(4 * fe, dtype=np.uint32)
    q = enc_aff(pts[4])
    a1 = np.zeros(4 * fe, dtype=np.uint32)
    lib.emu_pt_op(cid, group, 0, emu.ptr(one), emu.ptr(q), 0, emu.ptr(a1))
    a2 = np.zeros(4 * fe, dtype=np.uint32)
    lib.emu_pt_op(cid, group, 0, emu.ptr(a1), emu.ptr(q), 0, emu.ptr(a2))
    assert affine_of(a2) == cv.mul_affine(pts[4], 2)
    a3 = np.zeros(4 * fe, dtype=np.uint32)
    lib.emu_pt_op(cid, group, 0, emu.ptr(a2), emu.ptr(q), 1, emu.ptr(a3))   # 2P - P
    assert affine_of(a3) == pts[4]
    a4 = np.zeros(4 * fe, dtype=np.uint32)
    lib.emu_pt_op(cid, group, 0, emu.ptr(a3), emu.ptr(q), 1, emu.ptr(a4))   # P - P = inf
    assert affine_of(a4) is None
    # full add: acc + a2, acc + acc (doubling), acc + (-acc)
    o = np.zeros(4 * fe, dtype=np.uint32)
    lib.emu_pt_op(cid, group, 1, emu.ptr(acc), emu.ptr(a2), 0, emu.ptr(o))
    assert affine_of(o) == cv.to_affine(cv.add(ref, cv.mul(cv.from_affine(pts[4]), 2)))
    lib.emu_pt_op(cid, group, 1, emu.ptr(acc), emu.ptr(acc), 0, emu.ptr(o))
    assert affine_of(o) == cv.to_affine(cv.dbl(ref))
    lib.emu_pt_op(cid, group, 2, emu.ptr(acc), emu.ptr(acc), 0, emu.ptr(o))
    assert affine_of(o) == cv.to_affine(cv.dbl(ref))
    # scalar multiplication
    k = rng.randrange(cv.r)
    K = emu.to_u32(k, 8)
    lib.emu_pt_op(cid, group, 4, emu.ptr(a1), emu.ptr(K), 0, emu.ptr(o))
    assert affine_of(o) == cv.mul_affine(pts[4], k)


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_field_mul_karatsuba_sos(lib, fid):
    """Fp::mul_sos (one Karatsuba level + separated Montgomery reduction) == a b R^-1 mod p, on edge values that
    maximise the half sums' carries (all-ones halves) and on random operands."""
    fp, n = FIELDS[fid]
    p = fp.p
    Rinv = pow((1 << (32 * n)) % p, -1, p)
    rng = random.Random(100 + fid)
    half = 16 * n
    vals = edge_values(p) + [((1 << half) - 1), (((1 << half) - 1) << half) % p, p - (1 << half), (1 << (32 * n - 1)) % p]
    vals += [rng.randrange(p) for _ in range(300)]
    out = np.zeros(n, dtype=np.uint32)
    for i, a in enumerate(vals):
        for b in (vals[(i * 5 + 1) % len(vals)], a, p - 1):
            A, B = emu.to_u32(a % p, n), emu.to_u32(b % p, n)
            lib.emu_fp_op(fid, 8, emu.ptr(A), emu.ptr(B), emu.ptr(out))
            assert emu.from_u32(out) == (a % p) * (b % p) * Rinv % p, (hex(a), hex(b))
