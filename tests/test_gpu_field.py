"""GPU field / group arithmetic (the same code the MSM and NTT kernels inline) against the
Python big-int oracle, through the diagnostic C-ABI entry points."""
import random

import numpy as np
import pytest

from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, BLS_FQ, BLS_FR, BN_FQ, BN_FR, FQ
from tests import emu

pytestmark = pytest.mark.gpu

FIELDS = {0: (BN_FR, 8), 1: (BLS_FR, 8), 2: (BN_FQ, 8), 3: (BLS_FQ, 12)}


def pack(vals, n32):
    return np.stack([emu.to_u32(v, n32) for v in vals])


def unpack(arr):
    return [emu.from_u32(r) for r in arr]


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_field_ops(ctx, fid):
    fp, n = FIELDS[fid]
    p = fp.p
    R = (1 << (32 * n)) % p
    Rinv = pow(R, -1, p)
    rng = random.Random(100 + fid)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, 1 << (p.bit_length() - 1), 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]
    a = edge + [rng.randrange(p) for _ in range(500)]
    b = [a[(i * 7 + 3) % len(a)] for i in range(len(a))]
    A, B = pack(a, n), pack(b, n)
    run = lambda op: unpack(ctx.debug_fp_op(fid, op, A, B))
    assert run(0) == [x * y * Rinv % p for x, y in zip(a, b)]
    assert run(6) == [x * x * Rinv % p for x in a]
    assert run(1) == [(x + y) % p for x, y in zip(a, b)]
    assert run(2) == [(x - y) % p for x, y in zip(a, b)]
    assert run(7) == [(-x) % p for x in a]
    assert run(4) == [x * R % p for x in a]
    assert run(5) == [x * Rinv % p for x in a]
    nz = [x for x in a if x][:64]
    got = unpack(ctx.debug_fp_op(fid, 3, pack([x * R % p for x in nz], n), pack(nz, n)))
    assert got == [pow(x, -1, p) * R % p for x in nz]


def xyzz_pack(curve, cid, pts_jac, n32):
    """Jacobian oracle points -> XYZZ rows (X, Y, ZZ, ZZZ) in Montgomery form with a random Z."""
    F, fq = curve.F, FQ[cid]
    rows = []
    for (P, zval) in pts_jac:
        if P is None:
            coords = [F.zero] * 4
        else:
            zz = F.sqr(zval)
            zzz = F.mul(zz, zval)
            coords = [F.mul(P[0], zz), F.mul(P[1], zzz), zz, zzz]
        flat = []
        for c in coords:
            for e in (c if isinstance(c, tuple) else (c,)):
                flat.append(emu.to_u32(fq.to_mont(e), n32))
        rows.append(np.concatenate(flat))
    return np.stack(rows)


def aff_pack(cid, pts, n32, group):
    fq = FQ[cid]
    rows = []
    for P in pts:
        cs = [0] * (4 if group == 2 else 2) if P is None else (list(P[0]) + list(P[1]) if group == 2 else list(P))
        rows.append(np.concatenate([emu.to_u32(fq.to_mont(c), n32) for c in cs]))
    return np.stack(rows)


def aff_unpack(cid, rows, n32, group):
    fq = FQ[cid]
    out = []
    for r in rows:
        cs = [fq.from_mont(emu.from_u32(r[k * n32:(k + 1) * n32])) for k in range(len(r) // n32)]
        if all(c == 0 for c in cs):
            out.append(None)
        elif group == 2:
            out.append(((cs[0], cs[1]), (cs[2], cs[3])))
        else:
            out.append((cs[0], cs[1]))
    return out


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)])
def test_group_ops(ctx, cid, group):
    c = CURVES[(cid, group)]
    F = c.F
    n32 = FQ[cid].limbs * 2
    rng = random.Random(cid * 10 + group)
    rz = lambda: ((rng.randrange(1, F.p), rng.randrange(F.p)) if group == 2 else rng.randrange(1, F.p))
    ks = [rng.randrange(1, 1 << 64) for _ in range(24)]
    P = [c.mul_affine(c.gen, k) for k in ks]
    Q = [c.mul_affine(c.gen, k) for k in ks[1:] + ks[:1]]
    # special cases: identity operands, P + P (doubling branch), P + (-P)
    accs = [(p, rz()) for p in P] + [(None, None), (P[0], rz()), (P[1], rz()), (P[2], rz())]
    qs = Q + [Q[0], None, P[1], c.neg_affine(P[2])]
    want_add = [c.to_affine(c.add_mixed(c.from_affine(a[0]), q)) for a, q in zip(accs, qs)]
    acc_arr = xyzz_pack(c, cid, accs, n32)
    pt_words = acc_arr.shape[1]
    aff_words = pt_words // 2

    def affine_of(xyzz_rows):
        return aff_unpack(cid, ctx.debug_pt_op(cid, group, 3, xyzz_rows, None, False, aff_words), n32, group)

    got = ctx.debug_pt_op(cid, group, 0, acc_arr, aff_pack(cid, qs, n32, group), False, pt_words)
    assert affine_of(got) == want_add
    got = ctx.debug_pt_op(cid, group, 0, acc_arr, aff_pack(cid, qs, n32, group), True, pt_words)
    assert affine_of(got) == [c.to_affine(c.add_mixed(c.from_affine(a[0]), c.neg_affine(q))) for a, q in zip(accs, qs)]
    # full addition with both operands in XYZZ form
    q_xyzz = xyzz_pack(c, cid, [(q, rz() if q is not None else None) for q in qs], n32)
    got = ctx.debug_pt_op(cid, group, 1, acc_arr, q_xyzz, False, pt_words)
    assert affine_of(got) == want_add
    # doubling
    got = ctx.debug_pt_op(cid, group, 2, acc_arr, None, False, pt_words)
    assert affine_of(got) == [c.to_affine(c.dbl(c.from_affine(a[0]))) for a in accs]
    # scalar multiplication by 256-bit scalars
    sc = [rng.randrange(c.r) for _ in accs]
    sc[0], sc[1] = 0, 1
    k_arr = np.stack([emu.to_u32(k, 8) for k in sc])
    got = ctx.debug_pt_op(cid, group, 4, acc_arr, k_arr, False, pt_words)
    assert affine_of(got) == [c.to_affine(c.mul(c.from_affine(a[0]), k)) for a, k in zip(accs, sc)]
