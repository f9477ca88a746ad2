"""Marlin's polynomial-commitment prover side on the GPU (ckb_zkp_b200/kzg10.py) vs the oracle's
restatement of marlin/src/pc/kzg10.rs and marlin/src/pc/mod.rs; the shape of the cases follows the
reference's own tests (kzg10.rs:229-270, pc/mod.rs:253-418), with the pairing check replaced by its
restatement in the exponent (the trapdoor is known to the test)."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import kzg10 as zk
from oracle.pyref import kzg10 as OK
from oracle.pyref.fields import BLS12_381, BN254, FR
from tests import helpers as H

pytestmark = pytest.mark.gpu


class ReplayRng:
    """hands out a fixed list of values through randrange() and records nothing else"""

    def __init__(self, values):
        self.values, self.i = list(values), 0

    def randrange(self, _):
        v = self.values[self.i]
        self.i += 1
        return v


def make_key(ctx, cid, max_degree, supported, rng):
    mod = FR[cid].p
    pp = OK.setup(cid, max_degree, rng.randrange(mod), g_scalar=rng.randrange(1, mod), gamma=rng.randrange(1, mod))
    ock = OK.trim(pp, supported)
    ck = zk.CommitterKey(ctx, cid, H.points_array(cid, 1, ock["powers_of_g"]), H.points_array(cid, 1, ock["powers_of_gamma_g"]),
                         supported)
    return ock, ck


def pt(cid, got):
    return H.array_point(cid, 1, got[0], got[1])


@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_poly_helpers(ctx, cid):
    mod = FR[cid].p
    rng = random.Random(cid)
    for n in (1, 2, 3, 63, 64, 65, 200, 4096, 4097, 70000):
        p = [rng.randrange(mod) for _ in range(n)]
        if n > 5:
            p[0] = p[1] = 0
            p[-1] = 0
        z = rng.randrange(mod)
        q, rem = ctx.poly_div_linear(cid, H.fr_array(cid, p), H.fr_array(cid, [z])[0])
        assert H.fr_ints(cid, rem.reshape(1, 4))[0] == OK.poly_eval(p, z, mod)
        if n <= 4097:
            assert H.fr_ints(cid, q) == OK.poly_div_linear(p, z, mod)
        else:       # (x - z) * q + rem == p at a random point
            x = rng.randrange(mod)
            qv = H.fr_ints(cid, ctx.poly_eval(cid, q, H.fr_array(cid, [x])[0]).reshape(1, 4))[0]
            assert ((x - z) * qv + OK.poly_eval(p, z, mod)) % mod == OK.poly_eval(p, x, mod)
    # z = 0 and z = 1
    p = [rng.randrange(mod) for _ in range(130)]
    for z in (0, 1):
        q, rem = ctx.poly_div_linear(cid, H.fr_array(cid, p), H.fr_array(cid, [z])[0])
        assert H.fr_ints(cid, q) == OK.poly_div_linear(p, z, mod)
    # batch inversion keeps zeros (ark_ff::batch_inversion)
    a = [rng.randrange(mod) for _ in range(1000)]
    a[0] = a[17] = a[64] = a[999] = 0
    inv = H.fr_ints(cid, ctx.fr_batch_inverse(cid, H.fr_array(cid, a)))
    assert inv == [pow(x, -1, mod) if x else 0 for x in a]
    # linear combination with shifts
    polys = [[rng.randrange(mod) for _ in range(k)] for k in (5, 1, 33)]
    cs = [rng.randrange(mod) for _ in polys]
    shifts = [0, 7, 2]
    want = [0] * 40
    for pl, c, s in zip(polys, cs, shifts):
        for i, v in enumerate(pl):
            want[i + s] = (want[i + s] + c * v) % mod
    got = ctx.poly_lincomb(cid, [H.fr_array(cid, pl) for pl in polys], H.fr_array(cid, cs), shifts, out_len=40)
    assert H.fr_ints(cid, got) == want


@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_kzg10_commit_open(ctx, cid):
    """kzg10.rs:234-262: random degree < 20, trim to degree/2, commit with hiding bound 1, open, check"""
    mod = FR[cid].p
    rng = random.Random(20 + cid)
    for _ in range(6):
        degree = rng.randrange(2, 20)
        sup = max(degree // 2, 1)
        ock, ck = make_key(ctx, cid, degree, sup, rng)
        p = [rng.randrange(mod) for _ in range(sup + 1)]
        blind = [rng.randrange(mod) for _ in range(2)]
        comm, rand = zk.kzg_commit(ck, H.fr_array(cid, p), hiding_bound=1, rng=ReplayRng(blind), supported_degree=sup)
        assert pt(cid, comm) == OK.kzg_commit(cid, ock["powers_of_g"], ock["powers_of_gamma_g"], p, blind, sup)
        point = rng.randrange(mod)
        w, rand_v = zk.kzg_open(ck, H.fr_array(cid, p), H.fr_array(cid, [point])[0], rand)
        ow, orv = OK.kzg_open(cid, ock["powers_of_g"], ock["powers_of_gamma_g"], p, point, blind)
        assert pt(cid, w) == ow
        assert H.fr_ints(cid, rand_v.reshape(1, 4))[0] == orv
        # the reference's acceptance test (KZG10::check), in the exponent
        ce = OK.commitment_exponent(ock, p, blind)
        we = OK.commitment_exponent(ock, OK.poly_div_linear(p, point, mod), OK.poly_div_linear(blind, point, mod))
        assert pt(cid, comm) == OK.exponent_point(ock, ce) and pt(cid, w) == OK.exponent_point(ock, we)
        assert OK.kzg_check_in_exponent(ock, ce, point, OK.poly_eval(p, point, mod), we, orv)
        ck.free()


def test_kzg10_errors(ctx):
    cid = BLS12_381
    mod = FR[cid].p
    rng = random.Random(5)
    ock, ck = make_key(ctx, cid, 8, 4, rng)
    const = H.fr_array(cid, [5, 0, 0])
    with pytest.raises(zk.DegreeIsZero):
        zk.kzg_commit(ck, const, supported_degree=4)
    with pytest.raises(zk.DegreeOutOfBound):
        zk.kzg_commit(ck, H.fr_array(cid, [1] * 6), supported_degree=4)
    with pytest.raises(zk.MissingRng):
        zk.kzg_commit(ck, H.fr_array(cid, [1, 2]), hiding_bound=1, supported_degree=4)
    with pytest.raises(zk.HidingBoundIsZero):
        zk.kzg_commit(ck, H.fr_array(cid, [1, 2]), hiding_bound=0, rng=ReplayRng([1]), supported_degree=4)
    # no hiding: commitment of a polynomial with zero low coefficients (skip_leading_zeros)
    p = [0, 0, 3, 4]
    comm, rand = zk.kzg_commit(ck, H.fr_array(cid, p), supported_degree=4)
    assert not rand.is_hiding()
    assert pt(cid, comm) == OK.kzg_commit(cid, ock["powers_of_g"], ock["powers_of_gamma_g"], p, None, 4)
    ck.free()


@pytest.mark.parametrize("enforce_bounds", [False, True])
def test_pc_single_point_and_batch(ctx, enforce_bounds):
    """pc/mod.rs single_point_template / batch_template: several labelled polynomials, hiding bound 1,
    optional degree bounds (shifted commitments), one opening per query point."""
    cid = BLS12_381
    mod = FR[cid].p
    rng = random.Random(77 + enforce_bounds)
    max_degree, sup, n_polys = 24, 20, 4
    ock, ck = make_key(ctx, cid, max_degree, sup, rng)
    opening_challenge = rng.randrange(mod)
    labeled, opolys, draws = [], [], []
    for i in range(n_polys):
        degree = rng.randrange(1, sup + 1)
        coeffs = [rng.randrange(mod) for _ in range(degree + 1)]
        if coeffs[-1] == 0:
            coeffs[-1] = 1
        db = degree if enforce_bounds else None
        b1 = [rng.randrange(mod) for _ in range(2)]
        b2 = [rng.randrange(mod) for _ in range(2)] if enforce_bounds else None
        draws += b1 + (b2 or [])
        labeled.append(zk.LabeledPolynomial(str(i), H.fr_array(cid, coeffs), db, 1))
        opolys.append({"coeffs": coeffs, "degree_bound": db, "blinding": b1, "shifted_blinding": b2})
    comms, rands = zk.pc_commit(ck, labeled, ReplayRng(draws))
    ocomms = OK.pc_commit(ock, opolys)
    for (c, sc), (oc, osc) in zip(comms, ocomms):
        assert pt(cid, c) == oc
        assert (sc is None) == (osc is None)
        if sc is not None:
            assert pt(cid, sc) == osc
    point = rng.randrange(mod)
    w, rand_v = zk.pc_open(ck, labeled, H.fr_array(cid, [point])[0], opening_challenge, rands)
    ow, orv = OK.pc_open(ock, opolys, point, opening_challenge)
    assert pt(cid, w) == ow and H.fr_ints(cid, rand_v.reshape(1, 4))[0] == orv
    # batch_open over three points with overlapping label sets
    pts3 = sorted(rng.randrange(mod) for _ in range(3))
    query = [("0", pts3[0]), ("1", pts3[0]), ("2", pts3[1]), ("3", pts3[2]), ("0", pts3[2]), ("1", pts3[2])]
    proofs = zk.pc_batch_open(ck, labeled, query, opening_challenge, rands)
    groups = {pts3[0]: [0, 1], pts3[1]: [2], pts3[2]: [0, 1, 3]}
    for pr, point in zip(proofs, pts3):
        ow, orv = OK.pc_open(ock, [opolys[i] for i in groups[point]], point, opening_challenge)
        assert pt(cid, pr[0]) == ow and H.fr_ints(cid, pr[1].reshape(1, 4))[0] == orv
    with pytest.raises(zk.MissingPolynomial):
        zk.pc_batch_open(ck, labeled, [("nope", 1)], opening_challenge, rands)
    ck.free()


def test_kzg_commit_large(ctx):
    """Marlin-sized commitment (2^16 coefficients, BN254): GPU commit == (p(beta)) * g computed from
    the trapdoor, with the key generated on the GPU from the powers of beta."""
    from ckb_zkp_b200 import synth
    cid, n = BN254, 1 << 16
    mod = FR[cid].p
    rng = random.Random(3)
    beta = rng.randrange(mod)
    powers, cur = [], 1
    for _ in range(n):
        powers.append(cur)
        cur = cur * beta % mod
    gen = synth.generator_mont(cid, 1)
    xy, inf = ctx.fixed_base_mul(cid, 1, gen, H.ints_to_u64(powers, 4))
    ck = zk.CommitterKey(ctx, cid, (xy, inf), (xy[:4], inf[:4]))
    coeffs = [rng.randrange(mod) for _ in range(n)]
    coeffs[0] = coeffs[1] = 0
    comm, _ = zk.kzg_commit(ck, H.fr_array(cid, coeffs))
    e = OK.poly_eval(coeffs, beta, mod)
    want, winf = ctx.fixed_base_mul(cid, 1, gen, H.ints_to_u64([e], 4))
    assert not comm[1] and np.array_equal(comm[0], want[0])
    point = rng.randrange(mod)
    w, rand_v = zk.kzg_open(ck, H.fr_array(cid, coeffs), H.fr_array(cid, [point])[0], zk.Randomness())
    assert rand_v is None
    qe = (e - OK.poly_eval(coeffs, point, mod)) * pow(beta - point, -1, mod) % mod
    want, _ = ctx.fixed_base_mul(cid, 1, gen, H.ints_to_u64([qe], 4))
    assert np.array_equal(w[0], want[0])
    ck.free()


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_poly_eval_batch(ctx, cid):
    """zkb_poly_eval_batch: lengths 0, 1, one chunk, several tree levels; host and device-resident coefficients; each
    polynomial at its own point -- against Horner on integers and against the single-call entry"""
    import torch
    mod = FR[cid].p
    rng = random.Random(17 + cid)
    lens = [0, 1, 2, 127, 128, 129, 5000, 40000]
    polys = [[rng.randrange(mod) for _ in range(n)] for n in lens]
    points = [rng.randrange(mod) for _ in lens]
    points[3] = 0
    arrs = [H.fr_array(cid, p) for p in polys]
    arrs[6] = torch.from_numpy(arrs[6].view(np.int64)).cuda()
    got = ctx.poly_eval_batch(cid, arrs, H.fr_array(cid, points))
    want = [OK.poly_eval(p, z, mod) for p, z in zip(polys, points)]
    assert H.fr_ints(cid, got) == want
    for j in (1, 5, 7):
        assert np.array_equal(ctx.poly_eval(cid, arrs[j], H.fr_array(cid, [points[j]])[0]), got[j])
    assert ctx.poly_eval_batch(cid, [], np.zeros((0, 4), dtype=np.uint64)).shape == (0, 4)
    with pytest.raises(ValueError):
        ctx.poly_eval_batch(cid, arrs[:2], H.fr_array(cid, points[:3]))
