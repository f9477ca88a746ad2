"""Batched pairing checks on the GPU (zkb_multi_pairing, csrc/pairing.cuh) against the oracle's pairing
(oracle/pyref/pairing.py), and the reference's verifiers on top of it: groth16/src/verifier.rs:8-44 (one proof and a
batch), marlin/src/pc/kzg10.rs:158-173 (`check`, with the reference's own unit-test template kzg10.rs:235-270)."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import _lib
from ckb_zkp_b200 import generator as zgen
from ckb_zkp_b200 import groth16 as zg
from ckb_zkp_b200 import kzg10 as zk
from ckb_zkp_b200 import marlin as zm
from ckb_zkp_b200 import pairing as zp
from ckb_zkp_b200 import verifier as zv
from oracle.pyref import pairing as OP
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FQ, FR
from tests import helpers as H
from tests.test_gpu_generator import MiniCircuit

pytestmark = pytest.mark.gpu

W_POWER = [0, 2, 4, 1, 3, 5]           # tower slot (c0.a0, c0.a1, c0.a2, c1.a0, c1.a1, c1.a2) -> power of w


def gt_to_flat(cid, gt):
    """uint64[12 * limbs] (Montgomery, tower order) -> the oracle's 12 coefficients over Fq[w] / (w^12 - 2 c w^6 + c^2 + 1)"""
    q, L = FQ[cid].p, FQ[cid].limbs
    rinv = pow(1 << (64 * L), -1, q)
    t = [H.u64_to_int(gt[i * L:(i + 1) * L]) * rinv % q for i in range(12)]
    c = OP._C[cid]
    flat = [0] * 12
    for s in range(6):
        x, y, k = t[2 * s], t[2 * s + 1], W_POWER[s]
        flat[k] = (flat[k] + x - c * y) % q
        flat[k + 6] = (flat[k + 6] + y) % q
    return flat


def oracle_gt(cid, pairs):
    """the oracle's restatement of what the device computes: plain ate cubed on BLS12-381, optimal ate on BN254"""
    return OP.device_multi_pairing(cid, pairs)


def arr1(cid, P):
    xy, inf = H.points_array(cid, 1, [P])
    return xy[0], bool(inf[0])


def arr2(cid, Q):
    xy, inf = H.points_array(cid, 2, [Q])
    return xy[0], bool(inf[0])


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_multi_pairing_matches_oracle(ctx, cid):
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    rng = random.Random(31 + cid)
    groups = []
    for g in range(5):
        pairs = [(g1.mul_affine(g1.gen, rng.randrange(1, g1.r)), g2.mul_affine(g2.gen, rng.randrange(1, g2.r))) for _ in range(3)]
        if g == 1:
            pairs[2] = (None, pairs[2][1])             # an identity on either side contributes 1
        if g == 2:
            pairs[0] = (pairs[0][0], None)
        groups.append(pairs)
    got = zp.multi_pairing(ctx, cid, [[(arr1(cid, P), arr2(cid, Q)) for P, Q in g] for g in groups])
    for g, pairs in zip(got, groups):
        assert gt_to_flat(cid, g) == oracle_gt(cid, pairs)
    # all-identity group -> 1
    one = zp.multi_pairing(ctx, cid, [[(arr1(cid, None), arr2(cid, groups[0][0][1]))]])[0]
    assert np.array_equal(one, zp.gt_one(cid))


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_bilinearity_and_batch(ctx, cid):
    """e(aP, bQ) == e(abP, Q) == e(P, abQ), e(P, Q) e(-P, Q) == 1, over a batch large enough to fill several blocks"""
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    rng = random.Random(77)
    n = 200
    ks = [(rng.randrange(1, g1.r), rng.randrange(1, g1.r)) for _ in range(n)]
    can = lambda v: H.ints_to_u64(v, 4)
    gen1, gen2 = arr1(cid, g1.gen)[0], arr2(cid, g2.gen)[0]
    aP = ctx.fixed_base_mul(cid, _lib.G1, gen1, can([a for a, _ in ks]))
    bQ = ctx.fixed_base_mul(cid, _lib.G2, gen2, can([b for _, b in ks]))
    abP = ctx.fixed_base_mul(cid, _lib.G1, gen1, can([a * b % g1.r for a, b in ks]))
    abQ = ctx.fixed_base_mul(cid, _lib.G2, gen2, can([a * b % g1.r for a, b in ks]))
    e1 = ctx.multi_pairing(cid, aP, bQ, 1)
    e2 = ctx.multi_pairing(cid, abP, (np.tile(gen2, (n, 1)), None), 1)
    e3 = ctx.multi_pairing(cid, (np.tile(gen1, (n, 1)), None), abQ, 1)
    assert np.array_equal(e1, e2) and np.array_equal(e1, e3)
    assert len({e1[i].tobytes() for i in range(n)}) == n                       # non-degenerate: distinct exponents, distinct values
    negs = np.stack([zp.neg_point(cid, _lib.G1, (aP[0][i], False))[0] for i in range(n)])
    both = ctx.multi_pairing(cid, (np.stack([aP[0], negs], axis=1).reshape(2 * n, -1), None), (np.repeat(bQ[0], 2, axis=0), None), 2)
    assert all(np.array_equal(both[i], zp.gt_one(cid)) for i in range(n))
    with pytest.raises(ValueError):
        ctx.multi_pairing(cid, aP, bQ, 3)                                      # 200 pairs are not groups of 3


@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_groth16_verify_proof_on_gpu(ctx, cid):
    """groth16/tests/mini.rs:46-97: generate -> prove -> prepare_verifying_key -> verify_proof, all on the GPU; the
    decisions agree with the oracle's verifier; then the batched form over good and bad proofs"""
    rng = random.Random(2025)
    circuit = MiniCircuit(2, 3, 10, 10)
    data = zgen.generate_random_parameters(ctx, cid, circuit, rng)
    params = data.upload(ctx)
    proofs = [zg.create_random_proof(params, circuit, rng) for _ in range(6)]
    params.free()
    pvk = zv.prepare_verifying_key(ctx, cid, data.vk)
    assert zv.verify_proof(pvk, proofs[0], [10])
    assert not zv.verify_proof(pvk, proofs[0], [11])
    with pytest.raises(zv.MalformedVerifyingKey):
        zv.verify_proof(pvk, proofs[0], [10, 1])
    with pytest.raises(zv.MalformedVerifyingKey):
        zv.verify_proof(pvk, proofs[0], [])
    # alpha_g1_beta_g2 is the pairing the oracle computes (to the device's fixed power)
    pt1 = lambda x: H.array_point(cid, 1, x[0], x[1])
    pt2 = lambda x: H.array_point(cid, 2, x[0], x[1])
    assert gt_to_flat(cid, pvk.alpha_g1_beta_g2) == oracle_gt(cid, [(pt1(data.vk.alpha_g1), pt2(data.vk.beta_g2))])
    # a batch: proof 1 with a wrong input, proof 3 with a and c swapped, proof 4 with b negated
    bad3 = zg.Proof(proofs[3].c, proofs[3].b, proofs[3].a)
    bad4 = zg.Proof(proofs[4].a, zp.neg_point(cid, _lib.G2, proofs[4].b), proofs[4].c)
    batch = [proofs[0], proofs[1], proofs[2], bad3, bad4, proofs[5]]
    inputs = [[10], [9], [10], [10], [10], [10 + FR[cid].p]]                   # the last one is 10 again mod r
    got = zv.verify_proofs(pvk, batch, inputs)
    assert got == [True, False, True, False, False, True]
    opvk = OP.prepare_verifying_key(cid, {"alpha_g1": pt1(data.vk.alpha_g1), "beta_g2": pt2(data.vk.beta_g2),
                                          "gamma_g2": pt2(data.vk.gamma_g2), "delta_g2": pt2(data.vk.delta_g2),
                                          "gamma_abc_g1": H.array_points(cid, 1, *data.vk.gamma_abc_g1)})
    for pr, x, g in zip(batch[:4], inputs[:4], got[:4]):
        assert OP.verify_proof(cid, opvk, (pt1(pr.a), pt2(pr.b), pt1(pr.c)), x) == g
    assert zv.verify_proofs(pvk, [], []) == []
    # one decision for a whole batch (random linear combination of the equations): B + 3 Miller loops, one final exponentiation
    brng = random.Random(99)
    good = [proofs[i % 6] for i in range(40)]
    assert zv.verify_proofs_batched(pvk, good, [[10]] * 40, brng)                      # short-MSM path (>= 32 proofs)
    assert zv.verify_proofs_batched(pvk, good[:5], [[10]] * 5, brng)                   # one call per r_i * A_i
    assert not zv.verify_proofs_batched(pvk, good[:17] + [bad4] + good[18:], [[10]] * 40, brng)
    assert not zv.verify_proofs_batched(pvk, good, [[10]] * 39 + [[12]], brng)
    assert zv.verify_proofs_batched(pvk, [], [], brng)
    with pytest.raises(zv.MalformedVerifyingKey):
        zv.verify_proofs_batched(pvk, good[:2], [[10], []], brng)
    pvk.free()


@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_kzg10_commit_open_check(ctx, cid):
    """the reference's unit test of KZG10 (kzg10.rs:235-270): random polynomials, commit with a hiding bound, open at a
    random point, `check` accepts; a wrong value, a wrong point and a proof for another polynomial are rejected"""
    p = FR[cid].p
    rng = random.Random(11)
    degree = 19
    pp = zm.universal_setup(ctx, cid, degree, random.Random(5))
    ck, vk = zm.pc_trim(ctx, pp, degree)
    prev = None
    for it in range(6):
        d = rng.randrange(2, degree + 1)
        coeffs = [rng.randrange(p) for _ in range(d)] + [rng.randrange(1, p)]
        poly = H.fr_array(cid, coeffs)
        hiding = 1 if it % 2 == 0 else None
        comm, rand = zk.kzg_commit(ck, poly, hiding, rng)
        z = rng.randrange(p)
        value = sum(c * pow(z, i, p) for i, c in enumerate(coeffs)) % p
        proof = zk.kzg_open(ck, poly, H.fr_array(cid, [z])[0], rand)
        assert (proof[1] is not None) == (hiding is not None)
        assert zk.kzg_check(ctx, vk, comm, z, value, proof)
        assert not zk.kzg_check(ctx, vk, comm, z, (value + 1) % p, proof)
        assert not zk.kzg_check(ctx, vk, comm, (z + 1) % p, value, proof)
        if prev is not None:
            assert not zk.kzg_check(ctx, vk, prev, z, value, proof)
        prev = comm
    ck.free()


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_long_products(ctx, cid):
    """a group much longer than a verifier's three pairs (one random-linear-combination check over a whole batch): the
    chunked product path.  prod e(a_i G1, b_i G2) == e((sum a_i b_i) G1, G2), and == 1 when the sum vanishes mod r"""
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    rng = random.Random(5)
    can = lambda v: H.ints_to_u64(v, 4)
    gen1, gen2 = arr1(cid, g1.gen)[0], arr2(cid, g2.gen)[0]
    for n in (33, 100, 700):
        a = [rng.randrange(1, g1.r) for _ in range(n)]
        b = [rng.randrange(1, g1.r) for _ in range(n)]
        total = sum(x * y for x, y in zip(a, b)) % g1.r
        aP = ctx.fixed_base_mul(cid, _lib.G1, gen1, can(a + [1]))
        bQ = ctx.fixed_base_mul(cid, _lib.G2, gen2, can(b + [(g1.r - total) % g1.r]))
        got = ctx.multi_pairing(cid, (aP[0][:n], None), (bQ[0][:n], None), n)
        want = ctx.multi_pairing(cid, ctx.fixed_base_mul(cid, _lib.G1, gen1, can([total])), (gen2.reshape(1, -1), None), 1)
        assert np.array_equal(got, want)
        closed = ctx.multi_pairing(cid, aP, bQ, n + 1)
        assert np.array_equal(closed[0], zp.gt_one(cid))
