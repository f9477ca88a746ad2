"""CPU test of the verifier host layers (ckb_zkp_b200/{pairing,verifier}.py, kzg10.kzg_check / pc_check) over a mock
context: the pairings run the DEVICE code compiled for the host (tests/host_emu, csrc/pairing.cuh), the small MSMs run the
oracle.  Oracle-made Groth16 proofs (mini circuit, groth16/tests/mini.rs) and KZG10 openings (kzg10.rs:235-270) must be
accepted, tampered ones rejected, with the reference's error behaviour."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import _lib
from ckb_zkp_b200 import kzg10 as zk
from ckb_zkp_b200 import pairing as zp
from ckb_zkp_b200 import verifier as zv
from ckb_zkp_b200.generator import VerifyKey
from ckb_zkp_b200.groth16 import Proof
from ckb_zkp_b200.marlin import VerifierKey
from oracle.pyref import groth16 as OG
from oracle.pyref import kzg10 as OK
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FQ, FR, stream_field
from oracle.pyref.msm import msm_naive
from oracle.pyref.r1cs import ConstraintSystem, mini_circuit
from tests import emu
from tests import helpers as H


class _Srs:
    def __init__(self, curve, group, pts):
        self.curve, self.group, self.pts = curve, group, pts

    def free(self):
        pass


class MockVerifierContext:
    """the three calls the verifier layers make"""

    def __init__(self):
        self.lib = emu.build()

    def srs_upload(self, curve, group, xy, inf=None, precompute=True):
        inf = np.zeros(len(xy), dtype=np.uint8) if inf is None else inf
        return _Srs(curve, group, H.array_points(curve, group, np.asarray(xy), np.asarray(inf)))

    def msm_batch(self, srs_list, scalars_list, base_offsets=None, mont=False):
        assert not mont
        out = []
        for srs, sc in zip(srs_list, scalars_list):
            c = CURVES[(srs.curve, srs.group)]
            pt = c.to_affine(msm_naive(c, srs.pts, H.u64_to_ints(np.asarray(sc))))
            xy, inf = H.points_array(srs.curve, srs.group, [pt])
            out.append((xy[0], bool(inf[0])))
        return out

    def msm_many(self, srs, scalars, base_offsets=None, mont=False):
        k = len(scalars)
        offs = [0] * k if base_offsets is None else [int(o) for o in base_offsets]
        res = [self.msm(srs, scalars[i], base_offset=offs[i]) for i in range(k)]
        return np.stack([r[0] for r in res]), np.array([1 if r[1] else 0 for r in res], dtype=np.uint8)

    def msm(self, srs, scalars, base_offset=0, mont=False):
        assert not mont
        c = CURVES[(srs.curve, srs.group)]
        pt = c.to_affine(msm_naive(c, srs.pts[base_offset:], H.u64_to_ints(np.asarray(scalars))))
        xy, inf = H.points_array(srs.curve, srs.group, [pt])
        return xy[0], bool(inf[0])

    def fixed_base_mul(self, curve, group, base_xy, scalars):
        c = CURVES[(curve, group)]
        base = H.array_point(curve, group, np.asarray(base_xy), 0)
        return H.points_array(curve, group, [c.mul_affine(base, k) if k else None for k in H.u64_to_ints(np.asarray(scalars))])

    def multi_pairing(self, curve, g1, g2, group_size):
        """csrc/pairing.cuh on the host: Miller loops, product in the oracle's field, final exponentiation"""
        from oracle.pyref import pairing as OP
        from tests.test_pairing_emu import flat_to_tower, run, tower_to_flat
        (xy1, inf1), (xy2, inf2) = g1, g2
        n = len(xy1)
        inf1 = np.zeros(n, dtype=np.uint8) if inf1 is None else inf1
        inf2 = np.zeros(n, dtype=np.uint8) if inf2 is None else inf2
        L, q = FQ[curve].limbs, FQ[curve].p
        F12 = OP.Fq12(curve)
        out = np.zeros((n // group_size, 12 * L), dtype=np.uint64)
        for g in range(n // group_size):
            f = F12.one
            for i in range(g * group_size, (g + 1) * group_size):
                P = H.array_point(curve, 1, xy1[i], inf1[i])
                Q = H.array_point(curve, 2, xy2[i], inf2[i])
                f = F12.mul(f, tower_to_flat(curve, run(self.lib, curve, 0, P, Q)))
            gt = run(self.lib, curve, 1, None, None, flat_to_tower(curve, f))
            out[g] = H.ints_to_u64([v * (1 << (64 * L)) % q for v in gt], L).reshape(-1)
        return out


def arr(cid, group, P):
    xy, inf = H.points_array(cid, group, [P])
    return xy[0], bool(inf[0])


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_groth16_verifier_over_mock(cid):
    ctx = MockVerifierContext()
    p = FR[cid].p
    cs = mini_circuit(ConstraintSystem(p))
    alpha, beta, gamma, delta, t = [stream_field(3, i, p) for i in range(5)]
    pk = OG.generate_parameters(cs, cid, alpha, beta, gamma, delta, t)
    a, b, c = OG.create_proof(pk, cs, 12345, 67890)
    assert OG.verify_proof(pk, (a, b, c), [10])
    vk = VerifyKey(arr(cid, 1, pk.alpha_g1), arr(cid, 2, pk.beta_g2), arr(cid, 2, pk.gamma_g2), arr(cid, 2, pk.delta_g2),
                   H.points_array(cid, 1, pk.gamma_abc_g1))
    pvk = zv.prepare_verifying_key(ctx, cid, vk)
    proof = Proof(arr(cid, 1, a), arr(cid, 2, b), arr(cid, 1, c))
    bad = Proof(arr(cid, 1, c), arr(cid, 2, b), arr(cid, 1, a))
    assert zv.verify_proofs(pvk, [proof, proof, bad], [[10], [11], [10]]) == [True, False, False]
    assert zv.verify_proof(pvk, proof, [10 + p])
    # the random-linear-combination form: one decision for the batch
    a2, b2, c2 = OG.create_proof(pk, cs, 777, 888)
    proof2 = Proof(arr(cid, 1, a2), arr(cid, 2, b2), arr(cid, 1, c2))
    brng = random.Random(1)
    assert zv.verify_proofs_batched(pvk, [proof, proof2, proof], [[10], [10], [10]], brng)
    assert not zv.verify_proofs_batched(pvk, [proof, bad, proof2], [[10], [10], [10]], brng)
    assert not zv.verify_proofs_batched(pvk, [proof, proof2], [[10], [11]], brng)
    assert zv.verify_proofs_batched(pvk, [], [], brng)
    with pytest.raises(zv.MalformedVerifyingKey):
        zv.verify_proof(pvk, proof, [])
    with pytest.raises(ValueError):
        zv.verify_proofs(pvk, [proof], [])
    # -P of the identity stays the identity, and double negation is the identity map
    assert zp.neg_point(cid, 2, (vk.beta_g2[0], True))[1] is True
    back = zp.neg_point(cid, 2, zp.neg_point(cid, 2, vk.beta_g2))
    assert np.array_equal(back[0], vk.beta_g2[0])
    assert H.array_point(cid, 2, *zp.neg_point(cid, 2, vk.beta_g2)) == CURVES[(cid, 2)].neg_affine(pk.beta_g2)


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_kzg_check_over_mock(cid):
    ctx = MockVerifierContext()
    p = FR[cid].p
    g2 = CURVES[(cid, 2)]
    rng = random.Random(4)
    beta, kh = rng.randrange(1, p), rng.randrange(1, p)
    pp = OK.setup(cid, 8, beta, g_scalar=rng.randrange(1, p), gamma=rng.randrange(1, p))
    h = g2.mul_affine(g2.gen, kh)
    vk = VerifierKey(cid, arr(cid, 1, pp["powers_of_g"][0]), arr(cid, 1, pp["powers_of_gamma_g"][0]), arr(cid, 2, h),
                     arr(cid, 2, g2.mul_affine(h, beta)), 8)
    for hiding in (False, True):
        poly = [rng.randrange(p) for _ in range(7)]
        blind = [rng.randrange(p) for _ in range(2)] if hiding else None
        comm = OK.kzg_commit(cid, pp["powers_of_g"], pp["powers_of_gamma_g"], poly, blind)
        z = rng.randrange(p)
        value = OK.poly_eval(poly, z, p)
        w, rand_v = OK.kzg_open(cid, pp["powers_of_g"], pp["powers_of_gamma_g"], poly, z, blind)
        proof = (arr(cid, 1, w), None if rand_v is None else H.fr_array(cid, [rand_v])[0])
        assert (rand_v is not None) == hiding
        assert zk.kzg_check(ctx, vk, arr(cid, 1, comm), z, value, proof)
        assert not zk.kzg_check(ctx, vk, arr(cid, 1, comm), z, (value + 1) % p, proof)
        # PC::check over one commitment without a degree bound is KZG10::check on (comm, value)
        assert zk.pc_check(ctx, vk, [(arr(cid, 1, comm), None)], [None], z, [value], proof, 0x1234567)
        with pytest.raises(zk.MissingEvaluation):
            zk.pc_batch_check(ctx, vk, {"p": ((arr(cid, 1, comm), None), None)}, [("p", z)], {}, [proof], 5)
        with pytest.raises(zk.MissingPolynomial):
            zk.pc_batch_check(ctx, vk, {}, [("p", z)], {("p", z): value}, [proof], 5)
        assert zk.pc_batch_check(ctx, vk, {"p": ((arr(cid, 1, comm), None), None)}, [("p", z)], {("p", z): value}, [proof], 5)
