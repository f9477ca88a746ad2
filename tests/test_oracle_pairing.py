"""The oracle against the reference's OWN acceptance test: groth16/tests/mini.rs:46-97 proves the `Mini` circuit and
asserts `verify_proof(&pvk, &proof, &[10]) == true` (groth16/src/verifier.rs:18-44).  The pairing is restated in
oracle/pyref/pairing.py; here it is pinned by its defining properties (bilinear, non-degenerate, order r) on both
curves, then the oracle's proofs -- and with them the golden fixtures the GPU tests compare against -- are put through
the restated verifier: accepted as minted, rejected after any change.  CPU only."""
import os
import random

import numpy as np
import pytest

from oracle.pyref import groth16 as OG
from oracle.pyref import pairing as PR
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR, stream_field
from oracle.pyref.r1cs import ConstraintSystem, mimc_circuit, mini_circuit
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_pairing_is_bilinear_and_non_degenerate(cid):
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    r = FR[cid].p
    F12 = PR.Fq12(cid)
    e = PR.pairing(cid, g1.gen, g2.gen)
    assert e != F12.one and F12.pow(e, r) == F12.one
    rng = random.Random(cid)
    a, b = rng.randrange(r), rng.randrange(r)
    assert PR.pairing(cid, g1.mul_affine(g1.gen, a), g2.mul_affine(g2.gen, b)) == F12.pow(e, a * b % r)
    # e(aP, Q) * e(-P, aQ) == 1 through the shared final exponentiation, identities give 1
    assert PR.multi_pairing(cid, [(g1.mul_affine(g1.gen, a), g2.gen),
                                  (g1.neg_affine(g1.gen), g2.mul_affine(g2.gen, a))]) == F12.one
    assert PR.pairing(cid, None, g2.gen) == F12.one and PR.pairing(cid, g1.gen, None) == F12.one


def _mini(cid, seed=1):
    fr = FR[cid]
    cs = mini_circuit(ConstraintSystem(fr.p))
    alpha, beta, gamma, delta, t, r, s = [stream_field(seed, i, fr.p) for i in range(7)]
    pk = OG.generate_parameters(cs, cid, alpha, beta, gamma, delta, t)
    return cs, pk, r, s


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_oracle_proof_passes_the_reference_acceptance_test(cid):
    """mini.rs:85-89 with x = 2, y = 3, z = 10: accept; then every way of being wrong: reject"""
    cs, pk, r, s = _mini(cid)
    proof = OG.create_proof(pk, cs, r, s)
    assert OG.verify_proof(pk, proof, [10])
    assert OG.verify_proof(pk, OG.create_proof(pk, cs, 0, 0), [10])          # create_proof_no_zk (prover.rs:113-122)
    assert not OG.verify_proof(pk, proof, [11])                                # wrong public input
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    a, b, c = proof
    assert not OG.verify_proof(pk, (g1.mul_affine(a, 2), b, c), [10])
    assert not OG.verify_proof(pk, (a, g2.mul_affine(b, 3), c), [10])
    assert not OG.verify_proof(pk, (a, b, g1.neg_affine(c)), [10])
    with pytest.raises(PR.MalformedVerifyingKey):                              # verifier.rs:23-25
        OG.verify_proof(pk, proof, [10, 1])
    # an unsatisfying witness yields a proof the verifier rejects
    bad = mini_circuit(ConstraintSystem(FR[cid].p))
    bad.aux_assignment[0] = (bad.aux_assignment[0] + 1) % FR[cid].p
    assert not bad.is_satisfied()
    assert not OG.verify_proof(pk, OG.create_proof(pk, bad, r, s), [10])


@pytest.mark.parametrize("name,cid", [("groth16_mini_bls12_381", BLS12_381), ("groth16_mini_bn254", BN254)])
def test_golden_mini_proofs_are_accepted(name, cid):
    """the committed fixtures (what the GPU proofs are compared with byte for byte) hold proofs the reference's
    verifier accepts -- the proof points are read back from the fixture, not recomputed"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cs, pk, r, s = _mini(cid)
    assert H.u64_to_int(g["r"][0]) == r and H.u64_to_int(g["s"][0]) == s
    proof = tuple(H.array_point(cid, grp, g[k + "_xy"][0], bool(g[k + "_inf"][0]))
                  for k, grp in (("proof_a", 1), ("proof_b", 2), ("proof_c", 1)))
    assert OG.verify_proof(pk, proof, [10])


def test_golden_mimc_proof_is_accepted():
    """MiMC chain, 64 constraints, BLS12-381 (the smoke() instance): public input = the image"""
    cid = BLS12_381
    fr = FR[cid]
    cs = mimc_circuit(ConstraintSystem(fr.p), 64)
    alpha, beta, gamma, delta, t, r, s = [stream_field(2, i, fr.p) for i in range(7)]
    g = np.load(os.path.join(GOLD, "groth16_mimc_bls12_381_2e6.npz"))
    if H.u64_to_int(g["r"][0]) != r:
        pytest.fail("fixture minted from another toxic-waste stream than make_golden.py states (seed 2)")
    pk = OG.generate_parameters(cs, cid, alpha, beta, gamma, delta, t)
    proof = tuple(H.array_point(cid, grp, g[k + "_xy"][0], bool(g[k + "_inf"][0]))
                  for k, grp in (("proof_a", 1), ("proof_b", 2), ("proof_c", 1)))
    assert OG.verify_proof(pk, proof, cs.input_assignment[1:])


def test_optimal_ate_on_bn254_is_a_pairing():
    """miller_loop_optimal_bn (what csrc/pairing.cuh runs on BN254): the Frobenius of a twist point is [q]Q on G2, the
    result is bilinear and non-degenerate, and it decides pairing-product equalities like the plain ate pairing"""
    from oracle.pyref import pairing as OP
    from oracle.pyref.curves import CURVES
    from oracle.pyref.fields import FQ
    cid = BN254
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    Q = g2.mul_affine(g2.gen, 987654321)
    assert OP._frobenius_twist(cid, Q) == g2.mul_affine(Q, FQ[cid].p % g2.r)
    a, b = 0x1F2E3D4C5B6A, 0x123456789ABCDEF
    one = OP.Fq12(cid).one
    e_ab = OP.device_multi_pairing(cid, [(g1.mul_affine(g1.gen, a), g2.mul_affine(g2.gen, b))])
    assert e_ab == OP.device_multi_pairing(cid, [(g1.mul_affine(g1.gen, a * b % g1.r), g2.gen)]) != one
    assert e_ab == OP.device_multi_pairing(cid, [(g1.gen, g2.mul_affine(g2.gen, a * b % g1.r))])
    # e(aP, Q) e(-P, aQ) == 1 under both pairings; e(aP, Q) e(-P, (a + 1) Q) != 1 under both
    good = [(g1.mul_affine(g1.gen, a), Q), (g1.neg_affine(g1.gen), g2.mul_affine(Q, a))]
    bad = [(g1.mul_affine(g1.gen, a), Q), (g1.neg_affine(g1.gen), g2.mul_affine(Q, a + 1))]
    assert OP.device_multi_pairing(cid, good) == one and OP.multi_pairing(cid, good) == one
    assert OP.device_multi_pairing(cid, bad) != one and OP.multi_pairing(cid, bad) != one
