"""The Marlin oracle against the reference's OWN acceptance test, on the CPU: marlin/tests/mini.rs:43-88 sets up, indexes,
proves the `Mini` circuit and asserts `verify_proof(&vk, &proof, &[10]) == true` (marlin/src/lib.rs:184-260: Fiat-Shamir
replay, AHP verifier_equality_check, KZG10 pairing checks).  oracle/pyref/marlin_proof.py restates prover and verifier;
here the restated prover's proof must be accepted by the restated verifier (real pairings, oracle/pyref/pairing.py) and
every corruption rejected -- this is what pins the Marlin half of the oracle the GPU tests compare against."""
import random

import pytest

from oracle.pyref import marlin as OM
from oracle.pyref import marlin_proof as MP
from oracle.pyref.fields import BLS12_381, BN254, FR

ONE = ("in", 0)


def mini(cs, num=10):
    vx, vy = cs.alloc(2), cs.alloc(3)
    vz = cs.alloc_input(10)
    for _ in range(num):
        cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_marlin_mini_proof_is_accepted_and_corruptions_rejected(cid):
    p = FR[cid].p
    cs = OM.MarlinCS(p)
    mini(cs)
    pp = MP.universal_setup(cid, 128, beta=0x1234567, kg=3, kgamma=5, kh=7)
    ipk, ivk = MP.index(pp, cs)
    assert ivk["index_info"] == (10, 10, 20) and ivk["verifier_key"]["supported_degree"] == 93
    proof = MP.create_random_proof(ipk, cs, random.Random(1))
    assert len(proof["evaluations"]) == 21 and len(proof["opening_proofs"]) == 2
    assert [len(r) for r in proof["commitments"]] == [4, 3, 2]
    assert MP.verify_proof(ivk, proof, [10])
    assert not MP.verify_proof(ivk, proof, [11])                         # other public input: other transcript, AHP fails
    bad = dict(proof, evaluations=list(proof["evaluations"]))
    bad["evaluations"][3] = (bad["evaluations"][3] + 1) % p
    assert not MP.verify_proof(ivk, bad, [10])
    # a wrong opening witness with everything else intact: only the pairing check can catch it
    w, rv = proof["opening_proofs"][0]
    from oracle.pyref.curves import CURVES
    g1 = CURVES[(cid, 1)]
    bad = dict(proof, opening_proofs=[(g1.mul_affine(w, 2), rv), proof["opening_proofs"][1]])
    assert not MP.verify_proof(ivk, bad, [10])
    # a different proof of the same statement (other prover randomness) is accepted too
    assert MP.verify_proof(ivk, MP.create_random_proof(ipk, cs, random.Random(2)), [10])
