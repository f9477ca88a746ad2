"""Whole proofs on the CPU: the host layers of ckb_zkp_b200 (marlin.py, kzg10.py, plonk.py, verifier-side code) driven end to
end over tests/mock_backend.MockProverVerifierContext -- field / polynomial / group primitives by the oracle, pairings by
the device code compiled for the host.  What runs is exactly the orchestration the GPU runs (which primitive, which
operands, which transcript bytes); the kernels themselves are checked in the -m gpu tests."""
import random

import pytest

from ckb_zkp_b200 import marlin as zm
from oracle.pyref import marlin as OM
from oracle.pyref import marlin_proof as MP
from oracle.pyref.fields import BLS12_381, BN254, FR
from tests import helpers as H
from tests.mock_backend import MockProverVerifierContext
from tests.test_gpu_marlin_proof import Mini


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_marlin_setup_index_prove_verify_over_mock(cid):
    """zkp_marlin's crate-level flow (marlin/tests/mini.rs:46-90): universal_setup -> index -> create_random_proof ->
    verify_proof, the proof equal to the oracle prover's and accepted by both verifiers"""
    ctx = MockProverVerifierContext()
    p = FR[cid].p
    circuit = Mini()
    cs = OM.MarlinCS(p)
    circuit.generate_constraints(cs)
    oidx = OM.index(cs, cid)
    need = MP.max_degree(oidx["num_constraints"], oidx["num_variables"], oidx["num_non_zeros"])
    srs = zm.universal_setup(ctx, cid, need, random.Random(77))
    ipk, ivk = zm.index_keys(ctx, srs, circuit)
    proof = zm.create_random_proof(ctx, ipk, circuit, random.Random(4242), resident=False)
    assert len(proof.evaluations) == 21 and len(proof.opening_proofs) == 2
    # the oracle prover with the same setup draws and prover randomness
    setup_rng = random.Random(77)
    beta, kg, kgamma, kh = (setup_rng.randrange(1, p) for _ in range(4))
    opp = MP.universal_setup(cid, srs.max_degree(), beta, kg, kgamma, kh)
    oipk, oivk = MP.index(opp, cs)
    want = MP.create_random_proof(oipk, cs, random.Random(4242))
    assert proof.challenges == want["challenges"]
    assert H.fr_ints(cid, __import__("numpy").stack(proof.evaluations)) == want["evaluations"]
    point = lambda pt: H.array_point(cid, 1, pt[0], pt[1])
    for got_round, want_round in zip(proof.commitments, want["commitments"]):
        for (gc, gs), (wc, ws) in zip(got_round, want_round):
            assert point(gc) == wc and ((gs is None) == (ws is None)) and (gs is None or point(gs) == ws)
    # both verifiers
    assert zm.verify_proof(ctx, ivk, proof, cs.input[1:])
    assert not zm.verify_proof(ctx, ivk, proof, [(v + 1) % p for v in cs.input[1:]])
    forged = zm.Proof(proof.commitments, proof.evaluations, [proof.opening_proofs[1], proof.opening_proofs[0]])
    assert not zm.verify_proof(ctx, ivk, forged, cs.input[1:])
    gpu_proof = {"commitments": [[(point(c), None if s is None else point(s)) for c, s in rnd] for rnd in proof.commitments],
                 "evaluations": want["evaluations"],
                 "opening_proofs": [(point(w), None if rv is None else H.fr_ints(cid, rv.reshape(1, 4))[0]) for w, rv in proof.opening_proofs]}
    assert MP.verify_proof(oivk, gpu_proof, cs.input[1:])


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_plonk_keygen_prove_verify_over_mock(cid):
    """Plonk::{keygen, prove, verify} (plonk/src/lib.rs:62-290, the shape of test_plonk :361-376) on the CPU: the proof
    verifies, a changed public input and swapped openings do not; the evaluations satisfy the oracle's equality check
    under the challenges the transcript produced"""
    from ckb_zkp_b200 import plonk as zp
    from oracle.pyref import plonk as OP
    from tests.test_oracle_plonk import KS, random_circuit
    from tests.test_plonk_host import mirror
    ctx = MockProverVerifierContext()
    fr = FR[cid]
    p = fr.p
    ocs = random_circuit(p, 13, 3)
    cs = mirror(ocs)
    srs = zm.universal_setup(ctx, cid, 32, random.Random(5))
    pk, vk = zp.keygen(ctx, srs, cs, KS)
    proof, ch = zp.prove(ctx, pk, cs)
    assert [len(r) for r in proof.commitments] == [4, 1, 4] and len(proof.evaluations) == 11
    oidx = OP.index(ocs, fr, KS)
    assert OP.verifier_equality_check(oidx, ch["beta"], ch["gamma"], ch["alpha"], ch["zeta"], ch["evals"], ocs.public_inputs())
    pis = cs.public_inputs()
    assert zp.verify(ctx, vk, pis, proof)
    assert not zp.verify(ctx, vk, [(pis[0] + 1) % p] + list(pis[1:]), proof)
    swapped = zp.Proof(proof.commitments, proof.evaluations, {"zeta": proof.openings["shifted_zeta"],
                                                              "shifted_zeta": proof.openings["zeta"]})
    assert not zp.verify(ctx, vk, pis, swapped)
    tampered = zp.Proof(proof.commitments, [(proof.evaluations[0] + 1) % p] + proof.evaluations[1:], proof.openings)
    assert not zp.verify(ctx, vk, pis, tampered)


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_groth16_generator_over_mock(cid):
    """generate_parameters (groth16/src/generator.rs:135-286) host logic on the CPU against the oracle's generator on the
    same toxic waste: every query and the verifying key"""
    import numpy as np
    from ckb_zkp_b200 import generator as zgen
    from oracle.pyref import groth16 as OG
    from oracle.pyref.fields import stream_field
    from oracle.pyref.r1cs import ConstraintSystem, mini_circuit
    from tests.test_gpu_generator import MiniCircuit, same_points
    ctx = MockProverVerifierContext()
    p = FR[cid].p
    alpha, beta, gamma, delta, t = [stream_field(7, i, p) for i in range(5)]
    want = OG.generate_parameters(mini_circuit(ConstraintSystem(p)), cid, alpha, beta, gamma, delta, t)
    got = zgen.generate_parameters(ctx, cid, MiniCircuit(2, 3, 10, 10), alpha, beta, gamma, delta, t)
    same_points(cid, 1, got.a_query, want.a_query)
    same_points(cid, 1, got.b_g1_query, want.b_g1_query)
    same_points(cid, 2, got.b_g2_query, want.b_g2_query)
    same_points(cid, 1, got.h_query, want.h_query)
    same_points(cid, 1, got.l_query, want.l_query)
    same_points(cid, 1, got.vk.gamma_abc_g1, want.gamma_abc_g1)
    one = lambda pt: ([pt[0]], [1 if pt[1] else 0])
    same_points(cid, 1, one(got.vk.alpha_g1), [want.alpha_g1])
    same_points(cid, 2, one(got.vk.delta_g2), [want.delta_g2])
