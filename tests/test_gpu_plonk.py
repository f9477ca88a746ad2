"""The PLONK prover on the GPU (ckb_zkp_b200/plonk.py, restating plonk/src/{composer,ahp,lib.rs}) against the oracle
(oracle/pyref/plonk.py) and through the reference's own acceptance test for this layer (`fn ahp`, plonk/src/ahp/mod.rs:131-205:
index -> three prover rounds -> evaluations of the linear combinations -> verifier_equality_check)."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import marlin as zm
from ckb_zkp_b200 import plonk as zp
from oracle.pyref import plonk as OP
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR
from tests import helpers as H
from tests.test_oracle_plonk import KS, random_circuit

pytestmark = pytest.mark.gpu


def mirror(cs_oracle):
    """the same gates recorded by the product's Composer"""
    cs = zp.Composer(cs_oracle.p)
    cs.n, cs.pi = cs_oracle.n, list(cs_oracle.pi)
    cs.q = {k: list(v) for k, v in cs_oracle.q.items()}
    cs.w = [list(c) for c in cs_oracle.w]
    cs.variable_map = [list(w) for w in cs_oracle.variable_map]
    cs.assignment = list(cs_oracle.assignment)
    return cs


def product_test_circuit(p):
    """plonk/src/lib.rs:318-358 through the product's own gate API"""
    cs = zp.Composer(p)
    v1, v2, v3, v4, v6 = (cs.alloc_and_assign(x) for x in (1, 2, 3, 4, 6))
    cs.create_add_gate((v1, 1), (v2, 1), v3, None, 0, 0)
    cs.create_add_gate((v1, 1), (v3, 1), v4, None, 0, 0)
    cs.create_mul_gate(v2, v2, v4, None, 1, 0, 0)
    cs.create_mul_gate(v1, v2, v6, None, 2, 2, 0)
    cs.constrain_to_constant(v6, 6, 0)
    return cs


def ints(cid, arr):
    return H.fr_ints(cid, arr) if len(arr) else []


@pytest.mark.parametrize("cid,which", [(BLS12_381, "test"), (BN254, "test"), (BLS12_381, "random"), (BN254, "random")])
def test_index_and_rounds_match_oracle_and_pass_the_equality_check(ctx, cid, which):
    fr = FR[cid]
    p = fr.p
    ocs = OP.test_circuit(p) if which == "test" else random_circuit(p, 70, 3)
    cs = product_test_circuit(p) if which == "test" else mirror(ocs)
    rng = random.Random(cid * 7 + len(which))
    beta, gamma, alpha, zeta = (rng.randrange(p) for _ in range(4))

    oidx = OP.index(ocs, fr, KS)
    idx = zp.index(ctx, cid, cs, KS)
    assert idx.n == oidx.n
    for label in OP.SELECTOR_LABELS:
        assert ints(cid, idx.polys[label]) == oidx.polys[label], label
        assert ints(cid, idx.evals_4n[label]) == oidx.evals_4n[label], label
    assert ints(cid, idx.v_4n_inversed) == oidx.v_4n_inversed
    assert ints(cid, idx.l1_4n) == oidx.l1_4n

    ops = OP.prover_init(ocs, oidx)
    ps = zp.prover_init(ctx, cs, idx)
    assert ints(cid, ps.pi_4n) == ops.pi_4n
    polys, opolys = dict(idx.polys), dict(oidx.polys)
    for got, want in ((zp.prover_first_round(ps, cs), OP.prover_first_round(ops, ocs)),
                      (zp.prover_second_round(ps, beta, gamma), OP.prover_second_round(ops, beta, gamma)),
                      (zp.prover_third_round(ps, alpha), OP.prover_third_round(ops, alpha)[0])):
        assert sorted(got) == sorted(want)
        for label in want:
            assert ints(cid, got[label]) == want[label], label
        polys.update(got)
        opolys.update(want)

    # the reference's `fn ahp`: evaluations of the linear combinations at the query set, then the equality check --
    # evaluated by the product on the GPU, judged by the oracle's restatement of the verifier (and by the product's)
    lcs = zp.construct_linear_combinations(ctx, idx, beta, gamma, alpha, zeta, polys)
    olcs = OP.linear_combinations(oidx, beta, gamma, alpha, zeta, opolys)
    assert lcs == olcs
    qs = zp.verifier_query_set(idx, zeta)
    evals = {label: zp._eval(ctx, cid, zp.lc_polynomial(ctx, cid, lcs[label], polys), point) for label, (_, point) in qs.items()}
    assert evals == {label: OP.lc_eval(olcs[label], opolys, point, p) for label, point in OP.query_set(oidx, zeta).items()}
    assert OP.verifier_equality_check(oidx, beta, gamma, alpha, zeta, evals, ocs.public_inputs())
    assert zp.verifier_equality_check(ctx, idx, beta, gamma, alpha, zeta, evals, cs.public_inputs())
    bad = dict(evals)
    bad["t"] = (bad["t"] + 1) % p
    assert not zp.verifier_equality_check(ctx, idx, beta, gamma, alpha, zeta, bad, cs.public_inputs())


def test_unsatisfied_copy_constraint_trips_the_closing_assertion(ctx):
    """indexer/permutation.rs:118 `assert_eq!(z[n - 1] * perms[n - 1], F::one())`"""
    cid = BN254
    p = FR[cid].p
    cs = product_test_circuit(p)
    idx = zp.index(ctx, cid, cs, KS)
    cs.w[1][0] = 2                                # gate 0 now reads var_two on wire l without the permutation knowing
    ps = zp.prover_init(ctx, cs, idx)
    zp.prover_first_round(ps, cs)
    with pytest.raises(AssertionError):
        zp.prover_second_round(ps, 5, 6)


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_keygen_prove_self_verifies(ctx, cid, resident):
    """Plonk::{setup, keygen, prove} (plonk/src/lib.rs:53-204, test_plonk :361-376) with the Fiat-Shamir generator in the
    loop: commitments equal kg * p(beta) * G for the known trapdoor, the evaluations pass the equality check under the
    challenges the transcript produced, and each opening satisfies the KZG equation W * (beta - z) == P(beta) - v in the
    exponent (marlin/src/pc/kzg10.rs:158-172)."""
    fr = FR[cid]
    p = fr.p
    ocs = random_circuit(p, 45, 11)
    cs = mirror(ocs)
    setup_rng = random.Random(5)
    beta_t, kg = setup_rng.randrange(1, p), setup_rng.randrange(1, p)          # the first two draws of universal_setup
    srs = zm.universal_setup(ctx, cid, 64, random.Random(5))
    pk, vk = zp.keygen(ctx, srs, cs, KS, resident=resident)
    proof, ch = zp.prove(ctx, pk, cs)
    idx = pk.index
    c1 = CURVES[(cid, 1)]
    commit_of = lambda poly: c1.mul_affine(c1.gen, kg * OP.poly_eval(poly, beta_t, p) % p) if any(poly) else None

    # oracle run under the transcript's challenges: same polynomials, hence same commitments
    oidx = OP.index(ocs, fr, KS)
    ops = OP.prover_init(ocs, oidx)
    opolys = dict(oidx.polys)
    opolys.update(OP.prover_first_round(ops, ocs))
    opolys.update(OP.prover_second_round(ops, ch["beta"], ch["gamma"]))
    opolys.update(OP.prover_third_round(ops, ch["alpha"])[0])
    for (comm, shifted), label in zip(vk["comms"], OP.SELECTOR_LABELS):
        assert shifted is None and H.array_point(cid, 1, comm[0], comm[1]) == commit_of(opolys[label]), label
    flat = [c for rnd in proof.commitments for c in rnd]
    assert [len(r) for r in proof.commitments] == [4, 1, 4]
    for (comm, shifted), label in zip(flat, zp.ORACLE_LABELS):
        assert shifted is None and H.array_point(cid, 1, comm[0], comm[1]) == commit_of(opolys[label]), label

    evals = ch["evals"]
    assert proof.evaluations == [evals[l] for l in sorted(evals)] and len(proof.evaluations) == 11
    assert OP.verifier_equality_check(oidx, ch["beta"], ch["gamma"], ch["alpha"], ch["zeta"], evals, ocs.public_inputs())

    # openings: W = kg * (P(beta) - v) / (beta - z) * G with P = sum_j eps^(2 j) * lc_j over the labels of the point
    olcs = OP.linear_combinations(oidx, ch["beta"], ch["gamma"], ch["alpha"], ch["zeta"], opolys)
    qs = OP.query_set(oidx, ch["zeta"])
    for point_label, labels in (("shifted_zeta", ["z"]), ("zeta", sorted(l for l in qs if l != "z"))):
        z = qs[labels[0]]
        P_beta, v, c = 0, 0, 1
        for label in labels:
            P_beta = (P_beta + c * OP.lc_eval(olcs[label], opolys, beta_t, p)) % p
            v = (v + c * evals[label]) % p
            c = c * ch["epsilon"] % p * ch["epsilon"] % p
        want = c1.mul_affine(c1.gen, kg * (P_beta - v) % p * pow(beta_t - z, -1, p) % p)
        w, rand_v = proof.openings[point_label]
        assert rand_v is None and H.array_point(cid, 1, w[0], w[1]) == want

    # Plonk::verify (lib.rs:206-290) on the device: transcript re-derived, equality check, both openings by pairings
    pis = cs.public_inputs()
    assert zp.verify(ctx, vk, pis, proof)
    assert not zp.verify(ctx, vk, [(pis[0] + 1) % p] + list(pis[1:]), proof)              # another transcript, check fails
    swapped = zp.Proof(proof.commitments, proof.evaluations, {"zeta": proof.openings["shifted_zeta"],
                                                              "shifted_zeta": proof.openings["zeta"]})
    assert not zp.verify(ctx, vk, pis, swapped)                                         # equality check passes, pairings fail
    pk.ck.free()
