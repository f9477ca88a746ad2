"""Groth16 prove path on the GPU vs the oracle (restating groth16/src/prover.rs:124-228 and
groth16/src/r1cs_to_qap.rs:113-172) on the committed golden instances, including BASELINE
config 1 (BN256, 2^10-constraint MiMC chain), and through the reference-shaped host API."""
import os
import random

import numpy as np
import pytest

from ckb_zkp_b200 import groth16 as zg
from ckb_zkp_b200.backend import CsrMatrix
from ckb_zkp_b200.r1cs import ONE
from oracle.pyref import groth16 as OG
from oracle.pyref.fields import BLS12_381, BN254, FR, stream_field
from oracle.pyref.r1cs import ConstraintSystem, mimc_circuit, mini_circuit
from tests import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["groth16_mini_bls12_381", "groth16_mini_bn254", "groth16_mimc_bls12_381_2e6", "groth16_mimc_bn254_2e10"]


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def params_from_golden(ctx, g):
    cid = int(g["curve"])
    q = lambda k: (g[k + "_xy"], g[k + "_inf"])
    s1, s2 = g["g1_singles"], g["g2_singles"]
    return zg.Parameters(ctx, cid, q("a_query"), q("b_g1_query"), q("b_g2_query"), q("h_query"), q("l_query"), s1[0],
                         s1[1], s1[2], s2[0], s2[1])


def matrices(g):
    return [CsrMatrix(g[w + "_ptr"], g[w + "_col"], g[w + "_val"]) for w in "abc"]


def assert_proof(g, proof):
    for key, got in (("proof_a", proof[0]), ("proof_b", proof[1]), ("proof_c", proof[2])):
        assert bool(g[key + "_inf"][0]) == got[1], key
        if not got[1]:
            assert np.array_equal(g[key + "_xy"][0], got[0]), key


@pytest.mark.parametrize("name", CASES)
def test_witness_map_matches_golden(ctx, name):
    g = load(name)
    A, B, C = matrices(g)
    h = ctx.groth16_h(int(g["curve"]), A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]))
    assert np.array_equal(h, g["h"])


@pytest.mark.parametrize("name", CASES)
def test_prove_matches_golden(ctx, name):
    g = load(name)
    params = params_from_golden(ctx, g)
    A, B, C = matrices(g)
    proof = ctx.groth16_prove(params.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]), g["r"][0], g["s"][0])
    assert_proof(g, proof)
    # staged variant (inputs resident in HBM) gives the same bytes, twice in a row
    ctx.groth16_stage(params.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]))
    for _ in range(2):
        ctx.groth16_prove_staged(params.pk, g["r"][0], g["s"][0])
        assert_proof(g, ctx.groth16_fetch_proof(params.pk))
    params.free()


class MiniCircuit:
    """groth16/tests/mini.rs:12-44"""

    def __init__(self, x, y, z, num):
        self.x, self.y, self.z, self.num = x, y, z, num

    def generate_constraints(self, cs):
        vx = cs.alloc(lambda: self.x)
        vy = cs.alloc(lambda: self.y)
        vz = cs.alloc_input(lambda: self.z)
        for _ in range(self.num):
            cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])


def test_reference_api_mini(ctx):
    """mirrors groth16/tests/mini.rs:46-97: prove Mini through create_proof / create_random_proof /
    create_proof_no_zk and compare with the oracle's prover on the same parameters."""
    g = load("groth16_mini_bls12_381")
    params = params_from_golden(ctx, g)
    circuit = MiniCircuit(2, 3, 10, 10)
    r, s = H.u64_to_int(g["r"][0]), H.u64_to_int(g["s"][0])
    proof = zg.create_proof(params, circuit, r, s)
    assert_proof(g, (proof.a, proof.b, proof.c))

    # oracle parameters for fresh r, s (regenerated with the golden file's toxic waste stream)
    fr = FR[BLS12_381]
    cs = mini_circuit(ConstraintSystem(fr.p))
    alpha, beta, gamma, delta, t = [stream_field(1, i, fr.p) for i in range(5)]
    pk = OG.generate_parameters(cs, BLS12_381, alpha, beta, gamma, delta, t)

    def expect(r, s):
        a, b, c = OG.create_proof(pk, cs, r, s)
        return [H.points_array(BLS12_381, grp, [P]) for grp, P in ((1, a), (2, b), (1, c))]

    def check(proof, r, s):
        for got, (xy, inf) in zip((proof.a, proof.b, proof.c), expect(r, s)):
            assert got[1] == bool(inf[0])
            if not got[1]:
                assert np.array_equal(got[0], xy[0])

    check(zg.create_proof_no_zk(params, circuit), 0, 0)          # r = 0 takes the guard of prover.rs:170
    rng = random.Random(99)
    proof = zg.create_random_proof(params, circuit, rng)
    rng = random.Random(99)
    r = rng.randrange(fr.p)
    s = rng.randrange(fr.p)
    check(proof, r, s)
    check(zg.create_proof(params, circuit, 0, 12345), 0, 12345)
    check(zg.create_proof(params, circuit, 777, 0), 777, 0)
    params.free()


@pytest.mark.parametrize("name,cid", [("groth16_mini_bls12_381", BLS12_381), ("groth16_mini_bn254", BN254)])
def test_gpu_proofs_pass_the_reference_acceptance_test(ctx, name, cid):
    """groth16/tests/mini.rs:46-97 end to end with the GPU prover in the middle: parameters from the (restated) generator,
    `create_random_proof` on the GPU, then the reference's own acceptance test -- prepare_verifying_key + verify_proof
    (groth16/src/verifier.rs:8-44, pairing restated in oracle/pyref/pairing.py) -- on the GPU's proof points."""
    g = load(name)
    params = params_from_golden(ctx, g)
    fr = FR[cid]
    cs = mini_circuit(ConstraintSystem(fr.p))
    alpha, beta, gamma, delta, t = [stream_field(1, i, fr.p) for i in range(5)]
    pk = OG.generate_parameters(cs, cid, alpha, beta, gamma, delta, t)
    to_points = lambda proof: tuple(H.array_point(cid, grp, xy, inf) for grp, (xy, inf) in zip((1, 2, 1), (proof.a, proof.b, proof.c)))
    proof = zg.create_random_proof(params, MiniCircuit(2, 3, 10, 10), random.Random(2024))
    assert OG.verify_proof(pk, to_points(proof), [10])
    assert not OG.verify_proof(pk, to_points(proof), [11])
    assert OG.verify_proof(pk, to_points(zg.create_proof_no_zk(params, MiniCircuit(2, 3, 10, 10))), [10])
    # a witness that does not satisfy the circuit: the GPU still returns a proof (the prover does not check), the verifier rejects it
    assert not OG.verify_proof(pk, to_points(zg.create_random_proof(params, MiniCircuit(2, 4, 10, 10), random.Random(7))), [10])
    params.free()


def test_unsatisfied_witness_follows_the_pipeline(ctx):
    """For a witness that does not satisfy the constraints h is not a true quotient; parity then
    depends on following the literal 7-transform pipeline of r1cs_to_qap.rs:144-169."""
    cid = BN254
    fr = FR[cid]
    cs = ConstraintSystem(fr.p)
    mimc_circuit(cs, 32)
    cs.aux_assignment[5] = (cs.aux_assignment[5] + 1) % fr.p
    assert not cs.is_satisfied()
    want = OG.witness_map(cs, cid)
    mats = []
    for w in "abc":
        ptr, cols, vals = cs.csr(w)
        mats.append(CsrMatrix(np.asarray(ptr, dtype=np.uint32), np.asarray(cols, dtype=np.uint32), H.fr_array(cid, vals)))
    h = ctx.groth16_h(cid, mats[0], mats[1], mats[2], H.fr_array(cid, cs.full_assignment()), cs.num_inputs, cs.num_aux)
    assert H.fr_ints(cid, h, mont=False) == want
    assert want[-1] != 0          # a satisfied witness would give h[N-1] == 0


def test_error_behaviour(ctx):
    from ckb_zkp_b200.backend import ZkbError
    g = load("groth16_mini_bn254")
    A, B, C = matrices(g)
    # row counts differ
    bad = CsrMatrix(g["a_ptr"][:-1], g["a_col"][:int(g["a_ptr"][-2])], g["a_val"][:int(g["a_ptr"][-2])])
    with pytest.raises(ZkbError):
        ctx.groth16_h(BN254, bad, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]))
    with pytest.raises(ValueError):
        ctx.groth16_h(BN254, A, B, C, g["z"][:-1], int(g["n_inputs"]), int(g["n_aux"]))


def test_full_size_witness_map_and_proof(ctx):
    """BASELINE configs[1] shape at 2^17 constraints (domain 2^18): h bit-exact against the C++
    restatement, the proof bit-exact against the C++ prover on the same synthetic key, and the
    proof equal to the in-the-exponent derivation."""
    from ckb_zkp_b200 import synth
    from oracle import cref
    cid, n = BLS12_381, 1 << 17
    inst = synth.MimcInstance(cid, n)
    A, B, C, z = inst.device_form(ctx)
    h = ctx.groth16_h(cid, A, B, C, z, inst.n_inputs, inst.n_aux)
    mats = [(m.row_ptr, m.col_idx, m.coeff) for m in (A, B, C)]
    want_h = cref.witness_map(cid, mats[0], mats[1], mats[2], z, inst.n_inputs)
    assert np.array_equal(h, want_h)
    assert not h[-1].any()                      # satisfied witness: top coefficient of h is zero
    key = synth.SyntheticKey(inst.n_inputs + inst.n_aux, inst.n_inputs, len(h), b_zero_cols=np.arange(4, 4 + n, 2))
    params = key.upload(ctx, cid)
    r, s = synth.ints_to_limbs([0xABCDEF123]), synth.ints_to_limbs([0x13579BDF])
    proof = ctx.groth16_prove(params.pk, A, B, C, z, inst.n_inputs, inst.n_aux, r[0], s[0])
    # in the exponent
    ea, eb, ec = key.expected_exponents(inst.p, inst.z, synth.limbs_to_ints(h), 0xABCDEF123, 0x13579BDF)
    for grp, e, got in ((1, ea, proof[0]), (2, eb, proof[1]), (1, ec, proof[2])):
        xy, inf = ctx.fixed_base_mul(cid, grp, synth.generator_mont(cid, grp), synth.ints_to_limbs([e]))
        assert not got[1] and not inf[0] and np.array_equal(xy[0], got[0])
    # against the CPU restatement of the reference prover on the same key
    g1, g2 = synth.generator_mont(cid, 1), synth.generator_mont(cid, 2)
    pts = lambda grp, k: cref.fixed_base_mul(cid, grp, g1 if grp == 1 else g2, k)
    pk = {"a": pts(1, key.a), "b1": pts(1, key.b), "b2": pts(2, key.b), "h": pts(1, key.h), "l": pts(1, key.l),
          "g1_singles": pts(1, np.stack([key.alpha, key.beta, key.delta]))[0],
          "g2_singles": pts(2, np.stack([key.beta, key.delta]))[0]}
    ref = cref.groth16_prove(cid, pk, mats[0], mats[1], mats[2], z, inst.n_inputs, inst.n_aux, r[0], s[0])
    for got, want in zip(proof, ref):
        assert got[1] == want[1] and np.array_equal(got[0], want[0])
    params.free()


@pytest.mark.parametrize("name", ["groth16_mimc_bls12_381_2e6", "groth16_mimc_bn254_2e10"])
def test_serial_measurement_mode_is_the_same_proof(ctx, name):
    """zkb_set_serial (bench.py's kernel-timing pass: every kernel of the proof on one stream) changes the
    schedule only: the proof is still the golden one, before, during and after the mode."""
    g = load(name)
    params = params_from_golden(ctx, g)
    A, B, C = matrices(g)
    args = (params.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]), g["r"][0], g["s"][0])
    assert_proof(g, ctx.groth16_prove(*args))
    ctx.set_serial(True)
    try:
        assert_proof(g, ctx.groth16_prove(*args))
        ctx.groth16_stage(*args[:7])
        ctx.groth16_prove_staged(params.pk, g["r"][0], g["s"][0])
        assert_proof(g, ctx.groth16_fetch_proof(params.pk))
    finally:
        ctx.set_serial(False)
    assert_proof(g, ctx.groth16_prove(*args))


def test_curve_switch_on_a_fresh_context():
    """the stage of a context outlives a change of curve: BN254 first (the smaller result block), then BLS12-381, then
    BN254 again on a context of its own -- every proof equal to its golden vector (regression: the result block used to
    keep the size of the first curve proven on the context)"""
    from ckb_zkp_b200.backend import Context
    own = Context(0)
    try:
        for name in ("groth16_mini_bn254", "groth16_mimc_bls12_381_2e6", "groth16_mimc_bn254_2e10", "groth16_mini_bls12_381"):
            g = load(name)
            params = params_from_golden(own, g)
            A, B, C = matrices(g)
            proof = own.groth16_prove(params.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]), g["r"][0], g["s"][0])
            assert_proof(g, proof)
            params.free()
    finally:
        own.close()
