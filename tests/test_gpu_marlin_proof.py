"""zkp_marlin::create_random_proof on the GPU (ckb_zkp_b200.marlin.create_random_proof, marlin/src/lib.rs:97-181) with the
Fiat-Shamir generator of marlin/src/fs_rng.rs in the loop, against the oracle's restatement of the same function
(oracle/pyref/marlin_proof.py): every commitment, evaluation and opening proof bit for bit, the challenges drawn from the
transcript, and then the reference's own acceptance test -- verify_proof (lib.rs:184-260: AHP equality check + KZG
pairing checks, marlin/tests/mini.rs:81,87) -- on the GPU's proof."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import marlin as zm
from oracle.pyref import marlin as OM
from oracle.pyref import marlin_proof as MP
from oracle.pyref.fields import BLS12_381, BN254, FR, stream_field
from tests import helpers as H

pytestmark = pytest.mark.gpu
ONE = ("in", 0)


class Mini:
    """marlin/tests/mini.rs:12-41 (x = 2, y = 3, z = 10, num constraints = 10): more constraints than variables, so
    make_matrices_square appends padding variables"""

    def generate_constraints(self, cs):
        vx, vy = cs.alloc(2), cs.alloc(3)
        vz = cs.alloc_input(10)
        for _ in range(10):
            cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])


class Mimc:
    """MiMC chain shaped like marlin/examples/mimc.rs:26-118: more variables than constraints (padding constraints)"""

    def __init__(self, p, n, seed=5):
        self.p, self.n, self.seed = p, n, seed

    def generate_constraints(self, cs):
        p, seed = self.p, self.seed
        xl_v, xr_v = stream_field(seed, 0, p), stream_field(seed, 1, p)
        xl, xr = cs.alloc(xl_v), cs.alloc(xr_v)
        for i in range(self.n // 2):
            c = stream_field(seed, 2 + i, p)
            tmp_v = (xl_v + c) * (xl_v + c) % p
            tmp = cs.alloc(tmp_v)
            cs.enforce([(1, xl), (c, ONE)], [(1, xl), (c, ONE)], [(1, tmp)])
            new_v = ((xl_v + c) * tmp_v + xr_v) % p
            new = cs.alloc_input(new_v) if i == self.n // 2 - 1 else cs.alloc(new_v)
            cs.enforce([(1, tmp)], [(1, xl), (c, ONE)], [(1, new), (p - 1, xr)])
            xr, xr_v, xl, xl_v = xl, xl_v, new, new_v


def _point(cid, pt):
    return H.array_point(cid, 1, pt[0], pt[1])


@pytest.mark.parametrize("resident", [True, False])
@pytest.mark.parametrize("cid,make", [(BLS12_381, lambda p: Mini()), (BN254, lambda p: Mini()), (BN254, lambda p: Mimc(p, 28)),
                                      (BLS12_381, lambda p: Mimc(p, 120))])
def test_create_random_proof_matches_oracle_and_verifies(ctx, cid, make, resident):
    p = FR[cid].p
    circuit = make(p)
    # oracle side
    cs = OM.MarlinCS(p)
    circuit.generate_constraints(cs)
    setup_rng = random.Random(77)
    beta, kg, kgamma, kh = (setup_rng.randrange(1, p) for _ in range(4))
    probe = OM.MarlinCS(p)
    circuit.generate_constraints(probe)
    oidx = OM.index(probe, cid)
    need = MP.max_degree(oidx["num_constraints"], oidx["num_variables"], oidx["num_non_zeros"])
    size = 1 << max(need - 1, 0).bit_length()
    opp = MP.universal_setup(cid, size, beta, kg, kgamma, kh)
    oipk, oivk = MP.index(opp, cs)
    want = MP.create_random_proof(oipk, cs, random.Random(4242))
    assert MP.verify_proof(oivk, want, cs.input[1:])                  # the oracle's own proof passes the acceptance test

    # GPU side: same setup draws, same prover randomness, transcript restated in ckb_zkp_b200/fs_rng.py
    srs = zm.universal_setup(ctx, cid, need, random.Random(77))
    assert srs.max_degree() == size
    ipk, ivk = zm.index_keys(ctx, srs, circuit)
    assert ivk.index_info == oivk["index_info"] and ivk.verifier_key.supported_degree == need
    for got, (c, sh) in zip(ivk.index_comms, oivk["index_comms"]):
        assert _point(cid, got[0]) == c and got[1] is None and sh is None
    assert ivk.to_bytes() == MP.ivk_bytes(oivk)                       # the seed material of the transcript
    proof = zm.create_random_proof(ctx, ipk, circuit, random.Random(4242), resident=resident)
    assert proof.challenges == want["challenges"]
    for got_round, want_round in zip(proof.commitments, want["commitments"]):
        assert len(got_round) == len(want_round)
        for (gc, gs), (wc, ws) in zip(got_round, want_round):
            assert _point(cid, gc) == wc
            assert (gs is None) == (ws is None) and (gs is None or _point(cid, gs) == ws)
    assert H.fr_ints(cid, np.stack(proof.evaluations)) == want["evaluations"] and len(want["evaluations"]) == 21
    assert len(proof.opening_proofs) == 2
    for (gw, grv), (ww, wrv) in zip(proof.opening_proofs, want["opening_proofs"]):
        assert _point(cid, gw) == ww
        assert (grv is None) == (wrv is None) and (grv is None or H.fr_ints(cid, grv.reshape(1, 4))[0] == wrv)
    # the reference's acceptance test on what the GPU produced
    gpu_proof = {"commitments": [[(_point(cid, c), None if s is None else _point(cid, s)) for c, s in rnd] for rnd in proof.commitments],
                 "evaluations": H.fr_ints(cid, np.stack(proof.evaluations)),
                 "opening_proofs": [(_point(cid, w), None if rv is None else H.fr_ints(cid, rv.reshape(1, 4))[0])
                                    for w, rv in proof.opening_proofs]}
    assert MP.verify_proof(oivk, gpu_proof, cs.input[1:])
    assert not MP.verify_proof(oivk, gpu_proof, [(v + 1) % p for v in cs.input[1:]])
    # and the product's own verifier (lib.rs:183-260 with the pairings of PC::batch_check on the GPU)
    assert zm.verify_proof(ctx, ivk, proof, cs.input[1:])
    assert zm.verify_proof(ctx, ivk, proof, H.fr_array(cid, cs.input[1:]))          # Montgomery array form of the same input
    assert not zm.verify_proof(ctx, ivk, proof, [(v + 1) % p for v in cs.input[1:]])
    w0, rv0 = proof.opening_proofs[0]
    forged = zm.Proof(proof.commitments, proof.evaluations, [(proof.opening_proofs[1][0], rv0), proof.opening_proofs[1]])
    assert not zm.verify_proof(ctx, ivk, forged, cs.input[1:])                      # equality check passes, a pairing check fails
    ipk.committer_key.free()
