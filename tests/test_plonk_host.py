"""Host-side orchestration of ckb_zkp_b200/plonk.py (which primitive is called with which operands) over the mock
backend, i.e. without a GPU: index, the three prover rounds, the linear combinations and the equality check against the
oracle's PLONK AHP (oracle/pyref/plonk.py) -- the same flow tests/test_gpu_plonk.py runs on the device."""
import random

import pytest

from ckb_zkp_b200 import plonk as zp
from oracle.pyref import plonk as OP
from oracle.pyref.fields import BLS12_381, BN254, FR
from tests import helpers as H
from tests.mock_backend import MockContext
from tests.test_oracle_plonk import KS, random_circuit


def mirror(cs_oracle):
    cs = zp.Composer(cs_oracle.p)
    cs.n, cs.pi = cs_oracle.n, list(cs_oracle.pi)
    cs.q = {k: list(v) for k, v in cs_oracle.q.items()}
    cs.w = [list(c) for c in cs_oracle.w]
    cs.variable_map = [list(w) for w in cs_oracle.variable_map]
    cs.assignment = list(cs_oracle.assignment)
    return cs


def ints(cid, arr):
    return H.fr_ints(cid, arr) if len(arr) else []


@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_rounds_over_mock_backend(cid):
    fr = FR[cid]
    p = fr.p
    ctx = MockContext()
    rng = random.Random(cid)
    for ocs in (OP.test_circuit(p), random_circuit(p, 19, 4)):
        cs = mirror(ocs)
        beta, gamma, alpha, zeta = (rng.randrange(p) for _ in range(4))
        oidx, idx = OP.index(ocs, fr, KS), zp.index(ctx, cid, cs, KS)
        for label in OP.SELECTOR_LABELS:
            assert ints(cid, idx.polys[label]) == oidx.polys[label], label
            assert ints(cid, idx.evals_4n[label]) == oidx.evals_4n[label], label
        assert ints(cid, idx.v_4n_inversed) == oidx.v_4n_inversed and ints(cid, idx.l1_4n) == oidx.l1_4n
        ops, ps = OP.prover_init(ocs, oidx), zp.prover_init(ctx, cs, idx)
        polys, opolys = dict(idx.polys), dict(oidx.polys)
        for got, want in ((zp.prover_first_round(ps, cs), OP.prover_first_round(ops, ocs)),
                          (zp.prover_second_round(ps, beta, gamma), OP.prover_second_round(ops, beta, gamma)),
                          (zp.prover_third_round(ps, alpha), OP.prover_third_round(ops, alpha)[0])):
            assert {k: ints(cid, v) for k, v in got.items()} == want
            polys.update(got)
            opolys.update(want)
        lcs = zp.construct_linear_combinations(ctx, idx, beta, gamma, alpha, zeta, polys)
        assert lcs == OP.linear_combinations(oidx, beta, gamma, alpha, zeta, opolys)
        qs = zp.verifier_query_set(idx, zeta)
        evals = {l: zp._eval(ctx, cid, zp.lc_polynomial(ctx, cid, lcs[l], polys), pt) for l, (_, pt) in qs.items()}
        assert OP.verifier_equality_check(oidx, beta, gamma, alpha, zeta, evals, ocs.public_inputs())
        assert zp.verifier_equality_check(ctx, idx, beta, gamma, alpha, zeta, evals, cs.public_inputs())


def test_fiat_shamir_rng_is_blake2s_then_chacha20():
    """plonk/src/rng.rs: seed = Blake2s(material); absorb: Blake2s(material || seed); the stream is ChaCha20 keyed with it"""
    import hashlib

    from ckb_zkp_b200.fs_rng import ChaChaRng
    rng = zp.FiatShamirRng(b"PLONK", BLS12_381)
    seed0 = hashlib.blake2s(b"PLONK").digest()
    assert rng.seed == seed0
    ref = ChaChaRng(seed0)
    assert rng.r.next_u64() == ref.next_u64()
    rng.absorb(b"abc")
    assert rng.seed == hashlib.blake2s(b"abc" + seed0).digest()
    v = rng.rand_fr()
    assert 0 <= v < FR[BLS12_381].p
