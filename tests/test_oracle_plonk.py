"""The oracle's PLONK AHP (oracle/pyref/plonk.py) against the reference's own tests for this layer: `compose`
(plonk/src/composer/mod.rs:91-158), `permutation` (composer/permutation.rs:127-158), `ahp` (ahp/mod.rs:131-205:
the verifier's equality check accepts the prover's oracles) and the quad-split identities (ahp/prover.rs:257-320)."""
import random

import pytest

from oracle.pyref import plonk as OP
from oracle.pyref.fields import BLS12_381, BN254, FR
from oracle.pyref.ntt import Domain

KS = [1, 7, 13, 17]


def random_circuit(p, n_gates, seed):
    """a satisfied circuit of add / mul / constant gates over a growing pool of variables (some gates use the aux wire)"""
    rng = random.Random(seed)
    cs = OP.Composer(p)
    vals = [rng.randrange(p) for _ in range(4)]
    pool = [cs.alloc_and_assign(v) for v in vals]
    val = dict(zip(pool, vals))
    for g in range(n_gates):
        l, r = rng.choice(pool), rng.choice(pool)
        kind = rng.randrange(4)
        if kind == 0:
            ql, qr, qc = rng.randrange(p), rng.randrange(p), rng.randrange(p)
            aux = (rng.choice(pool), rng.randrange(p)) if rng.random() < 0.3 else None
            pi = rng.randrange(p) if rng.random() < 0.2 else 0
            out = (ql * val[l] + qr * val[r] + qc + pi + (aux[1] * val[aux[0]] if aux else 0)) % p
            o = cs.alloc_and_assign(out)
            cs.create_add_gate((l, ql), (r, qr), o, aux, qc, pi)
        elif kind == 1:
            qm, qc = rng.randrange(1, p), rng.randrange(p)
            out = (qm * val[l] * val[r] + qc) % p
            o = cs.alloc_and_assign(out)
            cs.create_mul_gate(l, r, o, None, qm, qc, 0)
        elif kind == 2:
            o = cs.alloc_and_assign(val[l])
            out = val[l]
            cs.assert_equal(l, o)
        else:
            pi = rng.randrange(p)
            out = val[l]
            o = l
            cs.constrain_to_constant(l, (val[l] - pi) % p, pi)
        val[o] = out
        if o not in pool:
            pool.append(o)
    return cs


@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_compose_like_the_reference(cid):
    fr = FR[cid]
    p = fr.p
    for cs in (OP.test_circuit(p), random_circuit(p, 23, 1)):
        n, s = cs.compose(fr, KS)
        pi = cs.public_inputs() + [0] * (n - cs.size())
        w = cs.synthesize(fr)
        for i in range(n):                                   # arithmetic (composer/mod.rs:108-120)
            assert (w[0][i] * s["q_0"][i] + w[1][i] * s["q_1"][i] + w[2][i] * s["q_2"][i] + w[3][i] * s["q_3"][i]
                    + w[1][i] * w[2][i] * s["q_m"][i] + s["q_c"][i] + pi[i]) % p == 0
        dom = Domain(fr, cs.size())
        roots = [pow(dom.group_gen, i, p) for i in range(n)]
        beta, gamma = 0x1234567, 0x89ABCDEF
        num = den = 1
        for i in range(n):                                   # permutation (:122-157)
            for k in range(4):
                num = num * (w[k][i] + beta * roots[i] * KS[k] + gamma) % p
                den = den * (w[k][i] + beta * s["sigma_%d" % k][i] + gamma) % p
        assert num == den
        sig = idn = 1                                        # composer/permutation.rs:127-158
        for k in range(4):
            for i in range(n):
                sig = sig * s["sigma_%d" % k][i] % p
                idn = idn * (KS[k] * roots[i]) % p
        assert sig == idn


@pytest.mark.parametrize("cid", [BLS12_381, BN254])
def test_ahp_equality_check_accepts(cid):
    fr = FR[cid]
    rng = random.Random(cid)
    for cs in (OP.test_circuit(fr.p), random_circuit(fr.p, 40, 2)):
        beta, gamma, alpha, zeta = (rng.randrange(fr.p) for _ in range(4))
        ok, evals, polys = OP.run_ahp(cs, fr, KS, beta, gamma, alpha, zeta)
        assert ok
        idx = OP.index(cs, fr, KS)
        bad = dict(evals)
        bad["w_1"] = (bad["w_1"] + 1) % fr.p
        assert not OP.verifier_equality_check(idx, beta, gamma, alpha, zeta, bad, cs.public_inputs())
        # quad split: t(zeta) = sum zeta^(k n) t_k(zeta)  (ahp/prover.rs:257-320)
        t_full = []
        for k in range(4):
            t_full += polys["t_%d" % k] + [0] * (idx.n - len(polys["t_%d" % k]))
        assert OP.poly_eval(t_full, zeta, fr.p) == evals["t"]


def test_unsatisfied_circuit_is_rejected():
    fr = FR[BLS12_381]
    cs = OP.test_circuit(fr.p)
    cs.assignment[3] = 5                                     # var_three no longer 1 + 2: the copy constraints still hold,
    ok, _, _ = OP.run_ahp(cs, fr, KS, 11, 12, 13, 14)        # the arithmetic identity does not
    assert not ok
