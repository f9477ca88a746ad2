"""Pairing and the Groth16 acceptance test (TEST INFRASTRUCTURE: part of the CPU oracle, never imported by the product).

Restates what groth16/src/verifier.rs:8-44 (`prepare_verifying_key`, `verify_proof`) obtains from ark-ec 0.2's
`PairingEngine::{miller_loop, final_exponentiation, pairing}` (un-vendored):

    e(A, B) * e(g_ic, -gamma) * e(C, -delta) == e(alpha, beta),   g_ic = gamma_abc[0] + sum_i x_i * gamma_abc[i + 1]

The reference accepts a proof iff that holds; its own tests pin nothing else (groth16/tests/mini.rs:89,96).  The
equation is invariant under the choice of non-degenerate bilinear pairing on (G1, G2), so the oracle uses the plain
ate pairing  a(Q, P) = f_{|t - 1|, Q}(P) ^ ((q^12 - 1) / r)  for both curves (t = trace of Frobenius; t - 1 = x for
BLS12-381, 6 x^2 for BN254) instead of arkworks' optimal-ate variants: no Frobenius line steps, one code path, and the
final exponentiation is a plain modular power.  PARITY UNPINNED at the byte level of GT (a different power of the same
pairing); what is checked against the reference is the accept / reject decision.

Representation: Fq12 = Fq[w] / (w^12 - 2 c w^6 + c^2 + 1) with w^6 = xi = c + u, u^2 = -1 (c = 1: BLS12-381, c = 9: BN254),
elements are 12-coefficient lists.  G2 points stay on the twist (affine, Fq2); the line through psi(T) is evaluated at
P in G1 as a sparse element (untwist psi(x', y') = (x' w^2, y' w^3) for the D-type twist of BN254 and
(x' / w^2, y' / w^3) for the M-type twist of BLS12-381; factors in Fq2 are dropped, the final exponentiation kills them).
"""
from .curves import CURVES
from .fields import BLS12_381, BN254, FQ, FR

_X = {BLS12_381: -0xD201000000010000, BN254: 4965661367192848881}
_C = {BLS12_381: 1, BN254: 9}
_TWIST_D = {BLS12_381: False, BN254: True}


def ate_loop_count(cid):
    x = _X[cid]
    return abs(x) if cid == BLS12_381 else 6 * x * x


class Fq12:
    def __init__(self, cid):
        self.cid, self.p, self.c = cid, FQ[cid].p, _C[cid]
        self.one = [1] + [0] * 11

    def mul(self, a, b):
        p, c = self.p, self.c
        t = [0] * 23
        for i, ai in enumerate(a):
            if ai:
                for j, bj in enumerate(b):
                    if bj:
                        t[i + j] += ai * bj
        # w^12 = 2 c w^6 - (c^2 + 1)
        k2, k0 = 2 * c, c * c + 1
        for d in range(22, 11, -1):
            v = t[d]
            if v:
                t[d - 6] += k2 * v
                t[d - 12] -= k0 * v
        return [x % p for x in t[:12]]

    def sqr(self, a):
        return self.mul(a, a)

    def pow(self, a, e):
        r = self.one
        for i in reversed(range(e.bit_length())):
            r = self.sqr(r)
            if (e >> i) & 1:
                r = self.mul(r, a)
        return r

    def embed_fq2(self, z, k=0):
        """(a + b u) * w^k  with u = w^6 - c"""
        a, b = z
        out = [0] * 12
        out[k] = (a - self.c * b) % self.p
        out[k + 6] = b % self.p
        return out


def _line(F12, F2, cid, T, lam, P):
    """line through psi(T) with twist-slope lam, evaluated at P = (xP, yP) in G1, up to a factor in Fq2"""
    xP, yP = P
    p = F12.p
    t3 = F2.sub(F2.mul(lam, T[0]), T[1])                      # lam x'_T - y'_T
    if _TWIST_D[cid]:
        # yP - lam xP w + (lam x'_T - y'_T) w^3
        out = [yP % p] + [0] * 11
        a = F12.embed_fq2(F2.neg(((lam[0] * xP) % p, (lam[1] * xP) % p)), 1)
        b = F12.embed_fq2(t3, 3)
    else:
        # xi yP - lam xP w^5 + (lam x'_T - y'_T) w^3      (the whole line scaled by xi)
        out = F12.embed_fq2((yP * F12.c % p, yP % p), 0)      # xi * yP = (c + u) yP
        a = F12.embed_fq2(F2.neg(((lam[0] * xP) % p, (lam[1] * xP) % p)), 5)
        b = F12.embed_fq2(t3, 3)
    return [(x + y + z) % p for x, y, z in zip(out, a, b)]


def miller_loop(cid, P, Q):
    """f_{|t-1|, Q}(P) for P in G1, Q in G2 (affine tuples, None = identity -> 1)"""
    F12 = Fq12(cid)
    if P is None or Q is None:
        return F12.one
    F2 = CURVES[(cid, 2)].F
    n = ate_loop_count(cid)
    f = F12.one
    T = Q
    three = F2.small(3)
    for i in reversed(range(n.bit_length() - 1)):
        lam = F2.mul(F2.mul(three, F2.sqr(T[0])), F2.inv(F2.add(T[1], T[1])))
        f = F12.mul(F12.sqr(f), _line(F12, F2, cid, T, lam, P))
        x3 = F2.sub(F2.sqr(lam), F2.add(T[0], T[0]))
        T = (x3, F2.sub(F2.mul(lam, F2.sub(T[0], x3)), T[1]))
        if (n >> i) & 1:
            if T[0] == Q[0]:
                # T == -Q only at the very end of a loop over the group order; cannot happen for t - 1 < r
                raise ArithmeticError("vertical line in the Miller loop")
            lam = F2.mul(F2.sub(Q[1], T[1]), F2.inv(F2.sub(Q[0], T[0])))
            f = F12.mul(f, _line(F12, F2, cid, T, lam, P))
            x3 = F2.sub(F2.sub(F2.sqr(lam), T[0]), Q[0])
            T = (x3, F2.sub(F2.mul(lam, F2.sub(T[0], x3)), T[1]))
    return f


def _frobenius_twist(cid, Q):
    """pi_q on a point of the (D-type) twist: psi^-1 o Frobenius o psi, (x', y') -> (conj(x') xi^((q-1)/3), conj(y') xi^((q-1)/2))"""
    F2 = CURVES[(cid, 2)].F
    q, c = FQ[cid].p, _C[cid]

    def f2_pow(a, e):
        r = (1, 0)
        while e:
            if e & 1:
                r = F2.mul(r, a)
            a = F2.sqr(a)
            e >>= 1
        return r

    conj = lambda z: (z[0], (-z[1]) % q)
    g2, g3 = f2_pow((c, 1), (q - 1) // 3), f2_pow((c, 1), (q - 1) // 2)
    return (F2.mul(conj(Q[0]), g2), F2.mul(conj(Q[1]), g3))


def miller_loop_optimal_bn(cid, P, Q):
    """The optimal ate Miller function of a BN curve (Vercauteren): f_{6x+2,Q}(P) * l_{[6x+2]Q, pi(Q)}(P) *
    l_{[6x+2]Q + pi(Q), -pi^2(Q)}(P), x > 0 -- half the loop length of miller_loop.  Same affine steps and line scaling
    as miller_loop; this is what csrc/pairing.cuh runs on BN254 (a different power of the same pairing)."""
    assert cid == BN254 and _X[cid] > 0
    F12 = Fq12(cid)
    if P is None or Q is None:
        return F12.one
    g2 = CURVES[(cid, 2)]
    F2 = g2.F
    n = 6 * _X[cid] + 2
    f, T = F12.one, Q
    three = F2.small(3)

    def add_step(f, T, R):
        lam = F2.mul(F2.sub(R[1], T[1]), F2.inv(F2.sub(R[0], T[0])))
        f = F12.mul(f, _line(F12, F2, cid, T, lam, P))
        x3 = F2.sub(F2.sub(F2.sqr(lam), T[0]), R[0])
        return f, (x3, F2.sub(F2.mul(lam, F2.sub(T[0], x3)), T[1]))

    for i in reversed(range(n.bit_length() - 1)):
        lam = F2.mul(F2.mul(three, F2.sqr(T[0])), F2.inv(F2.add(T[1], T[1])))
        f = F12.mul(F12.sqr(f), _line(F12, F2, cid, T, lam, P))
        x3 = F2.sub(F2.sqr(lam), F2.add(T[0], T[0]))
        T = (x3, F2.sub(F2.mul(lam, F2.sub(T[0], x3)), T[1]))
        if (n >> i) & 1:
            f, T = add_step(f, T, Q)
    Q1 = _frobenius_twist(cid, Q)
    Q2 = g2.neg_affine(_frobenius_twist(cid, Q1))
    f, T = add_step(f, T, Q1)
    f, T = add_step(f, T, Q2)
    return f


def device_miller_loop(cid, P, Q):
    """the Miller function csrc/pairing.cuh computes: optimal ate on BN254, plain ate (loop |x|) on BLS12-381"""
    return miller_loop_optimal_bn(cid, P, Q) if cid == BN254 else miller_loop(cid, P, Q)


def device_multi_pairing(cid, pairs):
    """what zkb_multi_pairing returns for one group: the product of the device's Miller functions to the power
    m (q^12 - 1) / r, m = 3 on BLS12-381 (x-chain of the hard part), 1 on BN254"""
    F12 = Fq12(cid)
    f = F12.one
    for P, Q in pairs:
        f = F12.mul(f, device_miller_loop(cid, P, Q))
    return F12.pow(final_exponentiation(cid, f), 3 if cid == BLS12_381 else 1)


def final_exponentiation(cid, f):
    q, r = FQ[cid].p, FR[cid].p
    return Fq12(cid).pow(f, (q ** 12 - 1) // r)


def pairing(cid, P, Q):
    return final_exponentiation(cid, miller_loop(cid, P, Q))


def multi_pairing(cid, pairs):
    """prod e(P_i, Q_i): the Miller loops multiplied, one final exponentiation (verifier.rs:31-41)"""
    F12 = Fq12(cid)
    f = F12.one
    for P, Q in pairs:
        f = F12.mul(f, miller_loop(cid, P, Q))
    return final_exponentiation(cid, f)


# ------------------------------------------------------------------------------------------------
# groth16/src/verifier.rs
# ------------------------------------------------------------------------------------------------
def prepare_verifying_key(cid, vk):
    """vk: dict alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1 (affine tuples)  (verifier.rs:8-16)"""
    g2 = CURVES[(cid, 2)]
    return {"vk": vk, "alpha_g1_beta_g2": pairing(cid, vk["alpha_g1"], vk["beta_g2"]),
            "gamma_g2_neg": g2.neg_affine(vk["gamma_g2"]), "delta_g2_neg": g2.neg_affine(vk["delta_g2"]),
            "gamma_abc_g1": vk["gamma_abc_g1"]}


class MalformedVerifyingKey(Exception):
    pass


def verify_proof(cid, pvk, proof, public_inputs):
    """verifier.rs:18-44; proof = (a, b, c) affine tuples"""
    if len(public_inputs) + 1 != len(pvk["gamma_abc_g1"]):
        raise MalformedVerifyingKey()
    g1 = CURVES[(cid, 1)]
    g_ic = g1.from_affine(pvk["gamma_abc_g1"][0])
    for x, b in zip(public_inputs, pvk["gamma_abc_g1"][1:]):
        g_ic = g1.add(g_ic, g1.mul(g1.from_affine(b), x % g1.r))
    a, b, c = proof
    test = multi_pairing(cid, [(a, b), (g1.to_affine(g_ic), pvk["gamma_g2_neg"]), (c, pvk["delta_g2_neg"])])
    return test == pvk["alpha_g1_beta_g2"]
