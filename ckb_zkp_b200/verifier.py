"""The reference's Groth16 verifier API (groth16/src/verifier.rs) on the B200 backend, plus the batched form SURVEY.md 8f-4
names: many proofs against one verifying key in one device call.

    prepare_verifying_key(ctx, curve, vk)            verifier.rs:8-16
    verify_proof(pvk, proof, public_inputs)          verifier.rs:18-44
    verify_proofs(pvk, proofs, public_inputs_list)   the same test for every proof: one batched MSM call for the g_ic,
                                                     one zkb_multi_pairing call for the 3 * B Miller loops and B final
                                                     exponentiations
    verify_proofs_batched(pvk, proofs, inputs, rng)  ONE decision for the batch by a random linear combination of the B
                                                     equations: B + 3 Miller loops and one final exponentiation

    e(A, B) * e(g_ic, -gamma) * e(C, -delta) == e(alpha, beta),     g_ic = gamma_abc[0] + sum_i x_i * gamma_abc[i + 1]
"""
import numpy as np

from . import _lib
from . import pairing as _pairing
from .groth16 import FR_MODULUS
from .r1cs import SynthesisError, ints_to_limbs


class MalformedVerifyingKey(SynthesisError):
    """SynthesisError::MalformedVerifyingKey (verifier.rs:23-25)"""


class PreparedVerifyingKey:
    """groth16/src/lib.rs:94-101; gamma_abc_g1 is kept resident as an MSM base set"""

    def __init__(self, ctx, curve, vk, alpha_g1_beta_g2, gamma_g2_neg, delta_g2_neg, gamma_abc_srs):
        self.ctx, self.curve, self.vk = ctx, curve, vk
        self.alpha_g1_beta_g2, self.gamma_g2_neg, self.delta_g2_neg = alpha_g1_beta_g2, gamma_g2_neg, delta_g2_neg
        self.gamma_abc_srs = gamma_abc_srs
        self.n_gamma_abc = len(vk.gamma_abc_g1[1])

    def free(self):
        self.gamma_abc_srs.free()


def prepare_verifying_key(ctx, curve, vk):
    """vk: generator.VerifyKey (single points (xy, is_identity), gamma_abc_g1 = (xy[n], inf[n]))"""
    xy, inf = vk.gamma_abc_g1
    srs = ctx.srs_upload(curve, _lib.G1, xy, inf, precompute=False)
    return PreparedVerifyingKey(ctx, curve, vk, _pairing.pairing(ctx, curve, vk.alpha_g1, vk.beta_g2),
                                _pairing.neg_point(curve, _lib.G2, vk.gamma_g2),
                                _pairing.neg_point(curve, _lib.G2, vk.delta_g2), srs)


def _proof_arrays(proofs, which):
    return (np.stack([np.asarray(getattr(pr, which)[0], dtype=np.uint64).reshape(-1) for pr in proofs]),
            np.array([1 if getattr(pr, which)[1] else 0 for pr in proofs], dtype=np.uint8))


def verify_proofs(pvk, proofs, public_inputs_list):
    """-> [bool] in order.  public inputs are canonical ints (E::Fr), one list per proof."""
    if len(proofs) != len(public_inputs_list):
        raise ValueError("one public-input list per proof")
    for x in public_inputs_list:
        if len(x) + 1 != pvk.n_gamma_abc:
            raise MalformedVerifyingKey()
    if not proofs:
        return []
    ctx, curve = pvk.ctx, pvk.curve
    p = FR_MODULUS[curve]
    B = len(proofs)
    # g_ic (verifier.rs:27-30): the MSM of gamma_abc_g1 with the scalars (1, x_1, ..., x_n), all proofs in one call
    scalars = ints_to_limbs([v for x in public_inputs_list for v in [1] + [int(e) % p for e in x]]).reshape(B, pvk.n_gamma_abc, 4)
    g_xy, g_inf = ctx.msm_many(pvk.gamma_abc_srs, scalars)
    (a_xy, a_inf), (b_xy, b_inf), (c_xy, c_inf) = (_proof_arrays(proofs, w) for w in "abc")
    g1 = (np.stack([a_xy, g_xy, c_xy], axis=1).reshape(3 * B, -1), np.stack([a_inf, g_inf, c_inf], axis=1).reshape(-1))
    fixed = lambda q: np.broadcast_to(np.asarray(q[0], dtype=np.uint64).reshape(1, -1), b_xy.shape)
    g2 = (np.stack([b_xy, fixed(pvk.gamma_g2_neg), fixed(pvk.delta_g2_neg)], axis=1).reshape(3 * B, -1),
          np.stack([b_inf, np.zeros_like(b_inf), np.zeros_like(b_inf)], axis=1).reshape(-1))
    tests = ctx.multi_pairing(curve, g1, g2, 3)                            # verifier.rs:31-41
    return [bool(t) for t in (tests == pvk.alpha_g1_beta_g2[None, :]).all(axis=1)]   # :43


def verify_proof(pvk, proof, public_inputs):
    return verify_proofs(pvk, [proof], [public_inputs])[0]


def verify_proofs_batched(pvk, proofs, public_inputs_list, rng):
    """True iff every proof verifies, except with probability ~2^-128 over `rng` (needs getrandbits): the B equations
    are raised to random 128-bit exponents r_i and multiplied, and bilinearity moves the exponents into G1:

        prod_i e(r_i A_i, B_i) * e(sum_i r_i g_ic_i, -gamma) * e(sum_i r_i C_i, -delta) * e(-(sum_i r_i) alpha, beta) == 1

    B + 3 Miller loops and ONE final exponentiation instead of 3 B and B (verifier.rs:31-43 per proof); the scalar
    multiplications r_i A_i are one thread each (zkb_msm_batch, short-MSM path from 32 proofs up), sum r_i C_i is one MSM over the B points,
    sum r_i g_ic_i one MSM over gamma_abc_g1 with the scalars (sum_i r_i, sum_i r_i x_i1, ...).  A False does not say
    which proof failed: fall back to verify_proofs for that."""
    if len(proofs) != len(public_inputs_list):
        raise ValueError("one public-input list per proof")
    for x in public_inputs_list:
        if len(x) + 1 != pvk.n_gamma_abc:
            raise MalformedVerifyingKey()
    if not proofs:
        return True
    ctx, curve, vk = pvk.ctx, pvk.curve, pvk.vk
    p = FR_MODULUS[curve]
    B = len(proofs)
    rs = [rng.getrandbits(128) | 1 for _ in range(B)]
    (a_xy, a_inf), (b_xy, b_inf), (c_xy, c_inf) = (_proof_arrays(proofs, w) for w in "abc")
    r_limbs = ints_to_limbs(rs)
    srs_a = ctx.srs_upload(curve, _lib.G1, a_xy, a_inf, precompute=False)
    srs_c = ctx.srs_upload(curve, _lib.G1, c_xy, c_inf, precompute=False)
    try:
        ra_xy, ra_inf = ctx.msm_many(srs_a, r_limbs.reshape(B, 1, 4), np.arange(B))                  # r_i * A_i
        c_sum = ctx.msm(srs_c, r_limbs)                                                              # sum r_i C_i
    finally:
        srs_a.free()
        srs_c.free()
    s = [sum(rs) % p] + [sum(r * (int(x[j]) % p) for r, x in zip(rs, public_inputs_list)) % p
                         for j in range(pvk.n_gamma_abc - 1)]
    g_ic_sum = ctx.msm(pvk.gamma_abc_srs, ints_to_limbs(s))                                          # sum r_i g_ic_i
    alpha_xy, alpha_inf = ctx.fixed_base_mul(curve, _lib.G1, vk.alpha_g1[0], ints_to_limbs([(p - s[0]) % p]))
    g1_xy = np.concatenate([ra_xy, g_ic_sum[0][None, :], c_sum[0][None, :], alpha_xy])
    g1_inf = np.concatenate([ra_inf, np.array([int(g_ic_sum[1]), int(c_sum[1]), int(alpha_inf[0])], dtype=np.uint8)])
    g2_xy = np.concatenate([b_xy, np.stack([np.asarray(q[0], dtype=np.uint64).reshape(-1)
                                            for q in (pvk.gamma_g2_neg, pvk.delta_g2_neg, vk.beta_g2)])])
    g2_inf = np.concatenate([b_inf, np.array([int(q[1]) for q in (pvk.gamma_g2_neg, pvk.delta_g2_neg, vk.beta_g2)], dtype=np.uint8)])
    gt = ctx.multi_pairing(curve, (g1_xy, g1_inf), (g2_xy, g2_inf), B + 3)
    return bool(np.array_equal(gt[0], _pairing.gt_one(curve)))

