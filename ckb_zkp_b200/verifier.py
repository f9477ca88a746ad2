"""The reference's Groth16 verifier API (groth16/src/verifier.rs) on the B200 backend, plus the batched form SURVEY.md 8f-4
names: many proofs against one verifying key in one device call.

    prepare_verifying_key(ctx, curve, vk)            verifier.rs:8-16
    verify_proof(pvk, proof, public_inputs)          verifier.rs:18-44
    verify_proofs(pvk, proofs, public_inputs_list)   the same test for every proof: one batched MSM call for the g_ic,
                                                     one zkb_multi_pairing call for the 3 * B Miller loops and B final
                                                     exponentiations

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
    # g_ic (verifier.rs:27-30): the MSM of gamma_abc_g1 with the scalars (1, x_1, ..., x_n)
    scalars = [ints_to_limbs([1] + [int(v) % p for v in x]) for x in public_inputs_list]
    g_ic = ctx.msm_batch([pvk.gamma_abc_srs] * len(proofs), scalars)
    groups = [[(pr.a, pr.b), (g, pvk.gamma_g2_neg), (pr.c, pvk.delta_g2_neg)] for pr, g in zip(proofs, g_ic)]
    tests = _pairing.multi_pairing(ctx, curve, groups)                    # verifier.rs:31-41
    return [bool(np.array_equal(t, pvk.alpha_g1_beta_g2)) for t in tests]  # :43


def verify_proof(pvk, proof, public_inputs):
    return verify_proofs(pvk, [proof], [public_inputs])[0]
