"""The reference's Groth16 prover API (`zkp_groth16`, groth16/src/prover.rs:97-228) on top of the
B200 backend: same entry points, same argument meaning, same error behaviour.

    create_random_proof(params, circuit, rng)   prover.rs:97-111
    create_proof_no_zk(params, circuit)         prover.rs:113-122
    create_proof(params, circuit, r, s)         prover.rs:124-211

`Parameters` holds the proving key resident in HBM (uploaded once, reused by every proof).  All
arithmetic between "prover filled" (prover.rs:146) and "Proof assembled" (:206) runs on the GPU
through one C-ABI call (zkb_groth16_prove).
"""
import numpy as np

from . import _lib
from .backend import Context, CsrMatrix, point_words
from .r1cs import (ONE, PolynomialDegreeTooLarge, ProvingAssignment, SynthesisError, ints_to_limbs, limbs_to_int)

FR_MODULUS = {
    _lib.BLS12_381: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
}


class Proof:
    """groth16/src/lib.rs:51-56: a, c in G1, b in G2 -- affine, Montgomery u64 limbs + identity flag."""

    def __init__(self, a, b, c):
        self.a, self.b, self.c = a, b, c

    def __eq__(self, o):
        return all(np.array_equal(x[0], y[0]) and x[1] == y[1] for x, y in ((self.a, o.a), (self.b, o.b), (self.c, o.c)))


class Parameters:
    """groth16/src/lib.rs:81-91 made resident on one GPU.

    Point arrays are (xy uint64[n, words], inf uint8[n]) pairs in the layout of include/zkb.h."""

    def __init__(self, ctx, curve, a_query, b_g1_query, b_g2_query, h_query, l_query, alpha_g1, beta_g1, delta_g1,
                 beta_g2, delta_g2, shard=None):
        """shard = (n_ranks, rank): one process per GPU, every rank passes the whole key and keeps only its slice of
        the pairs of each MSM resident; every rank then calls create_proof with the same arguments (collective) and
        gets the same proof (zkb_groth16_prove_sharded)."""
        self.ctx, self.curve, self.shard = ctx, curve, shard
        self.n_a, self.n_h, self.n_l = len(a_query[1]), len(h_query[1]), len(l_query[1])
        g1s = np.stack([np.asarray(x, dtype=np.uint64).reshape(point_words(curve, _lib.G1))
                        for x in (alpha_g1, beta_g1, delta_g1)])
        g2s = np.stack([np.asarray(x, dtype=np.uint64).reshape(point_words(curve, _lib.G2)) for x in (beta_g2, delta_g2)])
        self.pk = ctx.groth16_pk(curve, a_query, b_g1_query, b_g2_query, h_query, l_query, g1s, g2s, shard=shard)

    def free(self):
        self.pk.free()


def _synthesize(params, circuit):
    prover = ProvingAssignment(FR_MODULUS[params.curve])
    prover.alloc_input(1)                       # prover.rs:143
    circuit.generate_constraints(prover)        # prover.rs:146 (SynthesisError propagates)
    return prover


def prove_assignment(params, prover, r, s):
    """prover.rs:148-210 for an already synthesised ProvingAssignment (or any object exposing
    csr_arrays(): the CSR matrices and the assignment as Montgomery limb arrays)."""
    ctx = params.ctx
    curve = params.curve
    p = FR_MODULUS[curve]
    if hasattr(prover, "csr_arrays"):
        A, B, C, z_mont, n_inputs, n_aux = prover.csr_arrays()
    else:
        mats = []
        for which in "abc":
            ptr, cols, coeffs = prover.csr(which)
            mats.append(CsrMatrix(ptr, cols, ctx.fr_convert(curve, ints_to_limbs(coeffs), to_mont=True)))
        A, B, C = mats
        n_inputs, n_aux = prover.num_inputs, prover.num_aux
        z_mont = ctx.fr_convert(curve, ints_to_limbs(prover.input_assignment + prover.aux_assignment), to_mont=True)
    from .backend import ZkbError
    try:
        prove = ctx.groth16_prove_sharded if getattr(params, "shard", None) else ctx.groth16_prove
        a, b, c = prove(params.pk, A, B, C, z_mont, n_inputs, n_aux, ints_to_limbs([r % p])[0], ints_to_limbs([s % p])[0])
    except ZkbError as e:
        if e.code == _lib.E_TOO_LARGE:          # EvaluationDomain::new -> None (r1cs_to_qap.rs:123-125)
            raise PolynomialDegreeTooLarge() from e
        raise
    return Proof(a, b, c)


def create_proof(params, circuit, r, s):
    return prove_assignment(params, _synthesize(params, circuit), r, s)


def create_proof_no_zk(params, circuit):
    return create_proof(params, circuit, 0, 0)


def create_random_proof(params, circuit, rng):
    """`rng` needs randrange(); r and s are drawn in this order (prover.rs:107-108)."""
    p = FR_MODULUS[params.curve]
    r = rng.randrange(p)
    s = rng.randrange(p)
    return create_proof(params, circuit, r, s)
