"""The reference's Groth16 setup (`zkp_groth16::generate_random_parameters` / `generate_parameters`,
groth16/src/generator.rs:19-34,135-286) on the B200 backend -- the step before the prove path.

Same sequence as the reference: synthesise the circuit into a KeypairAssembly (generator.rs:38-132), sample t outside
the domain, `R1CStoQAP::instance_map_with_evaluation` (r1cs_to_qap.rs:58-110), gamma_abc and l (:187-199), then the five
queries as fixed-base multiples of the generators with batch normalisation (:205-274).  Everything of size O(circuit)
runs on the GPU through the C ABI:

    Lagrange coefficients at t   zkb_fr_powers, zkb_fr_vec_op, zkb_fr_batch_inverse   (ark-poly evaluate_all_lagrange_coefficients)
    a, b, c = M^T u              zkb_spmv on the transposed matrices                  (r1cs_to_qap.rs:82-107)
    gamma_abc, l, h scalars      zkb_fr_vec_op, zkb_fr_powers                         (generator.rs:187-199,235-242)
    queries                      zkb_fixed_base_mul (result-identical to FixedBaseMSM + batch_normalization: canonical affine)

The toxic waste and the two generators are explicit arguments of `generate_parameters`; `generate_random_parameters`
draws them from `rng` in the reference's order (alpha, beta, gamma, delta, t, g1, g2 -- arkworks' own byte stream is only
reproducible from Rust, so the generators are uniform multiples of the standard ones).
"""
import numpy as np

from . import _lib
from .backend import Context, CsrMatrix
from .groth16 import FR_MODULUS, Parameters
from .r1cs import PolynomialDegreeTooLarge, SynthesisError, Variable, ints_to_limbs, limbs_to_int
from .synth import generator_mont

TWO_ADICITY = {_lib.BLS12_381: 32, _lib.BN254: 28}
# Fr::multiplicative_generator() (7 / 5): the 2-adic root of unity is g^((p - 1) / 2^s)
FR_GENERATOR = {_lib.BLS12_381: 7, _lib.BN254: 5}


class UnexpectedIdentity(SynthesisError):
    pass


class KeypairAssembly:
    """generator.rs:38-132: records the constraint rows; assignments are never evaluated."""

    def __init__(self, modulus):
        self.p = modulus
        self.num_inputs = self.num_aux = 0
        self._rows = {"a": ([0], [], []), "b": ([0], [], []), "c": ([0], [], [])}

    def alloc(self, f=None):
        self.num_aux += 1
        return Variable("aux", self.num_aux - 1)

    def alloc_input(self, f=None):
        self.num_inputs += 1
        return Variable("in", self.num_inputs - 1)

    def enforce(self, a, b, c):
        for lc, key in ((a, "a"), (b, "b"), (c, "c")):
            ptr, vars_, coeffs = self._rows[key]
            for coeff, var in lc:
                vars_.append(var)
                coeffs.append(int(coeff) % self.p)
            ptr.append(len(vars_))

    @property
    def num_constraints(self):
        return len(self._rows["a"][0]) - 1

    def transposed(self, ctx, curve, which):
        """CSR of M^T (one row per variable: Input(i) -> i, Aux(i) -> num_inputs + i) with Montgomery coefficients"""
        ptr, vars_, coeffs = self._rows[which]
        nvars = self.num_inputs + self.num_aux
        cols = np.fromiter((v[1] if v[0] == "in" else self.num_inputs + v[1] for v in vars_), dtype=np.int64, count=len(vars_))
        rows = np.repeat(np.arange(len(ptr) - 1, dtype=np.int64), np.diff(np.asarray(ptr, dtype=np.int64)))
        order = np.argsort(cols, kind="stable")
        t_ptr = np.zeros(nvars + 1, dtype=np.uint32)
        np.cumsum(np.bincount(cols, minlength=nvars), out=t_ptr[1:])
        co = ctx.fr_convert(curve, ints_to_limbs(coeffs), to_mont=True) if coeffs else np.zeros((0, 4), dtype=np.uint64)
        return CsrMatrix(t_ptr, rows[order].astype(np.uint32), np.ascontiguousarray(co[order]))


class VerifyKey:
    """groth16/src/lib.rs:59-66"""

    def __init__(self, alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1):
        self.alpha_g1, self.beta_g2, self.gamma_g2, self.delta_g2 = alpha_g1, beta_g2, gamma_g2, delta_g2
        self.gamma_abc_g1 = gamma_abc_g1


class ParametersData:
    """groth16/src/lib.rs:81-91 as host arrays in the layout of include/zkb.h: point arrays are (xy uint64[n, words],
    inf uint8[n]), single points (xy uint64[words], is_identity)."""

    def __init__(self, curve, vk, beta_g1, delta_g1, a_query, b_g1_query, b_g2_query, h_query, l_query):
        self.curve, self.vk, self.beta_g1, self.delta_g1 = curve, vk, beta_g1, delta_g1
        self.a_query, self.b_g1_query, self.b_g2_query, self.h_query, self.l_query = a_query, b_g1_query, b_g2_query, h_query, l_query

    def upload(self, ctx, shard=None):
        """-> groth16.Parameters: the proving key resident in HBM (window tables built on the device)"""
        return Parameters(ctx, self.curve, self.a_query, self.b_g1_query, self.b_g2_query, self.h_query, self.l_query,
                          self.vk.alpha_g1[0], self.beta_g1[0], self.delta_g1[0], self.vk.beta_g2[0], self.vk.delta_g2[0],
                          shard=shard)


def _mont(ctx, curve, v):
    return ctx.fr_convert(curve, ints_to_limbs([v % FR_MODULUS[curve]]), to_mont=True)[0]


def lagrange_coefficients(ctx, curve, log_m, t):
    """EvaluationDomain::evaluate_all_lagrange_coefficients(t) for t outside the domain (ark-poly 0.2, used at
    r1cs_to_qap.rs:74): u_i = (t^m - 1) / m * w^i / (t - w^i), Montgomery uint64[m, 4]"""
    p = FR_MODULUS[curve]
    m = 1 << log_m
    root = pow(FR_GENERATOR[curve], (p - 1) >> TWO_ADICITY[curve], p)
    w = pow(root, 1 << (TWO_ADICITY[curve] - log_m), p)
    zt = (pow(t, m, p) - 1) % p
    omegas = ctx.fr_powers(curve, _mont(ctx, curve, w), m)
    denom = ctx.fr_vec_op(curve, Context.VEC_RSUB, omegas, s=_mont(ctx, curve, t))                   # t - w^i
    numer = ctx.fr_vec_op(curve, Context.VEC_SCALE, omegas, s=_mont(ctx, curve, zt * pow(m, -1, p)))  # (z / m) w^i
    return ctx.fr_vec_op(curve, Context.VEC_MUL, numer, ctx.fr_batch_inverse(curve, denom)), zt


def generate_parameters(ctx, curve, circuit, alpha, beta, gamma, delta, t, g1_generator=None, g2_generator=None):
    """generator.rs:135-286 with the rng draws (t, generators) as arguments.  Generators: one affine point each in the
    ABI layout (default: the curve's standard generators)."""
    p = FR_MODULUS[curve]
    asm = KeypairAssembly(p)
    asm.alloc_input()                                       # the "one" input (:158)
    circuit.generate_constraints(asm)                       # :161
    ncons, n_inputs, n_aux = asm.num_constraints, asm.num_inputs, asm.num_aux
    domain_size = ncons + (n_inputs - 1) + 1                # :165
    log_m = max(domain_size - 1, 0).bit_length()
    if log_m > TWO_ADICITY[curve]:
        raise PolynomialDegreeTooLarge()
    m_raw = 1 << log_m
    if pow(t, m_raw, p) == 1:
        raise ValueError("t lies inside the evaluation domain")
    g1 = generator_mont(curve, _lib.G1) if g1_generator is None else np.asarray(g1_generator, dtype=np.uint64)
    g2 = generator_mont(curve, _lib.G2) if g2_generator is None else np.asarray(g2_generator, dtype=np.uint64)

    # ---- instance_map_with_evaluation (r1cs_to_qap.rs:58-110)
    u, zt = lagrange_coefficients(ctx, curve, log_m, t)
    nvars = n_inputs + n_aux
    abc = []
    for which in "abc":
        mt = asm.transposed(ctx, curve, which)
        abc.append(ctx.spmv(curve, mt, u[:max(ncons, 1)]) if mt.nnz else np.zeros((nvars, 4), dtype=np.uint64))
    a, b, c = abc
    # a[i] += u[num_constraints + i] for the inputs (:78-80): a one-entry-per-row SpMV would do; the slice add is it
    a[:n_inputs] = ctx.fr_vec_op(curve, Context.VEC_ADD, np.ascontiguousarray(a[:n_inputs]),
                                 np.ascontiguousarray(u[ncons:ncons + n_inputs]))
    if gamma % p == 0 or delta % p == 0:
        raise UnexpectedIdentity()
    gamma_inv, delta_inv = pow(gamma, -1, p), pow(delta, -1, p)
    # beta * a + alpha * b + c (:187-199)
    comb = ctx.fr_vec_op(curve, Context.VEC_SCALE, a, s=_mont(ctx, curve, beta))
    comb = ctx.fr_vec_op(curve, Context.VEC_AXPY, comb, b, s=_mont(ctx, curve, alpha))
    comb = ctx.fr_vec_op(curve, Context.VEC_ADD, comb, c)
    gamma_abc = ctx.fr_vec_op(curve, Context.VEC_SCALE, np.ascontiguousarray(comb[:n_inputs]), s=_mont(ctx, curve, gamma_inv))
    l = ctx.fr_vec_op(curve, Context.VEC_SCALE, np.ascontiguousarray(comb[n_inputs:]), s=_mont(ctx, curve, delta_inv))   # :247
    h = ctx.fr_powers(curve, _mont(ctx, curve, t), m_raw - 1, scale_mont=_mont(ctx, curve, zt * delta_inv))              # :235-242

    # ---- the queries: fixed-base multiples, canonical affine (:205-274)
    canon = lambda v: ctx.fr_convert(curve, v, to_mont=False)
    mul1 = lambda v: ctx.fixed_base_mul(curve, _lib.G1, g1, canon(v))
    mul2 = lambda v: ctx.fixed_base_mul(curve, _lib.G2, g2, canon(v))
    singles = ints_to_limbs([alpha % p, beta % p, delta % p, gamma % p])
    s1, s1inf = ctx.fixed_base_mul(curve, _lib.G1, g1, singles[:3])
    s2, s2inf = ctx.fixed_base_mul(curve, _lib.G2, g2, singles[1:])
    pt = lambda arr, inf, i: (arr[i], bool(inf[i]))
    b_canon = canon(b)
    vk = VerifyKey(pt(s1, s1inf, 0), pt(s2, s2inf, 0), pt(s2, s2inf, 2), pt(s2, s2inf, 1), mul1(gamma_abc))
    return ParametersData(curve, vk, pt(s1, s1inf, 1), pt(s1, s1inf, 2), mul1(a),
                          ctx.fixed_base_mul(curve, _lib.G1, g1, b_canon), ctx.fixed_base_mul(curve, _lib.G2, g2, b_canon),
                          mul1(h), mul1(l))


def generate_random_parameters(ctx, curve, circuit, rng):
    """generator.rs:19-34: alpha, beta, gamma, delta <- Fr::rand(rng), then (inside generate_parameters) t outside the
    domain (:167) and the two generators (:201-202), in this order.  `rng` needs randrange()."""
    p = FR_MODULUS[curve]
    alpha, beta, gamma, delta = (rng.randrange(p) for _ in range(4))
    probe = KeypairAssembly(p)
    probe.alloc_input()
    circuit.generate_constraints(probe)
    m = 1 << max(probe.num_constraints + probe.num_inputs - 1, 0).bit_length()
    while True:                                             # sample_element_outside_domain
        t = rng.randrange(p)
        if pow(t, m, p) != 1:
            break
    k1, k2 = rng.randrange(1, p), rng.randrange(1, p)
    g1, _ = ctx.fixed_base_mul(curve, _lib.G1, generator_mont(curve, _lib.G1), ints_to_limbs([k1]))
    g2, _ = ctx.fixed_base_mul(curve, _lib.G2, generator_mont(curve, _lib.G2), ints_to_limbs([k2]))
    return generate_parameters(ctx, curve, circuit, alpha, beta, gamma, delta, t, g1[0], g2[0])
