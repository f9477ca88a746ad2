"""The reference's PLONK prover (`zkp-plonk`) on the B200 backend: same composer / indexer / prover
API, every vector of size n or 4n on the GPU through the C ABI.

    Composer                     plonk/src/composer/{mod,arithmetic,permutation,synthesize}.rs (gate recording: host)
    index(ctx, curve, cs, ks)    AHPForPLONK::index (ahp/indexer/mod.rs:166-259): 11 interpolations on n, 13 coset FFTs on 4n
    prover_init / first / second / third_round
                                 ahp/prover.rs:69-247; compute_z and both quotients: indexer/permutation.rs:81-170,
                                 indexer/arithmetic.rs:55-117
    keygen / prove               plonk/src/lib.rs:62-91,93-204: round oracles committed with PC::commit, challenges from
                                 the Fiat-Shamir generator (plonk/src/rng.rs: Blake2s + ChaCha20), evaluations of the
                                 query set, openings

GPU primitives used: zkb_ntt (ifft on n, coset fft / coset ifft on 4n), zkb_fr_vec_op (all pointwise loops),
zkb_fr_batch_inverse, zkb_fr_prefix_product (the accumulator z), zkb_fr_powers (domain elements), zkb_poly_div_linear
(evaluations), zkb_msm_batch through kzg10.pc_commit (commitments).

The reference commits and opens through the third-party `ark-poly-commit` crate (`MarlinKZG10`, `open_combinations`),
which is not vendored; its in-repo copy of the same scheme (marlin/src/pc) is what kzg10.py restates, and that is
what commits and opens here: one opening per query point over the linear combinations of ahp/mod.rs:30-113."""
import hashlib

import numpy as np

from . import _lib
from . import kzg10 as _kzg
from .backend import Context, is_dev
from .fs_rng import ChaChaRng, affine_to_bytes, fr_to_bytes
from .groth16 import FR_MODULUS
from .r1cs import ints_to_limbs, limbs_to_int

TWO_ADICITY = {_lib.BLS12_381: 32, _lib.BN254: 28}
FR_GENERATOR = {_lib.BLS12_381: 7, _lib.BN254: 5}
SELECTOR_LABELS = ["q_0", "q_1", "q_2", "q_3", "q_m", "q_c", "q_arith", "sigma_0", "sigma_1", "sigma_2", "sigma_3"]
ORACLE_LABELS = ["w_0", "w_1", "w_2", "w_3", "z", "t_0", "t_1", "t_2", "t_3"]          # AHPForPLONK::LABELS


class PlonkError(Exception):
    pass


class PolynomialDegreeTooLarge(PlonkError):
    pass


class CircuitTooLarge(PlonkError):
    pass


class Composer:
    """plonk/src/composer: a gate is one row (w_0 = aux, w_1 = l, w_2 = r, w_3 = o) with its selector values;
    values are plain ints mod r, converted to Montgomery limbs on the device when vectors are formed"""

    def __init__(self, modulus):
        self.p = modulus
        self.n = 0
        self.q = {k: [] for k in ("q_0", "q_1", "q_2", "q_3", "q_m", "q_c", "q_arith")}
        self.pi = []
        self.w = [[], [], [], []]
        self.variable_map = []
        self.assignment = []
        self.null_var = self.alloc_and_assign(0)

    def size(self):
        return self.n

    def alloc_and_assign(self, value):
        self.variable_map.append([])
        self.assignment.append(int(value) % self.p)
        self._limbs = None
        return len(self.assignment) - 1

    def witness_limbs(self):
        """(assignment as canonical uint64[n_vars, 4], the four wire columns as int64 index arrays), built once per
        circuit state: the reference's Composer holds field elements from `alloc_and_assign` on (composer/mod.rs:71-76),
        so this conversion is not part of its `prove` either"""
        if getattr(self, "_limbs", None) is None or self._limbs[2] != self.n:
            self._limbs = (ints_to_limbs(self.assignment), [np.asarray(c, dtype=np.int64) for c in self.w], self.n,
                           ints_to_limbs(self.pi) if self.pi else np.zeros((0, 4), dtype=np.uint64))
        return self._limbs

    def create_poly_gate(self, l, r, o, aux, q_m, q_c, pi):           # arithmetic.rs:5-44
        p = self.p
        index = self.n
        aux = aux if aux is not None else (self.null_var, 0)
        for col, var in enumerate((aux[0], l[0], r[0], o[0])):
            self.variable_map[var].append((col, index))
            self.w[col].append(var)
        self.pi.append(int(pi) % p)
        for key, v in (("q_0", aux[1]), ("q_1", l[1]), ("q_2", r[1]), ("q_3", o[1]), ("q_m", q_m), ("q_c", q_c), ("q_arith", 1)):
            self.q[key].append(int(v) % p)
        self.n += 1

    def constrain_to_constant(self, var, value, pi):
        self.create_poly_gate((var, 1), (var, 0), (var, 0), None, 0, -value, -pi)

    def assert_equal(self, l, r):
        self.create_poly_gate((l, 1), (r, -1), (self.null_var, 0), None, 0, 0, 0)

    def create_add_gate(self, l, r, o, aux, q_c, pi):
        self.create_poly_gate(l, r, (o, -1), aux, 0, q_c, pi)

    def create_mul_gate(self, l, r, o, aux, q_m, q_c, pi):
        self.create_poly_gate((l, 0), (r, 0), (o, -1), aux, q_m, q_c, pi)

    def public_inputs(self):
        return list(self.pi)

    def wire_permutation(self, n):
        """compute_wire_permutation (permutation.rs:84-118) as two index arrays per column: (target column, target row)"""
        cols = np.repeat(np.arange(4, dtype=np.int64)[:, None], n, axis=1)
        rows = np.repeat(np.arange(n, dtype=np.int64)[None, :], 4, axis=0)
        for wires in self.variable_map:
            if len(wires) <= 1:
                continue
            for curr, (col, i) in enumerate(wires):
                tc, ti = wires[len(wires) - 1 if curr == 0 else curr - 1]
                cols[col, i], rows[col, i] = tc, ti
        return cols, rows


def _log2_domain(curve, min_size):
    log = max(min_size - 1, 0).bit_length()
    if log > TWO_ADICITY[curve]:
        raise PolynomialDegreeTooLarge()
    return log


def _mont(ctx, curve, v):
    return ctx.fr_convert(curve, ints_to_limbs([int(v) % FR_MODULUS[curve]]), to_mont=True)[0]


def _mont_vec(ctx, curve, ints, size):
    """ints (canonical) padded with zeros to `size` -> Montgomery uint64[size, 4]"""
    out = np.zeros((size, 4), dtype=np.uint64)
    if ints:
        out[:len(ints)] = ctx.fr_convert(curve, ints_to_limbs(ints), to_mont=True)
    return out


def _group_gen(curve, log_n):
    p = FR_MODULUS[curve]
    root = pow(FR_GENERATOR[curve], (p - 1) >> TWO_ADICITY[curve], p)
    return pow(root, 1 << (TWO_ADICITY[curve] - log_n), p)


def _strip(poly):
    """DensePolynomial::from_coefficients_vec: trailing zero coefficients dropped (host array or resident tensor)"""
    from .marlin import contig, trim
    return contig(trim(poly))


class Index:
    """ahp/indexer/mod.rs:20-30: per selector the polynomial, its values on n and on the 4n coset"""

    def __init__(self):
        self.polys, self.evals_n, self.evals_4n = {}, {}, {}


def _interpolate(ctx, curve, evals_n, log_n):
    from .marlin import pad
    a = pad(evals_n, 1 << log_n)                                        # always a copy: the transform works in place
    ctx.ntt(curve, a, log_n, inverse=True)
    return a


def _coset_fft_4n(ctx, curve, poly, log_4n):
    from .marlin import pad
    a = pad(poly, 1 << log_4n)
    ctx.ntt(curve, a, log_4n, coset=True)
    return a


def index(ctx, curve, cs, ks, resident=False):
    """AHPForPLONK::index (ahp/indexer/mod.rs:166-259).  resident=True keeps every vector of the index and of the
    prover's round state in HBM (CUDA tensors, like marlin.Ops): the primitives then exchange device pointers."""
    from .marlin import Ops
    p = FR_MODULUS[curve]
    ops = Ops(ctx, curve, resident)
    log_n = _log2_domain(curve, cs.n)
    n = 1 << log_n
    log_4n = _log2_domain(curve, 4 * n)
    idx = Index()
    idx.resident = resident
    idx.curve, idx.n, idx.log_n, idx.log_4n, idx.ks = curve, n, log_n, log_4n, [int(k) % p for k in ks]
    idx.group_gen = _group_gen(curve, log_n)
    # Composer::compose (synthesize.rs:69-108): selectors padded with zeros, sigma = k_c * w^i at the permuted wire
    roots = ctx.fr_powers(curve, _mont(ctx, curve, idx.group_gen), n, device=ops.device)
    idx.roots = roots
    scaled = ops.cat([ctx.fr_vec_op(curve, Context.VEC_SCALE, roots, s=_mont(ctx, curve, k)) for k in idx.ks])   # [4 n, 4]
    cols, rows = cs.wire_permutation(n)
    sel = {k: ops.put(_mont_vec(ctx, curve, v, n)) for k, v in cs.q.items()}
    for c in range(4):
        sel["sigma_%d" % c] = ops.take(scaled, cols[c] * n + rows[c])
    for label in SELECTOR_LABELS:
        poly = _interpolate(ctx, curve, sel[label], log_n)
        idx.polys[label], idx.evals_n[label] = _strip(poly), sel[label]
        idx.evals_4n[label] = _coset_fft_4n(ctx, curve, idx.polys[label], log_4n)
    v_poly = np.zeros((n + 1, 4), dtype=np.uint64)                      # x^n - 1 (utils.rs:28-34)
    v_poly[0], v_poly[n] = _mont(ctx, curve, p - 1), _mont(ctx, curve, 1)
    idx.v_4n_inversed = ctx.fr_batch_inverse(curve, _coset_fft_4n(ctx, curve, ops.put(v_poly), log_4n))
    unit = np.zeros((n, 4), dtype=np.uint64)
    unit[0] = _mont(ctx, curve, 1)
    idx.l1_4n = _coset_fft_4n(ctx, curve, _strip(_interpolate(ctx, curve, ops.put(unit), log_n)), log_4n)   # utils.rs:41-45
    lin = np.zeros((2, 4), dtype=np.uint64)
    lin[1] = _mont(ctx, curve, 1)
    idx.linear_4n = _coset_fft_4n(ctx, curve, ops.put(lin), log_4n)     # coset_fft(&[0, 1]) of permutation.rs:139-142
    ops.release()                                                       # torch's current stream goes back to the caller
    return idx


class ProverState:
    pass


def prover_init(ctx, cs, idx):
    """ahp/prover.rs:69-93"""
    from .marlin import Ops
    ps = ProverState()
    ps.ctx, ps.index = ctx, idx
    ps.ops = Ops(ctx, idx.curve, idx.resident)      # resident: released by prove(); round-level callers call ps.ops.release()
    pi_canon = cs.witness_limbs()[3] if hasattr(cs, "witness_limbs") else ints_to_limbs(cs.public_inputs())
    pi_n = np.zeros((idx.n, 4), dtype=np.uint64)
    if len(pi_canon):
        pi_n[:len(pi_canon)] = ctx.fr_convert(idx.curve, pi_canon, to_mont=True)
    pi_poly = _strip(_interpolate(ctx, idx.curve, ps.ops.put(pi_n), idx.log_n))
    ps.pi_4n = _coset_fft_4n(ctx, idx.curve, pi_poly, idx.log_4n)
    return ps


def prover_first_round(ps, cs):
    """ahp/prover.rs:95-134 -> {label: polynomial}"""
    ctx, idx = ps.ctx, ps.index
    from .marlin import pad
    ps.w_n, ps.w_4n, oracles = [], [], {}
    # Composer::synthesize (synthesize.rs:114-132): the assignment table goes to the device once (one conversion per
    # variable, not per wire), the four wire columns are gathers from it, zero-padded to n
    canon, wires = cs.witness_limbs()[:2] if hasattr(cs, "witness_limbs") else (ints_to_limbs(cs.assignment), [np.asarray(c, dtype=np.int64) for c in cs.w])
    table = ps.ops.put(ctx.fr_convert(idx.curve, canon, to_mont=True))
    for k in range(4):
        w = pad(ps.ops.take(table, wires[k]), idx.n)
        poly = _strip(_interpolate(ctx, idx.curve, w, idx.log_n))
        oracles["w_%d" % k] = poly
        ps.w_n.append(w)
        ps.w_4n.append(_coset_fft_4n(ctx, idx.curve, poly, idx.log_4n))
    return oracles


def _factor_product(ctx, curve, ws, xs, scalars, gamma_m):
    """prod_k (w_k + s_k * x_k + gamma) elementwise"""
    V = Context
    acc = None
    for w, x, s in zip(ws, xs, scalars):
        f = ctx.fr_vec_op(curve, V.VEC_ADDC, ctx.fr_vec_op(curve, V.VEC_AXPY, w, x, s=s), s=gamma_m)
        acc = f if acc is None else ctx.fr_vec_op(curve, V.VEC_MUL, acc, f)
    return acc


def prover_second_round(ps, beta, gamma):
    """ahp/prover.rs:136-167 + PermutationKey::compute_z (indexer/permutation.rs:81-124); beta, gamma canonical ints"""
    ctx, idx = ps.ctx, ps.index
    curve, p = idx.curve, FR_MODULUS[idx.curve]
    gm = _mont(ctx, curve, gamma)
    num = _factor_product(ctx, curve, ps.w_n, [idx.roots] * 4, [_mont(ctx, curve, k * beta % p) for k in idx.ks], gm)
    den = _factor_product(ctx, curve, ps.w_n, [idx.evals_n["sigma_%d" % k] for k in range(4)], [_mont(ctx, curve, beta)] * 4, gm)
    perms = ctx.fr_vec_op(curve, Context.VEC_MUL, num, ctx.fr_batch_inverse(curve, den))
    z = ctx.fr_prefix_product(curve, perms)                             # z[0] = 1, z[i + 1] = z[i] * perms[i]
    from .marlin import contig, to_host
    closing = to_host(ctx.fr_vec_op(curve, Context.VEC_MUL, contig(z[-1:]), contig(perms[-1:])))
    if not np.array_equal(closing[0], _mont(ctx, curve, 1)):
        raise AssertionError("z[n - 1] * perms[n - 1] != 1: the copy constraints are not satisfied")   # permutation.rs:118
    z_poly = _strip(_interpolate(ctx, curve, z, idx.log_n))
    ps.z_n, ps.z_4n, ps.beta, ps.gamma = z, _coset_fft_4n(ctx, curve, z_poly, idx.log_4n), beta, gamma
    return {"z": z_poly}


def prover_third_round(ps, alpha):
    """ahp/prover.rs:169-247: the arithmetic quotient (indexer/arithmetic.rs:55-117), the permutation quotient
    (indexer/permutation.rs:126-170), division by the vanishing polynomial on the coset, coset iFFT, quad split"""
    ctx, idx = ps.ctx, ps.index
    curve, p, V = idx.curve, FR_MODULUS[idx.curve], Context
    e4, w = idx.evals_4n, ps.w_4n
    op = lambda o, a, b=None, s=None: ctx.fr_vec_op(curve, o, a, b, s)
    # (q_0 w_0 + q_1 w_1 + q_2 w_2 + q_3 w_3 + q_m w_1 w_2 + q_c + pi) * q_arith   (zero where q_arith is zero either way)
    acc = op(V.VEC_MUL, e4["q_0"], w[0])
    for k in (1, 2, 3):
        acc = op(V.VEC_ADD, acc, op(V.VEC_MUL, e4["q_%d" % k], w[k]))
    acc = op(V.VEC_ADD, acc, op(V.VEC_MUL, e4["q_m"], op(V.VEC_MUL, w[1], w[2])))
    acc = op(V.VEC_ADD, op(V.VEC_ADD, acc, e4["q_c"]), ps.pi_4n)
    t_arith = op(V.VEC_MUL, acc, e4["q_arith"])
    # permutation part
    beta, gamma = ps.beta, ps.gamma
    gm = _mont(ctx, curve, gamma)
    linear_4n = idx.linear_4n
    num = _factor_product(ctx, curve, w, [linear_4n] * 4, [_mont(ctx, curve, k * beta % p) for k in idx.ks], gm)
    den = _factor_product(ctx, curve, w, [e4["sigma_%d" % k] for k in range(4)], [_mont(ctx, curve, beta)] * 4, gm)
    # next = i + 4, wrapping to i % 4 in the last block: a rotation by four
    z_next = ps.z_4n.roll(-4, 0).contiguous() if is_dev(ps.z_4n) else np.ascontiguousarray(np.roll(ps.z_4n, -4, axis=0))
    diff = op(V.VEC_SUB, op(V.VEC_MUL, num, ps.z_4n), op(V.VEC_MUL, den, z_next))
    start = op(V.VEC_MUL, op(V.VEC_ADDC, ps.z_4n, s=_mont(ctx, curve, p - 1)), idx.l1_4n)     # (z - 1) * l1
    t_perm = op(V.VEC_AXPY, op(V.VEC_SCALE, diff, s=_mont(ctx, curve, alpha)), start, s=_mont(ctx, curve, alpha * alpha % p))
    t = op(V.VEC_MUL, op(V.VEC_ADD, t_arith, t_perm), idx.v_4n_inversed)
    ctx.ntt(curve, t, idx.log_4n, inverse=True, coset=True)             # coset_ifft
    t_poly = _strip(t)
    n = idx.n
    return {"t_%d" % k: _strip(t_poly[k * n:(k + 1) * n]) for k in range(4)}   # quad_split (:209-247)


# ---------------------------------------------------------------------------------------------------------
# linear combinations, query set, evaluations (ahp/mod.rs:30-113, ahp/verifier.rs:81-150, ahp/evaluations.rs)
# ---------------------------------------------------------------------------------------------------------
def _eval(ctx, curve, poly, point):
    """poly(point) as a canonical int; the zero polynomial evaluates to 0"""
    if len(poly) == 0:
        return 0
    rem = ctx.poly_eval(curve, poly, _mont(ctx, curve, point))
    return limbs_to_int(ctx.fr_convert(curve, rem.reshape(1, 4), to_mont=False)[0])


def first_lagrange_at(idx, zeta):
    p = FR_MODULUS[idx.curve]
    return (pow(zeta, idx.n, p) - 1) * pow(idx.n * (zeta - 1) % p, -1, p) % p


def construct_linear_combinations(ctx, idx, beta, gamma, alpha, zeta, polys):
    """-> {lc label: [(coefficient, polynomial label)]}; the scalars of `r` come from nine evaluations on the device"""
    curve, p, ks = idx.curve, FR_MODULUS[idx.curve], idx.ks
    lcs = {l: [(1, l)] for l in ("w_0", "w_1", "w_2", "w_3", "z", "sigma_0", "sigma_1", "sigma_2", "q_arith")}
    zn = pow(zeta, idx.n, p)
    lcs["t"] = [(1, "t_0"), (zn, "t_1"), (zn * zn % p, "t_2"), (zn * zn % p * zn % p, "t_3")]
    # the prover evaluates its polynomials (ahp/evaluations.rs:24-48), the verifier reads the proof's evaluations (:50-60)
    ev = (lambda label, x: polys[label]) if isinstance(polys.get("w_0"), int) else (lambda label, x: _eval(ctx, curve, polys[label], x))
    w = [ev("w_%d" % k, zeta) for k in range(4)]
    z_sh = ev("z", zeta * idx.group_gen % p)
    s = [ev("sigma_%d" % k, zeta) for k in range(3)]
    qa = ev("q_arith", zeta)
    arith = [(qa * w[0] % p, "q_0"), (qa * w[1] % p, "q_1"), (qa * w[2] % p, "q_2"), (qa * w[3] % p, "q_3"),
             (qa * w[1] % p * w[2] % p, "q_m"), (qa, "q_c")]
    num = 1
    for k in range(4):
        num = num * ((w[k] + ks[k] * beta % p * zeta + gamma) % p) % p
    den = beta * z_sh % p
    for k in range(3):
        den = den * ((w[k] + beta * s[k] + gamma) % p) % p
    l1 = first_lagrange_at(idx, zeta)
    lcs["r"] = arith + [((num * alpha + l1 * alpha % p * alpha) % p, "z"), ((-den * alpha) % p, "sigma_3")]
    return lcs


def lc_polynomial(ctx, curve, lc, polys):
    """the polynomial of a linear combination (EvaluationsProvider for Vec<LabeledPolynomial>, ahp/evaluations.rs:24-48)"""
    terms = [(c, polys[label]) for c, label in lc if len(polys[label])]
    if not terms:
        return polys[lc[0][1]][:0]
    coeffs = ctx.fr_convert(curve, ints_to_limbs([c for c, _ in terms]), to_mont=True)
    return ctx.poly_lincomb(curve, [q for _, q in terms], coeffs)


def verifier_query_set(idx, zeta):
    """ahp/verifier.rs:81-103 -> {lc label: (point label, point)}"""
    p = FR_MODULUS[idx.curve]
    qs = {l: ("zeta", zeta) for l in ("w_0", "w_1", "w_2", "w_3", "sigma_0", "sigma_1", "sigma_2", "q_arith", "t", "r")}
    qs["z"] = ("shifted_zeta", zeta * idx.group_gen % p)
    return qs


def verifier_equality_check(ctx, idx, beta, gamma, alpha, zeta, evals, public_inputs):
    """ahp/verifier.rs:105-150 (the verifier's side, here so that a proof can be self-checked on the same primitives)"""
    curve, p, n = idx.curve, FR_MODULUS[idx.curve], idx.n
    v_zeta = (pow(zeta, n, p) - 1) % p
    from .marlin import Ops
    ops = Ops(ctx, curve, idx.resident)
    try:
        pi_poly = _strip(_interpolate(ctx, curve, ops.put(_mont_vec(ctx, curve, list(public_inputs), n)), idx.log_n))
        pi_zeta = _eval(ctx, curve, pi_poly, zeta)
    finally:
        ops.release()
    l1 = first_lagrange_at(idx, zeta)
    prod = evals["z"]
    for k in range(3):
        prod = prod * ((evals["w_%d" % k] + beta * evals["sigma_%d" % k] + gamma) % p) % p
    prod = prod * ((evals["w_3"] + gamma) % p) % p
    rhs = (evals["r"] + evals["q_arith"] * pi_zeta - prod * alpha - l1 * alpha % p * alpha) % p
    return evals["t"] * v_zeta % p == rhs


# ---------------------------------------------------------------------------------------------------------
# keygen / prove (plonk/src/lib.rs:62-204)
# ---------------------------------------------------------------------------------------------------------
class FiatShamirRng:
    """plonk/src/rng.rs with D = Blake2s (plonk/src/lib.rs:306): seed = H(material), absorb: seed = H(material || seed),
    ChaCha20 keyed with the 32-byte digest"""

    def __init__(self, seed_material, curve):
        self.p = FR_MODULUS[curve]
        self.shave = 256 - self.p.bit_length()
        self.rinv = pow(1 << 256, -1, self.p)
        self.seed = hashlib.blake2s(bytes(seed_material)).digest()
        self.r = ChaChaRng(self.seed)

    def absorb(self, material):
        self.seed = hashlib.blake2s(bytes(material) + self.seed).digest()
        self.r = ChaChaRng(self.seed)

    def rand_fr(self):
        """`F::rand(rng)`: four u64 draws, top bits shaved, rejected when >= p, taken as the Montgomery residue"""
        while True:
            limbs = [self.r.next_u64() for _ in range(4)]
            limbs[3] &= ((1 << 64) - 1) >> self.shave
            v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
            if v < self.p:
                return v * self.rinv % self.p


def _labeled(polys, labels):
    return [_kzg.LabeledPolynomial(l, polys[l]) for l in labels]


def _pc_commit(ck, labeled):
    """PC::commit for PLONK's oracles.  ark-poly-commit's KZG10 accepts constant and zero polynomials (t_3 is zero for
    small circuits, a selector may be constant); the in-repo scheme (marlin/src/pc/kzg10.rs:175-183) rejects degree 0, so
    those are committed here as c * G / the identity and only the rest goes through kzg10.pc_commit."""
    ctx, curve = ck.ctx, ck.curve
    regular = [P for P in labeled if _kzg._degree(P.coeffs) >= 1]
    done = iter(_kzg.pc_commit(ck, regular, None)[0]) if regular else iter(())
    comms = []
    for P in labeled:
        if _kzg._degree(P.coeffs) >= 1:
            comms.append(next(done))
        elif len(P.coeffs) == 0:
            from .backend import point_words
            comms.append(((np.zeros(point_words(curve, _lib.G1), dtype=np.uint64), True), None))
        else:
            comms.append((ck.msm_g(P.coeffs[:1], 0), None))
    return comms, [(_kzg.Randomness(), None)] * len(labeled)


def _comms_to_bytes(curve, comms):
    """to_bytes![Vec<LabeledCommitment<marlin_pc::Commitment>>] (layout recalled from ark-poly-commit 0.2, the crate is
    not vendored: comm, shifted_comm option flag) -- only the prover and verifier of THIS package need to agree on it"""
    return b"".join(affine_to_bytes(curve, c) + b"\x00" for c, _ in comms)


class ProverKey:
    pass


def keygen(ctx, srs, cs, ks, resident=False):
    """Plonk::keygen (lib.rs:62-91): index, PC::trim(srs, index.size()), commitments to the 11 index polynomials.
    srs: marlin.UniversalParams (marlin.universal_setup) -> (ProverKey, verifier-key dict)"""
    from . import marlin as _marlin
    idx = index(ctx, srs.curve, cs, ks, resident=resident)
    if srs.max_degree() < idx.n:
        raise CircuitTooLarge()
    ck, rk = _marlin.pc_trim(ctx, srs, idx.n)
    comms, rands = _pc_commit(ck, _labeled(idx.polys, SELECTOR_LABELS))
    pk = ProverKey()
    pk.index, pk.ck, pk.rands, pk.comms = idx, ck, rands, comms
    vk = {"comms": comms, "labels": list(SELECTOR_LABELS), "rk": rk, "n": idx.n, "ks": idx.ks}
    pk.vk = vk
    return pk, vk


class Proof:
    """plonk/src/data_structures.rs:41-45: commitments per round, evaluations sorted by label, openings per query point"""

    def __init__(self, commitments, evaluations, openings):
        self.commitments, self.evaluations, self.openings = commitments, evaluations, openings


def prove(ctx, pk, cs, fs_rng=None):
    """Plonk::prove (lib.rs:93-204).  -> (Proof, challenges dict); `fs_rng` defaults to the Blake2s generator seeded
    with to_bytes![PROTOCOL_NAME, public_inputs]"""
    idx, ck, curve = pk.index, pk.ck, pk.index.curve
    p = FR_MODULUS[curve]
    if fs_rng is None:
        pi_bytes = (cs.witness_limbs()[3].tobytes() if hasattr(cs, "witness_limbs")          # canonical LE, 32 bytes each
                    else b"".join(fr_to_bytes(x) for x in cs.public_inputs()))
        fs_rng = FiatShamirRng(b"PLONK" + pi_bytes, curve)
    ps = prover_init(ctx, cs, idx)
    try:
        polys = dict(idx.polys)

        first = prover_first_round(ps, cs)
        first_comms, first_rands = _pc_commit(ck, _labeled(first, ORACLE_LABELS[:4]))
        fs_rng.absorb(_comms_to_bytes(curve, first_comms))
        beta, gamma = fs_rng.rand_fr(), fs_rng.rand_fr()                    # verifier_first_round
        polys.update(first)

        second = prover_second_round(ps, beta, gamma)
        second_comms, second_rands = _pc_commit(ck, _labeled(second, ["z"]))
        fs_rng.absorb(_comms_to_bytes(curve, second_comms))
        alpha = fs_rng.rand_fr()                                            # verifier_second_round
        polys.update(second)

        third = prover_third_round(ps, alpha)
        third_comms, third_rands = _pc_commit(ck, _labeled(third, ORACLE_LABELS[5:]))
        fs_rng.absorb(_comms_to_bytes(curve, third_comms))
        zeta = fs_rng.rand_fr()                                             # verifier_third_round
        polys.update(third)

        qs = verifier_query_set(idx, zeta)
        lcs = construct_linear_combinations(ctx, idx, beta, gamma, alpha, zeta, polys)
        lc_polys = {label: lc_polynomial(ctx, curve, lcs[label], polys) for label in qs}
        evals = {label: _eval(ctx, curve, lc_polys[label], point) for label, (_, point) in qs.items()}
        evaluations = [evals[label] for label in sorted(evals)]             # evals.sort_by label (lib.rs:171-173)
        fs_rng.absorb(b"".join(fr_to_bytes(e) for e in evaluations))
        epsilon = fs_rng.rand_fr()

        # one opening per query point over the linear-combination polynomials, combined with powers of epsilon
        # (the role of PC::open_combinations, lib.rs:186-197)
        openings = {}
        for point_label in ("shifted_zeta", "zeta"):
            group = [label for label in sorted(qs) if qs[label][0] == point_label]
            point = qs[group[0]][1]
            labeled = [_kzg.LabeledPolynomial(label, lc_polys[label]) for label in group]
            rands = [(_kzg.Randomness(), None)] * len(group)
            openings[point_label] = _kzg.pc_open(ck, labeled, _mont(ctx, curve, point), epsilon, rands)
    finally:
        ps.ops.release()
    proof = Proof([first_comms, second_comms, third_comms], evaluations, openings)
    return proof, {"beta": beta, "gamma": gamma, "alpha": alpha, "zeta": zeta, "epsilon": epsilon, "evals": evals}


class _VerifierInfo:
    """what the verifier knows of the index (ahp/verifier.rs:19-35: the domain and the coset representatives)"""

    def __init__(self, curve, n, ks):
        self.curve, self.n, self.ks, self.resident = curve, n, list(ks), False
        self.log_n = n.bit_length() - 1
        self.group_gen = _group_gen(curve, self.log_n)


def verify(ctx, vk, public_inputs, proof, fs_rng=None):
    """Plonk::verify (lib.rs:206-290): challenges re-derived from the transcript, the equality check on the evaluations,
    then the opening checks -- the role of PC::check_combinations (:277-286): the commitment of every linear combination
    is the same combination of the labeled commitments, one KZG check per query point with the combinations' evaluations
    batched by powers of epsilon -- with all pairings in one device call (kzg10._check_many)."""
    rk = vk["rk"]
    curve = rk.curve
    p = FR_MODULUS[curve]
    info = _VerifierInfo(curve, vk["n"], vk["ks"])
    public_inputs = [int(x) % p for x in public_inputs]
    if fs_rng is None:
        fs_rng = FiatShamirRng(b"PLONK" + b"".join(fr_to_bytes(x) for x in public_inputs), curve)
    first, second, third = proof.commitments
    fs_rng.absorb(_comms_to_bytes(curve, first))
    beta, gamma = fs_rng.rand_fr(), fs_rng.rand_fr()
    fs_rng.absorb(_comms_to_bytes(curve, second))
    alpha = fs_rng.rand_fr()
    fs_rng.absorb(_comms_to_bytes(curve, third))
    zeta = fs_rng.rand_fr()
    qs = verifier_query_set(info, zeta)
    fs_rng.absorb(b"".join(fr_to_bytes(e) for e in proof.evaluations))
    epsilon = fs_rng.rand_fr()
    evals = dict(zip(sorted(qs), proof.evaluations))                     # evaluation_labels.sort_by label (lib.rs:232-244)
    if not verifier_equality_check(ctx, info, beta, gamma, alpha, zeta, evals, public_inputs):
        return False
    labels = list(vk["labels"]) + list(ORACLE_LABELS)
    comms = dict(zip(labels, [c for c, _ in list(vk["comms"]) + list(first) + list(second) + list(third)]))
    lcs = construct_linear_combinations(ctx, info, beta, gamma, alpha, zeta, evals)
    items = []
    for point_label in ("shifted_zeta", "zeta"):
        group = [label for label in sorted(qs) if qs[label][0] == point_label]
        point = qs[group[0]][1]
        terms, acc_v, ch = [], 0, 1
        for label in group:                                              # accumulate_commitments_and_values over the combinations
            terms += [(comms[poly_label], ch * coeff % p) for coeff, poly_label in lcs[label]]
            acc_v = (acc_v + ch * evals[label]) % p
            ch = ch * epsilon % p * epsilon % p
        terms.append((rk.g, -acc_v))
        items.append((terms, point, proof.openings[point_label]))
    return all(_kzg._check_many(ctx, rk, items))

