"""PLONK AHP of the reference (TEST INFRASTRUCTURE): restatement of plonk/src/composer/*.rs, plonk/src/ahp/indexer/*.rs,
plonk/src/ahp/prover.rs, plonk/src/ahp/verifier.rs, plonk/src/ahp/mod.rs and plonk/src/utils.rs in Python integers.

* `Composer`            <- composer/mod.rs:13-77, arithmetic.rs:3-100, permutation.rs:20-118, synthesize.rs:69-132
* `index`               <- ahp/indexer/mod.rs:166-259 (11 interpolations on n, 13 coset FFTs on 4n)
* `prover_*`            <- ahp/prover.rs:69-247, indexer/arithmetic.rs:55-117, indexer/permutation.rs:81-170
* `linear_combinations` <- ahp/mod.rs:30-113 (+ the two key-specific ones), `lc_eval` <- ahp/evaluations.rs:24-48
* `verifier_equality_check` <- ahp/verifier.rs:105-150 -- the reference's own acceptance test for this layer
  (ahp/mod.rs:131-205 `fn ahp`): the product's GPU rounds are run through exactly that test.

Polynomials are coefficient lists (low degree first) with trailing zeros stripped like DensePolynomial."""
from .ntt import Domain


def strip(p):
    p = list(p)
    while p and p[-1] == 0:
        p.pop()
    return p


def poly_eval(p, x, mod):
    acc = 0
    for c in reversed(p):
        acc = (acc * x + c) % mod
    return acc


class Composer:
    """composer/mod.rs: rows of (w_0 = aux, w_1 = l, w_2 = r, w_3 = o) variables and selector values"""

    def __init__(self, p):
        self.p = p
        self.n = 0
        self.q = {k: [] for k in ("q_0", "q_1", "q_2", "q_3", "q_m", "q_c", "q_arith")}
        self.pi = []
        self.w = [[], [], [], []]
        self.variable_map = []          # Variable(i) -> [(wire column, row)] in insertion order (permutation.rs:20-58)
        self.assignment = []
        self.null_var = self.alloc_and_assign(0)

    def size(self):
        return self.n

    def alloc_and_assign(self, value):
        self.variable_map.append([])
        self.assignment.append(value % self.p)
        return len(self.assignment) - 1

    # arithmetic.rs:5-44
    def create_poly_gate(self, l, r, o, aux, q_m, q_c, pi):
        p = self.p
        index = self.n
        aux = aux if aux is not None else (self.null_var, 0)
        for col, var in enumerate((aux[0], l[0], r[0], o[0])):          # insert_gate: W0 = aux, W1 = l, W2 = r, W3 = o
            self.variable_map[var].append((col, index))
            self.w[col].append(var)
        self.pi.append(pi % p)
        for key, v in (("q_0", aux[1]), ("q_1", l[1]), ("q_2", r[1]), ("q_3", o[1]), ("q_m", q_m), ("q_c", q_c), ("q_arith", 1)):
            self.q[key].append(v % p)
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

    # permutation.rs:84-118
    def wire_permutation(self, n):
        perm = [[(col, i) for i in range(n)] for col in range(4)]
        for wires in self.variable_map:
            if len(wires) <= 1:
                continue
            for curr, (col, i) in enumerate(wires):
                nxt = len(wires) - 1 if curr == 0 else curr - 1
                perm[col][i] = wires[nxt]
        return perm

    # synthesize.rs:69-108
    def compose(self, fr, ks):
        dom = Domain(fr, self.n)
        n, p = dom.size, self.p
        roots, w = [], 1
        for _ in range(n):
            roots.append(w)
            w = w * dom.group_gen % p
        perm = self.wire_permutation(n)
        sel = {k: v + [0] * (n - self.n) for k, v in self.q.items()}
        for col in range(4):
            sel["sigma_%d" % col] = [roots[i] * ks[c] % p for (c, i) in perm[col]]
        return n, sel

    # synthesize.rs:114-132
    def synthesize(self, fr):
        n = Domain(fr, self.n).size
        return [[self.assignment[v] for v in col] + [0] * (n - self.n) for col in self.w]


SELECTOR_LABELS = ["q_0", "q_1", "q_2", "q_3", "q_m", "q_c", "q_arith", "sigma_0", "sigma_1", "sigma_2", "sigma_3"]


class Index:
    pass


def index(cs, fr, ks):
    """AHPForPLONK::index (indexer/mod.rs:166-259)"""
    n, sel = cs.compose(fr, ks)
    dn, d4 = Domain(fr, n), Domain(fr, 4 * n)
    p = fr.p
    idx = Index()
    idx.fr, idx.n, idx.ks, idx.domain_n, idx.domain_4n = fr, n, list(ks), dn, d4
    idx.polys, idx.evals_n, idx.evals_4n = {}, {}, {}
    for label in SELECTOR_LABELS:
        poly = strip(dn.ifft(sel[label]))
        idx.polys[label], idx.evals_n[label], idx.evals_4n[label] = poly, sel[label], d4.coset_fft(poly)
    v_poly = [p - 1] + [0] * (n - 1) + [1]                               # utils.rs:28-34
    idx.v_4n_inversed = [pow(v, -1, p) for v in d4.coset_fft(v_poly)]
    l1_poly = strip(dn.ifft([1] + [0] * (n - 1)))                        # utils.rs:41-45
    idx.l1_4n = d4.coset_fft(l1_poly)
    return idx


class ProverState:
    pass


def prover_init(cs, idx):
    """prover.rs:69-93"""
    ps = ProverState()
    ps.index = idx
    pi_n = cs.public_inputs() + [0] * (idx.n - cs.size())
    ps.pi_4n = idx.domain_4n.coset_fft(strip(idx.domain_n.ifft(pi_n)))
    return ps


def prover_first_round(ps, cs):
    """prover.rs:95-134 -> {label: poly}"""
    idx = ps.index
    ws = cs.synthesize(idx.fr)
    ps.w_n, ps.w_4n, oracles = [], [], {}
    for k, w in enumerate(ws):
        poly = strip(idx.domain_n.ifft(w))
        oracles["w_%d" % k] = poly
        ps.w_n.append(w)
        ps.w_4n.append(idx.domain_4n.coset_fft(poly))
    return oracles


def prover_second_round(ps, beta, gamma):
    """prover.rs:136-167 + PermutationKey::compute_z (indexer/permutation.rs:81-124)"""
    idx = ps.index
    p, n, ks = idx.fr.p, idx.n, idx.ks
    roots, w = [], 1
    for _ in range(n):
        roots.append(w)
        w = w * idx.domain_n.group_gen % p
    perms = []
    for i in range(n):
        num, den = 1, 1
        for k in range(4):
            num = num * ((ps.w_n[k][i] + ks[k] * beta % p * roots[i] + gamma) % p) % p
            den = den * ((ps.w_n[k][i] + beta * idx.evals_n["sigma_%d" % k][i] + gamma) % p) % p
        perms.append(num * pow(den, -1, p) % p)
    z, acc = [1], 1
    for i in range(n - 1):
        acc = acc * perms[i] % p
        z.append(acc)
    assert z[n - 1] * perms[n - 1] % p == 1
    z_poly = strip(idx.domain_n.ifft(z))
    ps.z_n, ps.z_4n, ps.beta, ps.gamma = z, idx.domain_4n.coset_fft(z_poly), beta, gamma
    return {"z": z_poly}


def prover_third_round(ps, alpha):
    """prover.rs:169-247, ArithmeticKey::compute_quotient (arithmetic.rs:55-117), PermutationKey::compute_quotient
    (permutation.rs:126-170), quad_split"""
    idx = ps.index
    p, n, ks = idx.fr.p, idx.n, idx.ks
    d4 = idx.domain_4n
    size = d4.size
    e4 = idx.evals_4n
    beta, gamma = ps.beta, ps.gamma
    linear_4n = d4.coset_fft([0, 1])
    alpha_2 = alpha * alpha % p
    t = []
    for i in range(size):
        w = [ps.w_4n[k][i] for k in range(4)]
        if e4["q_arith"][i] == 0:
            arith = 0
        else:
            arith = (e4["q_0"][i] * w[0] + e4["q_1"][i] * w[1] + e4["q_2"][i] * w[2] + e4["q_3"][i] * w[3]
                     + e4["q_m"][i] * w[1] % p * w[2] + e4["q_c"][i] + ps.pi_4n[i]) % p * e4["q_arith"][i] % p
        nxt = i % 4 if i // 4 == size // 4 - 1 else i + 4
        num, den = ps.z_4n[i], ps.z_4n[nxt]
        for k in range(4):
            num = num * ((w[k] + ks[k] * beta % p * linear_4n[i] + gamma) % p) % p
            den = den * ((w[k] + beta * e4["sigma_%d" % k][i] + gamma) % p) % p
        perm = ((num - den) * alpha + (ps.z_4n[i] - 1) * idx.l1_4n[i] % p * alpha_2) % p
        t.append((arith + perm) * idx.v_4n_inversed[i] % p)
    t_poly = strip(d4.coset_ifft(t))
    return {"t_%d" % k: strip(t_poly[k * n:(k + 1) * n]) for k in range(4)}, t_poly


def first_lagrange_at(idx, zeta):
    p, n = idx.fr.p, idx.n
    return (pow(zeta, n, p) - 1) * pow(n * (zeta - 1) % p, -1, p) % p


def linear_combinations(idx, beta, gamma, alpha, zeta, polys):
    """AHPForPLONK::construct_linear_combinations (ahp/mod.rs:30-113) -> {label: [(coeff, poly label)]}"""
    p, n, ks = idx.fr.p, idx.n, idx.ks
    lcs = {l: [(1, l)] for l in ("w_0", "w_1", "w_2", "w_3", "z", "sigma_0", "sigma_1", "sigma_2", "q_arith")}
    zn = pow(zeta, n, p)
    lcs["t"] = [(1, "t_0"), (zn, "t_1"), (zn * zn % p, "t_2"), (zn * zn % p * zn % p, "t_3")]
    ev = lambda label, x: poly_eval(polys[label], x, p)
    w = [ev("w_%d" % k, zeta) for k in range(4)]
    z_sh = ev("z", zeta * idx.domain_n.group_gen % p)
    s = [ev("sigma_%d" % k, zeta) for k in range(3)]
    qa = ev("q_arith", zeta)
    arith = [(qa * w[0] % p, "q_0"), (qa * w[1] % p, "q_1"), (qa * w[2] % p, "q_2"), (qa * w[3] % p, "q_3"),
             (qa * w[1] % p * w[2] % p, "q_m"), (qa, "q_c")]                                   # arithmetic.rs:24-41
    num = 1
    for k in range(4):
        num = num * ((w[k] + ks[k] * beta % p * zeta + gamma) % p) % p
    den = beta * z_sh % p
    for k in range(3):
        den = den * ((w[k] + beta * s[k] + gamma) % p) % p
    l1 = first_lagrange_at(idx, zeta)
    perm = [((num * alpha + l1 * alpha % p * alpha) % p, "z"), ((-den * alpha) % p, "sigma_3")]  # permutation.rs:36-79
    lcs["r"] = arith + perm
    return lcs


def lc_eval(lc, polys, point, mod):
    """EvaluationsProvider for a polynomial vector (ahp/evaluations.rs:24-48)"""
    return sum(c * poly_eval(polys[label], point, mod) for c, label in lc) % mod


def query_set(idx, zeta):
    """verifier_query_set (verifier.rs:81-103): {lc label: point}"""
    qs = {l: zeta for l in ("w_0", "w_1", "w_2", "w_3", "sigma_0", "sigma_1", "sigma_2", "q_arith", "t", "r")}
    qs["z"] = zeta * idx.domain_n.group_gen % idx.fr.p
    return qs


def verifier_equality_check(idx, beta, gamma, alpha, zeta, evals, public_inputs):
    """verifier.rs:105-150; evals = {lc label: value at its query point}"""
    p, n = idx.fr.p, idx.n
    v_zeta = (pow(zeta, n, p) - 1) % p
    pi_n = list(public_inputs) + [0] * (n - len(public_inputs))
    pi_zeta = poly_eval(strip(idx.domain_n.ifft(pi_n)), zeta, p)
    l1 = first_lagrange_at(idx, zeta)
    lhs = evals["t"] * v_zeta % p
    prod = evals["z"]
    for k in range(3):
        prod = prod * ((evals["w_%d" % k] + beta * evals["sigma_%d" % k] + gamma) % p) % p
    prod = prod * ((evals["w_3"] + gamma) % p) % p
    rhs = (evals["r"] + evals["q_arith"] * pi_zeta - prod * alpha - l1 * alpha % p * alpha) % p
    return lhs == rhs


def test_circuit(p):
    """plonk/src/lib.rs:318-358 `circuit()`"""
    cs = Composer(p)
    v1, v2, v3, v4, v6 = (cs.alloc_and_assign(x) for x in (1, 2, 3, 4, 6))
    cs.create_add_gate((v1, 1), (v2, 1), v3, None, 0, 0)
    cs.create_add_gate((v1, 1), (v3, 1), v4, None, 0, 0)
    cs.create_mul_gate(v2, v2, v4, None, 1, 0, 0)
    cs.create_mul_gate(v1, v2, v6, None, 2, 2, 0)
    cs.constrain_to_constant(v6, 6, 0)
    return cs


def run_ahp(cs, fr, ks, beta, gamma, alpha, zeta):
    """the reference's `fn ahp` test body (ahp/mod.rs:131-205) -> (accepted, evaluations, all polynomials)"""
    idx = index(cs, fr, ks)
    ps = prover_init(cs, idx)
    polys = dict(idx.polys)
    polys.update(prover_first_round(ps, cs))
    polys.update(prover_second_round(ps, beta, gamma))
    third, _ = prover_third_round(ps, alpha)
    polys.update(third)
    lcs = linear_combinations(idx, beta, gamma, alpha, zeta, polys)
    evals = {label: lc_eval(lcs[label], polys, point, fr.p) for label, point in query_set(idx, zeta).items()}
    return verifier_equality_check(idx, beta, gamma, alpha, zeta, evals, cs.public_inputs()), evals, polys
