"""Marlin AHP: restatement of the indexer and the three prover rounds (TEST INFRASTRUCTURE).

Follows marlin/src/ahp/{constraint_systems,indexer,arithmetic,prover,verifier}.rs of the reference:

    index()                 indexer.rs:71-116, arithmetic.rs:97-172 (compose_matrix_polynomials),
                            constraint_systems.rs:9-31,83-133 (square, balance, sort)
    prover_init()           prover.rs:86-147
    prover_first_round()    prover.rs:150-222      w, z_a, z_b, mask
    prover_second_round()   prover.rs:230-321      t, g_1, h_1
    prover_third_round()    prover.rs:331-427      g_2, h_2
    verifier_equality_check()  verifier.rs:128-209 (the AHP acceptance test: outer and inner sumcheck)

Everything the reference draws from an rng (blinding coefficients, the mask polynomial, the verifier's
challenges -- Fiat-Shamir bytes are only reproducible from the Rust host)
is an explicit argument.  Polynomials are coefficient lists (low degree first) of canonical ints;
`ark-poly` behaviours relied on (un-vendored, recalled): EvaluationDomain::new = next power of two,
reindex_by_subdomain, divide_by_vanishing_poly, DensePolynomial trimming of high zero coefficients.
"""
from .fields import FR
from .ntt import Domain


# ---- polynomial helpers --------------------------------------------------------------------------
def trim(p):
    p = list(p)
    while p and p[-1] == 0:
        p.pop()
    return p


def poly_eval(p, x, mod):
    acc = 0
    for c in reversed(p):
        acc = (acc * x + c) % mod
    return acc


def poly_add(a, b, mod):
    n = max(len(a), len(b))
    return trim([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % mod for i in range(n)])


def poly_sub(a, b, mod):
    n = max(len(a), len(b))
    return trim([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % mod for i in range(n)])


def poly_mul(a, b, fr):
    a, b = trim(a), trim(b)
    if not a or not b:
        return []
    d = Domain(fr, len(a) + len(b) - 1)
    ea, eb = d.fft(a), d.fft(b)
    return trim(d.ifft([x * y % fr.p for x, y in zip(ea, eb)]))


def divide_by_vanishing_poly(p, n, mod):
    """ark-poly DensePolynomial::divide_by_vanishing_poly for x^n - 1 -> (quotient, remainder)"""
    p = list(p)
    if len(p) < n:
        return [], trim(p)
    q = p[n:]
    for i in range(1, len(p) // n):
        for k, c in enumerate(p[n * (i + 1):]):
            q[k] = (q[k] + c) % mod
    r = p[:n]
    for k, c in enumerate(q[:n]):
        r[k] = (r[k] + c) % mod
    return trim(q), trim(r)


def mul_by_vanishing_poly(p, n, mod):
    out = [0] * (len(p) + n)
    for i, c in enumerate(p):
        out[i + n] = c
        out[i] = (out[i] - c) % mod
    return trim(out)


def reindex_by_subdomain(h_size, x_size, index):
    """ark-poly EvaluationDomain::reindex_by_subdomain"""
    period = h_size // x_size
    if index < x_size:
        return index * period
    i = index - x_size
    x = period - 1
    return i + (i // x) + 1


def batch_evals(domain, x):
    """arithmetic.rs:28-34: u_H(x, w^i) = v_H(x) / (x - w^i)"""
    p = domain.p
    v = domain.vanishing_at(x)
    out, w = [], 1
    for _ in range(domain.size):
        out.append(v * pow((x - w) % p, -1, p) % p)
        w = w * domain.group_gen % p
    return out


def bivariate_eval(domain, x, y):
    """arithmetic.rs:19-26"""
    p = domain.p
    if x != y:
        return (domain.vanishing_at(x) - domain.vanishing_at(y)) * pow((x - y) % p, -1, p) % p
    return domain.size * pow(x, domain.size - 1, p) % p


# ---- constraint systems -----------------------------------------------------------------------------
class MarlinCS:
    """IndexerConstraintSystem + ProverConstraintSystem in one (constraint_systems.rs:33-284):
    rows as [(coeff, ('in'|'aux', i))], input 0 = ONE."""

    def __init__(self, mod):
        self.p = mod
        self.input = [1]
        self.witness = []
        self.a, self.b, self.c = [], [], []

    def alloc(self, v):
        self.witness.append(v % self.p)
        return ("aux", len(self.witness) - 1)

    def alloc_input(self, v):
        self.input.append(v % self.p)
        return ("in", len(self.input) - 1)

    def enforce(self, a, b, c):
        for lc, dst in ((a, self.a), (b, self.b), (c, self.c)):
            dst.append([(co % self.p, v) for co, v in lc])

    def make_matrices_square(self):
        """constraint_systems.rs:9-31"""
        nv = len(self.input) + len(self.witness)
        nc = len(self.a)
        if nv < nc:
            for _ in range(nc - nv):
                self.alloc(1)
        else:
            for _ in range(nv - nc):
                self.enforce([], [], [])

    def matrices(self):
        """process_matrices (constraint_systems.rs:83-99): reindex, balance a/b, sort columns"""
        ni = len(self.input)
        re = lambda m: [[(co, v[1] if v[0] == "in" else ni + v[1]) for co, v in row] for row in m]
        a, b, c = re(self.a), re(self.b), re(self.c)
        da, db = sum(map(len, a)), sum(map(len, b))
        denser = da > db
        for i in range(len(a)):                           # balance_matrices :100-114
            if denser:
                ra, rb = len(a[i]), len(b[i])
                a[i], b[i] = b[i], a[i]
                da += rb - ra
                db += ra - rb
                denser = da > db
        srt = lambda m: [sorted(row, key=lambda t: t[1]) for row in m]      # stable, like slice::sort_by
        return srt(a), srt(b), srt(c)


def compose_matrix_polynomials(matrix, dx, dh, dk, db):
    """arithmetic.rs:97-172"""
    p = dh.p
    h_elems, w = [], 1
    for _ in range(dh.size):
        h_elems.append(w)
        w = w * dh.group_gen % p
    diag = [dh.size * u % p for u in h_elems]
    diag[1:] = diag[1:][::-1]
    row_vec, col_vec, val_vec = [], [], []
    for i, row in enumerate(matrix):
        for v, j in row:
            j = reindex_by_subdomain(dh.size, dx.size, j)
            row_vec.append(h_elems[j])
            col_vec.append(h_elems[i])
            val_vec.append(v * pow(diag[j], -1, p) % p)
    pad = dk.size - len(row_vec)
    row_vec += [h_elems[0]] * pad
    col_vec += [h_elems[0]] * pad
    val_vec += [0] * pad
    row_col_vec = [r * c % p for r, c in zip(row_vec, col_vec)]
    polys = {k: trim(dk.ifft(v)) for k, v in (("row", row_vec), ("col", col_vec), ("val", val_vec),
                                              ("row_col", row_col_vec))}
    out = {"row": polys["row"], "col": polys["col"], "val": polys["val"], "row_col": polys["row_col"],
           "row_evals_on_k": row_vec, "col_evals_on_k": col_vec, "val_evals_on_k": val_vec}
    for k in ("row", "col", "val", "row_col"):
        out[k + "_evals_on_b"] = db.fft(polys[k])
    return out


def index(cs, curve_id):
    """indexer.rs:71-116 for an already synthesised MarlinCS (made square here)"""
    fr = FR[curve_id]
    cs.make_matrices_square()
    a, b, c = cs.matrices()
    nnz = max(sum(map(len, m)) for m in (a, b, c))
    ni, nc = len(cs.input), len(cs.a)
    nv = len(cs.input) + len(cs.witness)
    dx, dh, dk = Domain(fr, ni), Domain(fr, nv), Domain(fr, nnz)
    db = Domain(fr, 3 * dk.size - 3)
    idx = {"curve": curve_id, "num_constraints": nc, "num_variables": nv, "num_non_zeros": nnz, "a": a, "b": b, "c": c,
           "dx": dx, "dh": dh, "dk": dk, "db": db}
    for name, m in (("a", a), ("b", b), ("c", c)):
        idx[name + "_star"] = compose_matrix_polynomials(m, dx, dh, dk, db)
    return idx


# ---- prover -------------------------------------------------------------------------------------------
def prover_init(idx, cs):
    """prover.rs:86-147; cs must already be square (index() squares the same object)"""
    p = FR[idx["curve"]].p
    x, w = cs.input, cs.witness
    ni = len(x)
    if idx["num_constraints"] != len(cs.a) or idx["num_constraints"] != ni + len(w):
        raise ValueError("InstanceDoesNotMatchIndex")
    val = lambda j: x[j] if j < ni else w[j - ni]
    inner = lambda row: sum(co * val(j) for co, j in row) % p
    return {"idx": idx, "x": list(x), "w": list(w), "z_a": [inner(r) for r in idx["a"]], "z_b": [inner(r) for r in idx["b"]]}


def prover_first_round(st, rand_w, rand_za, rand_zb, mask_coeffs):
    """prover.rs:150-222.  rand_*: the single coefficient of DensePolynomial::rand(zk_bound - 1);
    mask_coeffs: the 3|H| coefficients of DensePolynomial::rand(3|H| + 2*zk_bound - 3)."""
    idx = st["idx"]
    fr = FR[idx["curve"]]
    p = fr.p
    dh, dx = idx["dh"], idx["dx"]
    H, X = dh.size, dx.size
    x_poly = trim(dx.ifft(st["x"]))
    x_evals_on_h = dh.fft(x_poly)
    ratio = H // X
    w_ext = st["w"] + [0] * (H - X - len(st["w"]))
    w_evals = [0 if i % ratio == 0 else (w_ext[i - i // ratio - 1] - x_evals_on_h[i]) % p for i in range(H)]
    v_h = lambda r: mul_by_vanishing_poly([r], H, p)
    w_poly = poly_add(trim(dh.ifft(w_evals)), v_h(rand_w), p)
    w_poly, rem = divide_by_vanishing_poly(w_poly, X, p)
    assert not rem, "w is not divisible by v_X"
    z_a_poly = poly_add(trim(dh.ifft(st["z_a"])), v_h(rand_za), p)
    z_b_poly = poly_add(trim(dh.ifft(st["z_b"])), v_h(rand_zb), p)
    assert len(mask_coeffs) == 3 * H
    mask = list(mask_coeffs)
    rem = divide_by_vanishing_poly(mask, H, p)[1]
    mask[0] = (mask[0] - (rem[0] if rem else 0)) % p
    st.update(w_poly=w_poly, z_a_poly=z_a_poly, z_b_poly=z_b_poly, mask_poly=trim(mask), x_poly=x_poly)
    return {"w": w_poly, "z_a": z_a_poly, "z_b": z_b_poly, "mask": trim(mask)}


def prover_second_round(st, alpha, eta_a, eta_b, eta_c):
    """prover.rs:230-321"""
    idx = st["idx"]
    fr = FR[idx["curve"]]
    p = fr.p
    dh, dx = idx["dh"], idx["dx"]
    H, X = dh.size, dx.size
    za, zb = st["z_a_poly"], st["z_b_poly"]
    m = [c * eta_c % p for c in poly_mul(za, zb, fr)]
    for i in range(min(len(m), len(za), len(zb))):
        m[i] = (m[i] + eta_a * za[i] + eta_b * zb[i]) % p
    m_poly = trim(m)
    r_alpha_evals = batch_evals(dh, alpha)
    r_alpha_poly = trim(dh.ifft(r_alpha_evals))
    t_evals = [0] * H
    for matrix, eta in ((idx["a"], eta_a), (idx["b"], eta_b), (idx["c"], eta_c)):
        for i, row in enumerate(matrix):
            for coeff, j in row:
                k = reindex_by_subdomain(H, X, j)
                t_evals[k] = (t_evals[k] + eta * coeff % p * r_alpha_evals[i]) % p
    t_poly = trim(dh.ifft(t_evals))
    z_poly = mul_by_vanishing_poly(st["w_poly"], X, p)
    for i in range(min(len(z_poly), len(st["x_poly"]))):
        z_poly[i] = (z_poly[i] + st["x_poly"][i]) % p
    size = max(len(st["mask_poly"]), len(r_alpha_poly) + len(m_poly), len(t_poly) + len(z_poly))
    d = Domain(fr, size)
    ev = [(r * mm - t * z) % p for r, mm, t, z in zip(d.fft(r_alpha_poly), d.fft(m_poly), d.fft(t_poly), d.fft(z_poly))]
    q1 = poly_add(st["mask_poly"], trim(d.ifft(ev)), p)
    h_1, x_g_1 = divide_by_vanishing_poly(q1, H, p)
    g_1 = trim(x_g_1[1:])
    st.update(t_poly=t_poly, first_msg=(alpha, eta_a, eta_b, eta_c), x_g_1_const=(x_g_1[0] if x_g_1 else 0))
    return {"t": t_poly, "g_1": g_1, "h_1": h_1}


def prover_third_round(st, beta):
    """prover.rs:331-427"""
    idx = st["idx"]
    fr = FR[idx["curve"]]
    p = fr.p
    dh, dk, db = idx["dh"], idx["dk"], idx["db"]
    alpha, eta_a, eta_b, eta_c = st["first_msg"]
    vha, vhb = dh.vanishing_at(alpha), dh.vanishing_at(beta)
    stars = [idx["a_star"], idx["b_star"], idx["c_star"]]
    etas = [eta_a, eta_b, eta_c]
    inv = []
    for s in stars:
        inv.append([pow((beta - r) * (alpha - c) % p, -1, p) if (beta - r) * (alpha - c) % p else 0
                    for r, c in zip(s["row_evals_on_k"], s["col_evals_on_k"])])
    t_evals_on_k = []
    for i in range(dk.size):
        t = sum(eta * s["val_evals_on_k"][i] * iv[i] for eta, s, iv in zip(etas, stars, inv)) % p
        t_evals_on_k.append(t * vha % p * vhb % p)
    t_poly = trim(dk.ifft(t_evals_on_k))
    g_2 = trim(t_poly[1:])
    den = [[(beta * alpha - alpha * r - beta * c + rc) % p
            for r, c, rc in zip(s["row_evals_on_b"], s["col_evals_on_b"], s["row_col_evals_on_b"])] for s in stars]
    a_evals, b_evals = [], []
    for i in range(db.size):
        da, dbb, dc = den[0][i], den[1][i], den[2][i]
        tmp = (eta_a * stars[0]["val_evals_on_b"][i] * dbb * dc + eta_b * stars[1]["val_evals_on_b"][i] * dc * da
               + eta_c * stars[2]["val_evals_on_b"][i] * da * dbb) % p
        a_evals.append(tmp * vha % p * vhb % p)
        b_evals.append(da * dbb * dc % p)
    a_poly, b_poly = trim(db.ifft(a_evals)), trim(db.ifft(b_evals))
    h_2 = divide_by_vanishing_poly(poly_sub(a_poly, poly_mul(b_poly, t_poly, fr), p), dk.size, p)[0]
    st.update(beta=beta)
    return {"g_2": g_2, "h_2": h_2}


# ---- the AHP acceptance test ---------------------------------------------------------------------------
def verifier_equality_check(idx, public_input, polys, alpha, eta_a, eta_b, eta_c, beta, gamma):
    """verifier.rs:128-209 with the evaluations taken directly from the polynomials.
    polys: label -> coefficient list for the 9 prover polynomials; indexer polynomials come from idx."""
    fr = FR[idx["curve"]]
    p = fr.p
    dh, dk = idx["dh"], idx["dk"]
    ev = lambda label, x: poly_eval(polys[label], x, p)
    vha, vhb = dh.vanishing_at(alpha), dh.vanishing_at(beta)
    r_alpha_at_beta = bivariate_eval(dh, alpha, beta)
    formatted = [1] + list(public_input)
    dx = Domain(fr, len(formatted))
    vxb = dx.vanishing_at(beta)
    x_at_beta = poly_eval(trim(dx.ifft(formatted)), beta, p)
    za, zb = ev("z_a", beta), ev("z_b", beta)
    lhs = (ev("mask", beta) + r_alpha_at_beta * (eta_a * za + eta_b * zb + eta_c * za * zb)
           - ev("t", beta) * (vxb * ev("w", beta) + x_at_beta)) % p
    rhs = (ev("h_1", beta) * vhb + beta * ev("g_1", beta)) % p
    if lhs != rhs:
        return False
    vkg = dk.vanishing_at(gamma)
    ab = alpha * beta % p
    den, val = [], []
    for name in ("a", "b", "c"):
        s = idx[name + "_star"]
        e = {k: poly_eval(s[k], gamma, p) for k in ("row", "col", "val", "row_col")}
        den.append((ab - alpha * e["row"] - beta * e["col"] + e["row_col"]) % p)
        val.append(e["val"])
    a_at = (eta_a * val[0] * den[1] * den[2] + eta_b * val[1] * den[2] * den[0] + eta_c * val[2] * den[0] * den[1]) % p
    a_at = a_at * vha % p * vhb % p
    b_at = den[0] * den[1] * den[2] % p
    lhs = ev("h_2", gamma) * vkg % p
    rhs = (a_at - b_at * (gamma * ev("g_2", gamma) + ev("t", beta) * pow(dk.size, -1, p))) % p
    return lhs == rhs
