"""KZG10 with hiding and degree-bound shifted commitments: restatement of the prover side of
marlin/src/pc/kzg10.rs:27-156,185-226 and marlin/src/pc/mod.rs:34-160,241-250 (TEST INFRASTRUCTURE).

Polynomials are coefficient lists (low degree first, canonical ints mod r).  Randomness that the
reference draws from an rng (`Rand::rand(hiding_bound, rng)` = a random polynomial of degree
hiding_bound, data_structures.rs:176-180) is an explicit argument here.  With the trapdoor beta and
gamma (gamma_g = gamma * g) known, `commitment_exponent` / `kzg_check_in_exponent` restate the
pairing check of kzg10.rs:158-172 in Fr.
"""
from .curves import G1
from .fields import FR
from .msm import msm_pippenger


class KzgError(Exception):
    pass


def poly_degree(p):
    d = len(p) - 1
    while d > 0 and p[d] == 0:
        d -= 1
    return d


def poly_eval(p, x, mod):
    acc = 0
    for c in reversed(p):
        acc = (acc * x + c) % mod
    return acc


def poly_div_linear(p, z, mod):
    """p / (x - z): quotient (remainder p(z) dropped), as `p / &divisor` in kzg10.rs:216-217"""
    if len(p) < 2:
        return []
    q = [0] * (len(p) - 1)
    carry = 0
    for i in reversed(range(1, len(p))):
        carry = (p[i] + carry * z) % mod
        q[i - 1] = carry
    return q


def setup(curve_id, max_degree, beta, g_scalar=1, gamma=7):
    """kzg10.rs:27-72 with explicit trapdoor: g = g_scalar * G1 generator, gamma_g = gamma * g."""
    fr = FR[curve_id]
    c = G1(curve_id)
    g = c.mul_affine(c.gen, g_scalar)
    gamma_g = c.mul_affine(g, gamma)
    tg = c.fixed_base_table(g, fr.bits)
    tgg = c.fixed_base_table(gamma_g, fr.bits)
    powers, cur = [], 1
    for _ in range(max_degree + 1):
        powers.append(cur)
        cur = cur * beta % fr.p
    return {"curve": curve_id, "powers_of_g": c.batch_to_affine([c.fixed_base_mul(tg, k) for k in powers]),
            "powers_of_gamma_g": c.batch_to_affine([c.fixed_base_mul(tgg, k) for k in powers]),
            "beta": beta, "gamma": gamma, "g": g, "g_scalar": g_scalar}


def trim(pp, supported_degree):
    """kzg10.rs:74-98"""
    if supported_degree > len(pp["powers_of_g"]) - 1:
        raise KzgError("TrimmingDegreeTooLarge")
    ck = dict(pp)
    ck["powers_of_g"] = pp["powers_of_g"][:supported_degree + 1]
    ck["powers_of_gamma_g"] = pp["powers_of_gamma_g"][:supported_degree + 1]
    ck["supported_degree"] = supported_degree
    return ck


def _skip_leading_zeros(p):
    n = 0
    while n < len(p) and p[n] == 0:
        n += 1
    return n, p[n:]


def kzg_commit(curve_id, powers_of_g, powers_of_gamma_g, p, blinding=None, supported_degree=None):
    """kzg10.rs:100-123.  blinding = coefficient list of the hiding polynomial or None."""
    c = G1(curve_id)
    bits = FR[curve_id].bits
    deg = poly_degree(p)
    sup = len(powers_of_g) - 1 if supported_degree is None else supported_degree
    if deg < 1:
        raise KzgError("DegreeIsZero")
    if deg > sup:
        raise KzgError("DegreeOutOfBound")
    nz, coeffs = _skip_leading_zeros(p)
    comm = msm_pippenger(c, powers_of_g[nz:], coeffs, bits)
    if blinding is not None:
        hb = len(blinding) - 1
        if hb == 0:
            raise KzgError("HidingBoundIsZero")
        if hb > len(powers_of_g):
            raise KzgError("HidingBoundTooLarge")
        rc = c.to_affine(msm_pippenger(c, powers_of_gamma_g, blinding, bits))
        comm = c.add_mixed(comm, rc)
    return c.to_affine(comm)


def kzg_open(curve_id, powers_of_g, powers_of_gamma_g, p, point, blinding=None):
    """kzg10.rs:125-156 -> (w affine, rand_v or None)"""
    fr = FR[curve_id]
    c = G1(curve_id)
    deg = poly_degree(p)
    if deg < 1:
        raise KzgError("DegreeIsZero")
    if deg > len(powers_of_g):
        raise KzgError("DegreeOutOfBound")
    witness = poly_div_linear(p, point, fr.p)
    nz, coeffs = _skip_leading_zeros(witness)
    w = msm_pippenger(c, powers_of_g[nz:], coeffs, fr.bits)
    rand_v = None
    if blinding is not None and any(blinding):
        rand_v = poly_eval(blinding, point, fr.p)
        rw = poly_div_linear(blinding, point, fr.p)
        w = c.add(w, msm_pippenger(c, powers_of_gamma_g, rw, fr.bits))
    return c.to_affine(w), rand_v


# ---- PC layer (pc/mod.rs) ---------------------------------------------------------------------
def pc_commit(ck, polys):
    """polys: list of dicts {coeffs, degree_bound (or None), blinding (or None), shifted_blinding (or None)}
    -> list of (comm, shifted_comm or None)   (pc/mod.rs:34-71)"""
    cid = ck["curve"]
    out = []
    for P in polys:
        comm = kzg_commit(cid, ck["powers_of_g"], ck["powers_of_gamma_g"], P["coeffs"], P.get("blinding"),
                          ck["supported_degree"])
        shifted = None
        if P.get("degree_bound") is not None:
            db = P["degree_bound"]
            if db > ck["supported_degree"]:
                raise KzgError("DegreeOutOfBound")
            sp = ck["powers_of_g"][ck["supported_degree"] - db:]
            shifted = kzg_commit(cid, sp, ck["powers_of_gamma_g"], P["coeffs"], P.get("shifted_blinding"), len(sp) - 1)
        out.append((comm, shifted))
    return out


def _axpy(acc, f, p, mod):
    if len(acc) < len(p):
        acc = acc + [0] * (len(p) - len(acc))
    for i, c in enumerate(p):
        acc[i] = (acc[i] + f * c) % mod
    return acc


def pc_open(ck, polys, point, opening_challenge):
    """pc/mod.rs:73-100: linear combination with powers of the opening challenge, then KZG10::open"""
    cid = ck["curve"]
    mod = FR[cid].p
    p, r = [], []
    challenge = 1
    for P in polys:
        p = _axpy(p, challenge, P["coeffs"], mod)
        r = _axpy(r, challenge, P.get("blinding") or [], mod)
        if P.get("degree_bound") is not None:
            sc = challenge * opening_challenge % mod
            shift = ck["supported_degree"] - P["degree_bound"]
            shifted = [0] * shift + list(P["coeffs"]) if any(P["coeffs"]) else []      # shift_polynomial :241-250
            p = _axpy(p, sc, shifted, mod)
            r = _axpy(r, sc, P.get("shifted_blinding") or [], mod)
        challenge = challenge * opening_challenge % mod * opening_challenge % mod
    return kzg_open(cid, ck["powers_of_g"], ck["powers_of_gamma_g"], p, point, r if any(r) else None)


# ---- the pairing checks restated in the exponent (trapdoor known) -----------------------------------
def commitment_exponent(ck, p, blinding=None, shift=0):
    """discrete log (base g) of commit(p): beta^shift * p(beta) + gamma * blinding(beta)"""
    mod = FR[ck["curve"]].p
    beta, gamma = ck["beta"], ck["gamma"]
    e = pow(beta, shift, mod) * poly_eval(p, beta, mod)
    if blinding is not None:
        e += gamma * poly_eval(blinding, beta, mod)
    return e % mod


def kzg_check_in_exponent(ck, comm_exp, point, value, w_exp, rand_v):
    """kzg10.rs:158-172 with e(a*g, h) == e(b*g, c*h)  <=>  a == b*c"""
    mod = FR[ck["curve"]].p
    lhs = (comm_exp - value - (rand_v or 0) * ck["gamma"]) % mod
    return lhs == w_exp * (ck["beta"] - point) % mod


def exponent_point(ck, e):
    c = G1(ck["curve"])
    return c.mul_affine(ck["g"], e)
