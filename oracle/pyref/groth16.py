"""Groth16 setup / prove / verify: restatement of groth16/src/{generator,prover,
r1cs_to_qap,verifier}.rs of the reference.

* `generate_parameters`  <- generator.rs:135-286 (toxic waste and the generators are
  explicit arguments instead of rng draws; FixedBaseMSM is replaced by a
  result-identical fixed-base table multiplication).
* `witness_map`          <- r1cs_to_qap.rs:113-172 (the literal 7-transform pipeline).
* `create_proof`         <- prover.rs:124-228 (including the `r != 0` guard on B-G1).
* `verify_proof`         <- verifier.rs:18-44 (needs pairing.py).
* `proof_in_exponent`    <- independent check: with toxic waste known, the three
  proof elements are fixed multiples of the generators.
"""
from .curves import G1, G2
from .fields import FR
from .msm import msm_pippenger
from .ntt import Domain


class Parameters:
    """groth16/src/lib.rs:81-91 (+ VerifyKey :59-66)."""
    pass


def generate_parameters(cs, curve_id, alpha, beta, gamma, delta, t, g1_gen=None, g2_gen=None):
    fr = FR[curve_id]
    p = fr.p
    c1, c2 = G1(curve_id), G2(curve_id)
    g1_gen = g1_gen or c1.gen
    g2_gen = g2_gen or c2.gen
    num_inputs, num_aux, ncons = cs.num_inputs, cs.num_aux, cs.num_constraints
    # generator.rs:165-168 / r1cs_to_qap.rs:62-70
    domain = Domain(fr, ncons + (num_inputs - 1) + 1)
    m_raw = domain.size
    zt = domain.vanishing_at(t)
    assert zt != 0, "t must lie outside the domain"
    u = domain.lagrange_coeffs_at(t)
    nvars = (num_inputs - 1) + num_aux
    a = [0] * (nvars + 1)
    b = [0] * (nvars + 1)
    c = [0] * (nvars + 1)
    for i in range(num_inputs):                      # r1cs_to_qap.rs:78-80
        a[i] = u[ncons + i]
    for which, dst in (("a", a), ("b", b), ("c", c)):  # :82-107
        for i, row in enumerate(cs.rows(which)):
            for co, idx in row:
                dst[idx] = (dst[idx] + u[i] * co) % p
    gamma_inv, delta_inv = pow(gamma, -1, p), pow(delta, -1, p)
    gamma_abc = [(beta * a[i] + alpha * b[i] + c[i]) * gamma_inv % p for i in range(num_inputs)]
    l = [(beta * ai + alpha * bi + ci) * delta_inv % p for ai, bi, ci in zip(a, b, c)]
    t1 = c1.fixed_base_table(g1_gen, fr.bits)
    t2 = c2.fixed_base_table(g2_gen, fr.bits)
    m1 = lambda k: c1.fixed_base_mul(t1, k % p)
    m2 = lambda k: c2.fixed_base_mul(t2, k % p)
    pk = Parameters()
    pk.curve_id = curve_id
    pk.alpha_g1 = c1.to_affine(m1(alpha))
    pk.beta_g1 = c1.to_affine(m1(beta))
    pk.beta_g2 = c2.to_affine(m2(beta))
    pk.gamma_g2 = c2.to_affine(m2(gamma))
    pk.delta_g1 = c1.to_affine(m1(delta))
    pk.delta_g2 = c2.to_affine(m2(delta))
    pk.a_query = c1.batch_to_affine([m1(x) for x in a])
    pk.b_g1_query = c1.batch_to_affine([m1(x) for x in b])
    pk.b_g2_query = c2.batch_to_affine([m2(x) for x in b])
    hs, tp = [], 1                                    # generator.rs:235-242
    for _ in range(m_raw - 1):
        hs.append(zt * delta_inv % p * tp % p)
        tp = tp * t % p
    pk.h_query = c1.batch_to_affine([m1(x) for x in hs])
    pk.l_query = c1.batch_to_affine([m1(x) for x in l])[num_inputs:]   # :247
    pk.gamma_abc_g1 = c1.batch_to_affine([m1(x) for x in gamma_abc])
    # kept for the in-the-exponent check only (never crosses the ABI)
    pk._dlog = dict(a=a, b=b, c=c, l=l, hs=hs, alpha=alpha, beta=beta, gamma=gamma, delta=delta,
                    g1=g1_gen, g2=g2_gen)
    return pk


def evaluate_constraint(row, z, p):
    """r1cs_to_qap.rs:15-52."""
    return sum(co * z[idx] for co, idx in row) % p


def witness_map(cs, curve_id):
    """r1cs_to_qap.rs:113-172.  Returns h (canonical ints, length = domain size)."""
    fr = FR[curve_id]
    p = fr.p
    num_inputs, ncons = cs.num_inputs, cs.num_constraints
    z = cs.full_assignment()
    domain = Domain(fr, ncons + num_inputs)
    n = domain.size
    a = [0] * n
    b = [0] * n
    for i, (ra, rb) in enumerate(zip(cs.rows("a"), cs.rows("b"))):
        a[i] = evaluate_constraint(ra, z, p)
        b[i] = evaluate_constraint(rb, z, p)
    for i in range(num_inputs):
        a[ncons + i] = z[i]
    a = domain.coset_fft(domain.ifft(a))
    b = domain.coset_fft(domain.ifft(b))
    ab = [x * y % p for x, y in zip(a, b)]
    c = [0] * n
    for i, rc in enumerate(cs.rows("c")):
        c[i] = evaluate_constraint(rc, z, p)
    c = domain.coset_fft(domain.ifft(c))
    zinv = pow(domain.vanishing_at(fr.generator), -1, p)
    ab = [(x - y) * zinv % p for x, y in zip(ab, c)]
    return domain.coset_ifft(ab)


def calculate_coeff(curve, initial, query, vk_param, assignment, bits):
    """prover.rs:213-228."""
    acc = msm_pippenger(curve, query[1:], assignment, bits)
    res = curve.add_mixed(initial, query[0])
    res = curve.add(res, acc)
    return curve.add_mixed(res, vk_param)


def create_proof(pk, cs, r, s):
    """prover.rs:124-211.  Returns (A, B, C) affine (None = identity)."""
    cid = pk.curve_id
    fr = FR[cid]
    p, bits = fr.p, fr.bits
    c1, c2 = G1(cid), G2(cid)
    h = witness_map(cs, cid)
    assignment = cs.input_assignment[1:] + cs.aux_assignment
    r_g1 = c1.mul(c1.from_affine(pk.delta_g1), r)
    g_a = calculate_coeff(c1, r_g1, pk.a_query, pk.alpha_g1, assignment, bits)
    if r != 0:
        s_g1 = c1.mul(c1.from_affine(pk.delta_g1), s)
        g1_b = calculate_coeff(c1, s_g1, pk.b_g1_query, pk.beta_g1, assignment, bits)
    else:
        g1_b = c1.identity()
    s_g2 = c2.mul(c2.from_affine(pk.delta_g2), s)
    g2_b = calculate_coeff(c2, s_g2, pk.b_g2_query, pk.beta_g2, assignment, bits)
    h_acc = msm_pippenger(c1, pk.h_query, h, bits)
    l_acc = msm_pippenger(c1, pk.l_query, cs.aux_assignment, bits)
    s_g_a = c1.mul(g_a, s)
    r_g1_b = c1.mul(g1_b, r)
    r_s_delta = c1.mul(c1.mul(c1.from_affine(pk.delta_g1), r), s)
    g_c = c1.add(s_g_a, r_g1_b)
    g_c = c1.add(g_c, c1.neg(r_s_delta))
    g_c = c1.add(g_c, l_acc)
    g_c = c1.add(g_c, h_acc)
    return c1.to_affine(g_a), c2.to_affine(g2_b), c1.to_affine(g_c)


def proof_in_exponent(pk, cs, r, s, h=None):
    """Independent derivation of the proof from the toxic waste:
    A = (alpha + sum z_i a_i(t) + r delta) G1, B = (beta + sum z_i b_i(t) + s delta) G2,
    C = (sum_aux z_i l_i + sum h_j hs_j + s A' + r B' - r s delta) G1."""
    cid = pk.curve_id
    p = FR[cid].p
    d = pk._dlog
    z = cs.full_assignment()
    if h is None:
        h = witness_map(cs, cid)
    A = (d["alpha"] + sum(zi * ai for zi, ai in zip(z, d["a"])) + r * d["delta"]) % p
    B = (d["beta"] + sum(zi * bi for zi, bi in zip(z, d["b"])) + s * d["delta"]) % p
    ni = cs.num_inputs
    C = (sum(zi * li for zi, li in zip(z[ni:], d["l"][ni:]))
         + sum(hj * x for hj, x in zip(h, d["hs"]))
         + s * A + r * B - r * s * d["delta"]) % p
    if r == 0:                                   # the literal guard of prover.rs:170
        C = (C - r * B) % p
    c1, c2 = G1(cid), G2(cid)
    return c1.mul_affine(d["g1"], A), c2.mul_affine(d["g2"], B), c1.mul_affine(d["g1"], C)


def verifying_key(pk):
    """VerifyKey<E> of the parameters (groth16/src/lib.rs:59-66) in the form pairing.prepare_verifying_key takes"""
    return {"alpha_g1": pk.alpha_g1, "beta_g2": pk.beta_g2, "gamma_g2": pk.gamma_g2, "delta_g2": pk.delta_g2,
            "gamma_abc_g1": pk.gamma_abc_g1}


def verify_proof(pk, proof, public_inputs):
    """prepare_verifying_key + verify_proof (verifier.rs:8-44) -- the reference's own acceptance test
    (groth16/tests/mini.rs:85-89).  public_inputs excludes ONE, like the reference's `&[Fr]` argument."""
    from . import pairing
    pvk = pairing.prepare_verifying_key(pk.curve_id, verifying_key(pk))
    return pairing.verify_proof(pk.curve_id, pvk, proof, public_inputs)
