"""zkp_marlin's crate-level API restated on the oracle's integers (TEST INFRASTRUCTURE, never imported by the product):

    index               marlin/src/lib.rs:67-95       AHP::index, PC::trim, PC::commit of the index polynomials
    create_random_proof marlin/src/lib.rs:97-181      three AHP rounds <-> PC::commit <-> Fiat-Shamir, evaluations, batch_open
    verify_proof        marlin/src/lib.rs:184-260     the reference's acceptance test (marlin/tests/mini.rs:81,87):
                                                      verifier_equality_check + PC::batch_check with REAL pairings

Byte layouts hashed into the transcript are written here independently of ckb_zkp_b200/fs_rng.py (same recalled
arkworks-0.2 `ToBytes` rules, see oracle/pyref/transcript.py); the prover's own randomness is a Python `random.Random`
consumed in the reference's draw order (prover.rs:150-222, then `Rand::rand` per hiding commitment, kzg10.rs:112-116).
"""
from . import kzg10 as K
from . import marlin as M
from . import pairing as PR
from .curves import CURVES
from .fields import FQ, FR
from .ntt import Domain
from .transcript import FiatShamirRng

INDEXER = ["a_row", "a_col", "a_val", "a_row_col", "b_row", "b_col", "b_val", "b_row_col", "c_row", "c_col", "c_val",
           "c_row_col"]


# ---- ToBytes -------------------------------------------------------------------------------------------
def _fq(cid, v):
    return int(v).to_bytes(8 * FQ[cid].limbs, "little")


def g1_bytes(cid, P):
    if P is None:
        return _fq(cid, 0) + _fq(cid, 1) + b"\x01"
    return _fq(cid, P[0]) + _fq(cid, P[1]) + b"\x00"


def g2_bytes(cid, P):
    if P is None:
        return _fq(cid, 0) * 2 + _fq(cid, 1) + _fq(cid, 0) + b"\x01"
    return b"".join(_fq(cid, c) for c in (P[0][0], P[0][1], P[1][0], P[1][1])) + b"\x00"


def commitment_bytes(cid, comm):
    c, shifted = comm
    return g1_bytes(cid, c) + (b"\x01" if shifted is not None else b"\x00") + g1_bytes(cid, shifted)


def ivk_bytes(ivk):
    cid = ivk["curve"]
    vk = ivk["verifier_key"]
    out = b"".join(int(v).to_bytes(8, "little") for v in ivk["index_info"])
    out += len(ivk["index_comms"]).to_bytes(4, "little")
    out += b"".join(commitment_bytes(cid, c) for c in ivk["index_comms"])
    out += g1_bytes(cid, vk["g"]) + g1_bytes(cid, vk["gamma_g"]) + g2_bytes(cid, vk["h"]) + g2_bytes(cid, vk["beta_h"])
    return out + int(vk["supported_degree"]).to_bytes(8, "little")


def fr_vec_bytes(vals):
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


# ---- setup / index ---------------------------------------------------------------------------------------
def max_degree(nc, nv, nnz):
    size = lambda n: 1 << max(n - 1, 0).bit_length()         # compute_size_of_domain
    h, k = size(max(nc, nv)), size(nnz)
    return max(3 * h + 2 * 1 - 1, 3 * k - 3)                 # AHP::max_degree (ahp/mod.rs:66-84), zk_bound = 1


def universal_setup(cid, max_deg, beta, kg, kgamma, kh):
    """KZG10::setup with g = kg * G1, gamma_g = kgamma * G1, h = kh * G2, beta_h = beta * h (kzg10.rs:27-72)"""
    p = FR[cid].p
    pp = K.setup(cid, max_deg, beta, g_scalar=kg, gamma=kgamma * pow(kg, -1, p) % p)
    g2 = CURVES[(cid, 2)]
    pp["h"] = g2.mul_affine(g2.gen, kh)
    pp["beta_h"] = g2.mul_affine(g2.gen, kh * beta % p)
    return pp


def index(pp, cs):
    """lib.rs:67-95 -> (ipk, ivk) as dicts"""
    cid = pp["curve"]
    idx = M.index(cs, cid)
    deg = max_degree(idx["num_constraints"], idx["num_variables"], idx["num_non_zeros"])
    if len(pp["powers_of_g"]) - 1 < deg:
        raise ValueError("IndexTooLarge")
    ck = K.trim(pp, deg)
    polys = [{"label": l, "coeffs": idx[l[0] + "_star"][l[2:]], "degree_bound": None} for l in INDEXER]
    comms = K.pc_commit(ck, polys)
    vk = {"g": ck["powers_of_g"][0], "gamma_g": ck["powers_of_gamma_g"][0], "h": pp["h"], "beta_h": pp["beta_h"],
          "supported_degree": deg}
    ivk = {"curve": cid, "index_info": (idx["num_variables"], idx["num_constraints"], idx["num_non_zeros"]),
           "index_comms": comms, "verifier_key": vk}
    return {"index": idx, "index_polys": polys, "ivk": ivk, "ck": ck}, ivk


# ---- prover ------------------------------------------------------------------------------------------------
def _outside(domain, fs, p):
    t = fs.rand_fr(p)
    while domain.vanishing_at(t) == 0:
        t = fs.rand_fr(p)
    return t


def query_set(beta, gamma):
    return sorted([(l, beta) for l in ("w", "z_a", "z_b", "mask", "t", "g_1", "h_1")] + [(l, gamma) for l in ["g_2", "h_2"] + INDEXER])


def create_random_proof(ipk, cs, zk_rng):
    """lib.rs:97-181; cs is the (already squared) MarlinCS the index was built from, holding the assignment"""
    idx, ck, ivk = ipk["index"], ipk["ck"], ipk["ivk"]
    cid = idx["curve"]
    p = FR[cid].p
    H, Ksz = idx["dh"].size, idx["dk"].size
    st = M.prover_init(idx, cs)
    fs = FiatShamirRng(ivk_bytes(ivk) + fr_vec_bytes(cs.input[1:]))
    labeled = list(ipk["index_polys"])
    rounds = []

    def commit(polys):
        for P in polys:                                      # Rand::rand(hiding_bound): hiding_bound + 1 coefficients
            if P.get("hiding_bound") is not None:
                P["blinding"] = [zk_rng.randrange(p) for _ in range(P["hiding_bound"] + 1)]
                if P.get("degree_bound") is not None:
                    P["shifted_blinding"] = [zk_rng.randrange(p) for _ in range(P["hiding_bound"] + 1)]
        comms = K.pc_commit(ck, polys)
        labeled.extend(polys)
        rounds.append(comms)
        fs.absorb(b"".join(commitment_bytes(cid, c) for c in comms))

    draws = [zk_rng.randrange(p) for _ in range(3)]
    mask = [zk_rng.randrange(p) for _ in range(3 * H)]
    o1 = M.prover_first_round(st, draws[0], draws[1], draws[2], mask)
    commit([{"label": "w", "coeffs": o1["w"], "degree_bound": None, "hiding_bound": 1},
            {"label": "z_a", "coeffs": o1["z_a"], "degree_bound": None, "hiding_bound": 1},
            {"label": "z_b", "coeffs": o1["z_b"], "degree_bound": None, "hiding_bound": 1},
            {"label": "mask", "coeffs": o1["mask"], "degree_bound": None}])
    alpha = _outside(idx["dh"], fs, p)
    eta_a, eta_b, eta_c = fs.rand_fr(p), fs.rand_fr(p), fs.rand_fr(p)
    o2 = M.prover_second_round(st, alpha, eta_a, eta_b, eta_c)
    commit([{"label": "t", "coeffs": o2["t"], "degree_bound": None},
            {"label": "g_1", "coeffs": o2["g_1"], "degree_bound": H - 2, "hiding_bound": 1},
            {"label": "h_1", "coeffs": o2["h_1"], "degree_bound": None}])
    beta = _outside(idx["dh"], fs, p)
    o3 = M.prover_third_round(st, beta)
    commit([{"label": "g_2", "coeffs": o3["g_2"], "degree_bound": Ksz - 2},
            {"label": "h_2", "coeffs": o3["h_2"], "degree_bound": None}])
    gamma = fs.rand_fr(p)
    qs = query_set(beta, gamma)
    by_label = {P["label"]: P for P in labeled}
    evaluations = [K.poly_eval(by_label[l]["coeffs"], pt, p) for l, pt in qs]
    fs.absorb(fr_vec_bytes(evaluations))
    opening_challenge = fs.rand_u128()
    proofs = []
    for point in sorted({pt for _, pt in qs}):                # BTreeMap<point, BTreeSet<label>> (pc/mod.rs:122-160)
        labels = sorted(l for l, pt in qs if pt == point)
        proofs.append(K.pc_open(ck, [by_label[l] for l in labels], point, opening_challenge))
    return {"commitments": rounds, "evaluations": evaluations, "opening_proofs": proofs,
            "challenges": {"alpha": alpha, "eta_a": eta_a, "eta_b": eta_b, "eta_c": eta_c, "beta": beta, "gamma": gamma,
                           "opening_challenge": opening_challenge}}


# ---- verifier ------------------------------------------------------------------------------------------------
def _equality_check(info, cid, public_input, ev, alpha, eta_a, eta_b, eta_c, beta, gamma):
    """AHP::verifier_equality_check (ahp/verifier.rs:128-209) on the proof's evaluations (not on polynomials)"""
    fr = FR[cid]
    p = fr.p
    nv, nc, nnz = info
    dh, dk = Domain(fr, nc), Domain(fr, nnz)
    vha, vhb = dh.vanishing_at(alpha), dh.vanishing_at(beta)
    r_alpha_at_beta = M.bivariate_eval(dh, alpha, beta)
    formatted = [1] + list(public_input)
    dx = Domain(fr, len(formatted))
    x_at_beta = M.poly_eval(M.trim(dx.ifft(formatted)), beta, p)
    za, zb = ev[("z_a", beta)], ev[("z_b", beta)]
    lhs = (ev[("mask", beta)] + r_alpha_at_beta * (eta_a * za + eta_b * zb + eta_c * za * zb)
           - ev[("t", beta)] * (dx.vanishing_at(beta) * ev[("w", beta)] + x_at_beta)) % p
    if lhs != (ev[("h_1", beta)] * vhb + beta * ev[("g_1", beta)]) % p:
        return False
    ab = alpha * beta % p
    den, val = [], []
    for name in "abc":
        e = {k: ev[("%s_%s" % (name, k), gamma)] for k in ("row", "col", "val", "row_col")}
        den.append((ab - alpha * e["row"] - beta * e["col"] + e["row_col"]) % p)
        val.append(e["val"])
    a_at = (eta_a * val[0] * den[1] * den[2] + eta_b * val[1] * den[2] * den[0] + eta_c * val[2] * den[0] * den[1]) % p
    a_at = a_at * vha % p * vhb % p
    b_at = den[0] * den[1] * den[2] % p
    lhs = ev[("h_2", gamma)] * dk.vanishing_at(gamma) % p
    return lhs == (a_at - b_at * (gamma * ev[("g_2", gamma)] + ev[("t", beta)] * pow(dk.size, -1, p))) % p


def _pc_check(cid, vk, comms, bounds, point, values, proof, opening_challenge):
    """PC::check (pc/mod.rs:102-121): accumulate_commitments_and_values (:213-250) then KZG10::check (kzg10.rs:158-172)"""
    p = FR[cid].p
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    acc, acc_v, ch = g1.identity(), 0, 1
    for (comm, shifted), db, v in zip(comms, bounds, values):
        assert (db is not None) == (shifted is not None)
        acc = g1.add(acc, g1.mul(g1.from_affine(comm), ch))
        acc_v = (acc_v + v * ch) % p
        if db is not None:
            sc = ch * opening_challenge % p
            acc = g1.add(acc, g1.mul(g1.from_affine(shifted), sc))
            acc_v = (acc_v + pow(point, vk["supported_degree"] - db, p) * v % p * sc) % p
        ch = ch * opening_challenge % p * opening_challenge % p
    w, rand_v = proof
    u = g1.add(acc, g1.neg(g1.mul(g1.from_affine(vk["g"]), acc_v)))
    if rand_v is not None:
        u = g1.add(u, g1.neg(g1.mul(g1.from_affine(vk["gamma_g"]), rand_v)))
    v2 = g2.add(g2.from_affine(vk["beta_h"]), g2.neg(g2.mul(g2.from_affine(vk["h"]), point)))
    return PR.pairing(cid, g1.to_affine(u), vk["h"]) == PR.pairing(cid, w, g2.to_affine(v2))


def verify_proof(ivk, proof, public_input):
    """lib.rs:184-260.  proof: dict with commitments (3 rounds of (comm, shifted)), evaluations, opening_proofs"""
    cid = ivk["curve"]
    p = FR[cid].p
    nv, nc, nnz = ivk["index_info"]
    if nc != nv:
        raise ValueError("NonSquareMatrix")
    dh = Domain(FR[cid], nc)
    fs = FiatShamirRng(ivk_bytes(ivk) + fr_vec_bytes(public_input))
    first, second, third = proof["commitments"]
    fs.absorb(b"".join(commitment_bytes(cid, c) for c in first))
    alpha = _outside(dh, fs, p)
    eta_a, eta_b, eta_c = fs.rand_fr(p), fs.rand_fr(p), fs.rand_fr(p)
    fs.absorb(b"".join(commitment_bytes(cid, c) for c in second))
    beta = _outside(dh, fs, p)
    fs.absorb(b"".join(commitment_bytes(cid, c) for c in third))
    gamma = fs.rand_fr(p)
    qs = query_set(beta, gamma)
    fs.absorb(fr_vec_bytes(proof["evaluations"]))
    opening_challenge = fs.rand_u128()
    H, Ksz = dh.size, Domain(FR[cid], nnz).size
    labels = INDEXER + ["w", "z_a", "z_b", "mask", "t", "g_1", "h_1", "g_2", "h_2"]
    bounds = dict.fromkeys(labels)
    bounds["g_1"], bounds["g_2"] = H - 2, Ksz - 2
    comms = dict(zip(labels, list(ivk["index_comms"]) + list(first) + list(second) + list(third)))
    ev = {(l, pt): e for (l, pt), e in zip(qs, proof["evaluations"])}
    if not _equality_check(ivk["index_info"], cid, public_input, ev, alpha, eta_a, eta_b, eta_c, beta, gamma):
        return False
    points = sorted({pt for _, pt in qs})
    assert len(points) == len(proof["opening_proofs"])
    ok = True
    for point, pr in zip(points, proof["opening_proofs"]):
        ls = sorted(l for l, pt in qs if pt == point)
        ok &= _pc_check(cid, ivk["verifier_key"], [comms[l] for l in ls], [bounds[l] for l in ls], point,
                        [ev[(l, point)] for l in ls], pr, opening_challenge)
    return ok
