"""Marlin AHP prover rounds on the GPU (ckb_zkp_b200/marlin.py) vs the oracle's restatement of
marlin/src/ahp/prover.rs, on the reference's own test circuit (marlin/tests/mini.rs:12-41) and on a
MiMC chain shaped like marlin/examples/mimc.rs; then the reference's AHP acceptance test
(verifier_equality_check, ahp/verifier.rs:128-209) on the GPU's polynomials, and the whole
commit -> open flow through the KZG10 layer as in marlin/src/lib.rs:97-181 (challenges supplied)."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import kzg10 as zk
from ckb_zkp_b200 import marlin as zm
from ckb_zkp_b200.backend import CsrMatrix
from oracle.pyref import kzg10 as OK
from oracle.pyref import marlin as OM
from oracle.pyref.fields import BLS12_381, BN254, FR, stream_field
from tests import helpers as H

pytestmark = pytest.mark.gpu
ONE = ("in", 0)


def mini(cs, num=10):
    vx, vy = cs.alloc(2), cs.alloc(3)
    vz = cs.alloc_input(10)
    for _ in range(num):
        cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])


def mimc(cs, n, seed=5):
    p = cs.p
    xl_v, xr_v = stream_field(seed, 0, p), stream_field(seed, 1, p)
    xl, xr = cs.alloc(xl_v), cs.alloc(xr_v)
    for i in range(n // 2):
        c = stream_field(seed, 2 + i, p)
        tmp_v = (xl_v + c) * (xl_v + c) % p
        tmp = cs.alloc(tmp_v)
        cs.enforce([(1, xl), (c, ONE)], [(1, xl), (c, ONE)], [(1, tmp)])
        new_v = ((xl_v + c) * tmp_v + xr_v) % p
        new = cs.alloc_input(new_v) if i == n // 2 - 1 else cs.alloc(new_v)
        cs.enforce([(1, tmp)], [(1, xl), (c, ONE)], [(1, new), (p - 1, xr)])
        xr, xr_v, xl, xl_v = xl, xl_v, new, new_v


class ReplayRng:
    def __init__(self, values):
        self.values, self.i = list(values), 0

    def randrange(self, _):
        v = self.values[self.i]
        self.i += 1
        return v


def device_index(cid, oidx):
    """oracle index (ints) -> ckb_zkp_b200.marlin.Index (Montgomery arrays)"""
    mats = {}
    for name in "abc":
        ptr, cols, vals = [0], [], []
        for row in oidx[name]:
            for co, j in row:
                cols.append(j)
                vals.append(co)
            ptr.append(len(cols))
        mats[name] = CsrMatrix(np.asarray(ptr, dtype=np.uint32), np.asarray(cols, dtype=np.uint32), H.fr_array(cid, vals))
    stars = {n: {k: H.fr_array(cid, oidx[n + "_star"][k]) for k in
                 ("row_evals_on_k", "col_evals_on_k", "val_evals_on_k", "row_evals_on_b", "col_evals_on_b",
                  "val_evals_on_b", "row_col_evals_on_b")} for n in "abc"}
    return zm.Index(cid, oidx["num_constraints"], oidx["num_variables"], oidx["num_non_zeros"], oidx["dx"].size, mats, stars)


def outside(domain, rng, p):
    while True:
        t = rng.randrange(p)
        if domain.vanishing_at(t) != 0:
            return t


@pytest.mark.parametrize("resident", [False, True])      # host-buffer API / round state resident in HBM
@pytest.mark.parametrize("cid,build", [(BLS12_381, mini), (BN254, lambda cs: mimc(cs, 12)), (BN254, lambda cs: mimc(cs, 60)),
                                       (BN254, lambda cs: mimc(cs, 1000))])   # |H| = 2^10, |K| = 2^11, |B| = 2^13
def test_ahp_rounds_match_oracle(ctx, cid, build, resident):
    fr = FR[cid]
    p = fr.p
    rng = random.Random(31)
    cs = OM.MarlinCS(p)
    build(cs)
    oidx = OM.index(cs, cid)
    ost = OM.prover_init(oidx, cs)
    idx = device_index(cid, oidx)
    assert (idx.x_size, idx.h_size, idx.k_size, idx.b_size) == (oidx["dx"].size, oidx["dh"].size, oidx["dk"].size, oidx["db"].size)
    st = zm.prover_init(ctx, idx, H.fr_array(cid, cs.input), H.fr_array(cid, cs.witness), resident=resident)
    assert H.fr_ints(cid, zm.to_host(st.z_a)) == ost["z_a"] and H.fr_ints(cid, zm.to_host(st.z_b)) == ost["z_b"]

    Hs = oidx["dh"].size
    draws = [rng.randrange(p) for _ in range(3 + 3 * Hs)]
    o1 = OM.prover_first_round(ost, draws[0], draws[1], draws[2], draws[3:])
    g1 = zm.prover_first_round(st, ReplayRng(draws))
    alpha = outside(oidx["dh"], rng, p)
    etas = [rng.randrange(p) for _ in range(3)]
    o2 = OM.prover_second_round(ost, alpha, *etas)
    g2 = zm.prover_second_round(st, alpha, *etas)
    beta = outside(oidx["dh"], rng, p)
    o3 = OM.prover_third_round(ost, beta)
    g3 = zm.prover_third_round(st, beta)
    want = {**o1, **o2, **o3}
    got = {label: H.fr_ints(cid, zm.to_host(poly)) for label, poly, _, _ in g1 + g2 + g3}
    for label in ("w", "z_a", "z_b", "mask", "t", "g_1", "h_1", "g_2", "h_2"):
        assert got[label] == want[label], label
    bounds = {label: (db, hb) for label, _, db, hb in g1 + g2 + g3}
    assert bounds["g_1"] == (Hs - 2, 1) and bounds["g_2"] == (oidx["dk"].size - 2, None) and bounds["w"] == (None, 1)
    # the reference's AHP acceptance test on the GPU's polynomials
    gamma = rng.randrange(p)
    assert OM.verifier_equality_check(oidx, cs.input[1:], got, alpha, *etas, beta, gamma)


@pytest.mark.parametrize("resident", [False, True])
def test_marlin_commit_and_open_flow(ctx, resident):
    """marlin/src/lib.rs:97-181 with the challenges supplied: three rounds of PC::commit over the AHP
    oracles, evaluations at beta / gamma, PC::batch_open -- commitments and opening proofs compared with
    the oracle's KZG layer, and every opening checked in the exponent (PC::check)."""
    cid = BN254
    fr = FR[cid]
    p = fr.p
    rng = random.Random(8)
    cs = OM.MarlinCS(p)
    mimc(cs, 12)
    oidx = OM.index(cs, cid)
    idx = device_index(cid, oidx)
    st = zm.prover_init(ctx, idx, H.fr_array(cid, cs.input), H.fr_array(cid, cs.witness), resident=resident)
    Hs, Ks = idx.h_size, idx.k_size
    max_degree = max(3 * Hs + 2 - 1, 3 * Ks - 3)                              # AHP::max_degree (ahp/mod.rs:66-84)
    pp = OK.setup(cid, max_degree, rng.randrange(p), g_scalar=rng.randrange(1, p), gamma=rng.randrange(1, p))
    ock = OK.trim(pp, max_degree)
    ck = zk.CommitterKey(ctx, cid, H.points_array(cid, 1, ock["powers_of_g"]), H.points_array(cid, 1, ock["powers_of_gamma_g"]),
                         max_degree)
    draws = [rng.randrange(p) for _ in range(3 + 3 * Hs)]
    r1 = zm.prover_first_round(st, ReplayRng(draws))
    alpha, etas = outside(oidx["dh"], rng, p), [rng.randrange(p) for _ in range(3)]
    r2 = zm.prover_second_round(st, alpha, *etas)
    beta = outside(oidx["dh"], rng, p)
    r3 = zm.prover_third_round(st, beta)
    gamma = rng.randrange(p)

    labeled, opolys, all_comms, all_rands = [], [], [], []
    for rnd in (r1, r2, r3):
        polys = [zk.LabeledPolynomial(label, poly, db, hb) for label, poly, db, hb in rnd]
        blinds = []
        for label, poly, db, hb in rnd:
            b1 = [rng.randrange(p) for _ in range(2)] if hb else None
            b2 = [rng.randrange(p) for _ in range(2)] if (hb and db is not None) else None
            blinds += (b1 or []) + (b2 or [])
            opolys.append({"label": label, "coeffs": H.fr_ints(cid, zm.to_host(poly)), "degree_bound": db, "blinding": b1,
                           "shifted_blinding": b2})
        comms, rands = zk.pc_commit(ck, polys, ReplayRng(blinds))              # lib.rs:109-110,117-118,124-125
        labeled += polys
        all_comms += comms
        all_rands += rands
    ocomms = OK.pc_commit(ock, opolys)
    pt = lambda g: H.array_point(cid, 1, g[0], g[1])
    for (c, sc), (oc, osc), P in zip(all_comms, ocomms, opolys):
        assert pt(c) == oc, P["label"]
        assert (sc is None) == (osc is None) and (sc is None or pt(sc) == osc), P["label"]
    # query set of the prover polynomials (ahp/verifier.rs:79-104) and the evaluations (lib.rs:147-156)
    at_beta = ["w", "z_a", "z_b", "mask", "t", "g_1", "h_1"]
    at_gamma = ["g_2", "h_2"]
    query = [(l, beta) for l in at_beta] + [(l, gamma) for l in at_gamma]
    by_label = {P.label: P for P in labeled}
    f = zm.Field(cid)
    for label, point in query:
        ev = f.to_int(ctx.poly_eval(cid, by_label[label].coeffs, f.mont(point)))
        assert ev == OK.poly_eval(H.fr_ints(cid, zm.to_host(by_label[label].coeffs)), point, p)
    opening_challenge = rng.randrange(1 << 128)                               # u128::rand (lib.rs:158)
    proofs = zk.pc_batch_open(ck, labeled, query, opening_challenge, all_rands)
    o_by_label = {P["label"]: P for P in opolys}
    for (w, rand_v), (point, labels) in zip(proofs, sorted([(beta, at_beta), (gamma, at_gamma)])):
        group = [o_by_label[l] for l in sorted(labels)]
        ow, orv = OK.pc_open(ock, group, point, opening_challenge)
        assert pt(w) == ow
        assert (rand_v is None) == (orv is None) and (rand_v is None or f.to_int(rand_v) == orv)
        # PC::check (pc/mod.rs:102-120) in the exponent: accumulate commitments and values, then KZG10::check
        acc_c, acc_v, ch = 0, 0, 1
        sup = ock["supported_degree"]
        for P in group:
            acc_c += ch * OK.commitment_exponent(ock, P["coeffs"], P["blinding"])
            v = OK.poly_eval(P["coeffs"], point, p)
            acc_v += ch * v
            if P["degree_bound"] is not None:
                sc = ch * opening_challenge % p
                sh = sup - P["degree_bound"]
                acc_c += sc * OK.commitment_exponent(ock, P["coeffs"], P["shifted_blinding"], shift=sh)
                acc_v += sc * pow(point, sh, p) * v
            ch = ch * opening_challenge % p * opening_challenge % p
        # recover the witness exponent through the oracle (its point equals the GPU's, asserted above)
        comb, rcomb, ch = [], [], 1
        for P in group:
            comb = OK._axpy(comb, ch, P["coeffs"], p)
            rcomb = OK._axpy(rcomb, ch, P["blinding"] or [], p)
            if P["degree_bound"] is not None:
                sc = ch * opening_challenge % p
                comb = OK._axpy(comb, sc, [0] * (sup - P["degree_bound"]) + P["coeffs"], p)
                rcomb = OK._axpy(rcomb, sc, P["shifted_blinding"] or [], p)
            ch = ch * opening_challenge % p * opening_challenge % p
        w_exp = OK.commitment_exponent(ock, OK.poly_div_linear(comb, point, p), OK.poly_div_linear(rcomb, point, p) if any(rcomb) else None)
        assert ow == OK.exponent_point(ock, w_exp)
        assert OK.kzg_check_in_exponent(ock, acc_c % p, point, acc_v % p, w_exp, orv)
    ck.free()


@pytest.mark.parametrize("cid,build", [(BLS12_381, mini), (BN254, lambda cs: mimc(cs, 60))])
def test_gpu_indexer_matches_oracle(ctx, cid, build):
    """AHP::index (indexer.rs:71-116, arithmetic.rs:97-172) computed with the GPU primitives from CSR matrices
    == the oracle's index: squared matrices, domain sizes and all 7 x 3 evaluation tables."""
    fr = FR[cid]
    cs = OM.MarlinCS(fr.p)
    build(cs)
    ni, nv, nc = len(cs.input), len(cs.input) + len(cs.witness), len(cs.a)

    def csr(rows):
        ptr, cols, vals = [0], [], []
        for row in rows:
            for co, v in row:
                cols.append(v[1] if v[0] == "in" else ni + v[1])
                vals.append(co)
            ptr.append(len(cols))
        return CsrMatrix(np.asarray(ptr, dtype=np.uint32), np.asarray(cols, dtype=np.uint32), H.fr_array(cid, vals))

    a, b, c = csr(cs.a), csr(cs.b), csr(cs.c)          # before squaring: the device indexer squares them itself
    idx, extra = zm.index(ctx, cid, a, b, c, ni, nv)
    oidx = OM.index(cs, cid)
    assert extra == max(nc - nv, 0)
    assert (idx.num_constraints, idx.num_variables, idx.num_non_zeros) == (oidx["num_constraints"], oidx["num_variables"],
                                                                            oidx["num_non_zeros"])
    assert (idx.x_size, idx.h_size, idx.k_size, idx.b_size) == (oidx["dx"].size, oidx["dh"].size, oidx["dk"].size, oidx["db"].size)
    want = device_index(cid, oidx)
    for name in "abc":
        assert np.array_equal(idx.matrices[name].row_ptr, want.matrices[name].row_ptr)
        assert np.array_equal(idx.matrices[name].col_idx, want.matrices[name].col_idx)
        assert np.array_equal(idx.matrices[name].coeff, want.matrices[name].coeff)
        for key, arr in want.stars[name].items():
            assert np.array_equal(idx.stars[name][key], arr), (name, key)


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
@pytest.mark.parametrize("resident", [False, True])
def test_spmv_long_and_empty_rows(ctx, cid, resident):
    """zkb_spmv (ahp/prover.rs:110-123, 259-269) with the row shapes of Marlin's transposed matrices: one row far
    longer than the per-thread limit (the ONE variable), empty rows, coefficient 1 fast path, duplicate columns."""
    fr = FR[cid]
    p = fr.p
    rng = random.Random(cid + 17)
    n_cols = 700
    rows = [[(rng.randrange(p), rng.randrange(n_cols)) for _ in range(1500)],        # long row -> block reduction
            [], [(1, 5), (1, 5), (p - 1, 6)], [], [(rng.randrange(p), 699)],
            [(rng.randrange(p), rng.randrange(n_cols)) for _ in range(257)],          # just above the limit
            [(rng.randrange(p), rng.randrange(n_cols)) for _ in range(256)]]          # exactly at the limit
    x = [rng.randrange(p) for _ in range(n_cols)]
    ptr, cols, vals = [0], [], []
    for row in rows:
        for co, j in row:
            cols.append(j)
            vals.append(co)
        ptr.append(len(cols))
    m = CsrMatrix(np.asarray(ptr, dtype=np.uint32), np.asarray(cols, dtype=np.uint32), H.fr_array(cid, vals))
    xv = H.fr_array(cid, x)
    if resident:
        import torch
        dev = torch.device("cuda", 0)
        m.to_device(dev)
        xv = torch.from_numpy(xv.view(np.int64)).to(dev)
    got = H.fr_ints(cid, zm.to_host(ctx.spmv(cid, m, xv)))
    assert got == [sum(co * x[j] for co, j in row) % p for row in rows]
