"""CPU test of the Marlin host orchestration (ckb_zkp_b200/marlin.py) over a mock backend whose
primitives are computed by the oracle: checks that the three AHP rounds call the device primitives
with the right operands, independently of the CUDA kernels (those are checked in test_gpu_marlin.py)."""
import random

import pytest

from ckb_zkp_b200 import marlin as zm
from oracle.pyref import marlin as OM
from oracle.pyref.fields import BLS12_381, BN254, FR
from tests import helpers as H
from tests.mock_backend import MockContext
from tests.test_gpu_marlin import ReplayRng, device_index, mimc, mini, outside


@pytest.mark.parametrize("cid,build", [(BLS12_381, mini), (BN254, lambda cs: mimc(cs, 12))])
def test_rounds_over_mock_backend(cid, build):
    p = FR[cid].p
    rng = random.Random(31)
    cs = OM.MarlinCS(p)
    build(cs)
    oidx = OM.index(cs, cid)
    ost = OM.prover_init(oidx, cs)
    idx = device_index(cid, oidx)
    st = zm.prover_init(MockContext(), idx, H.fr_array(cid, cs.input), H.fr_array(cid, cs.witness))
    Hs = idx.h_size
    draws = [rng.randrange(p) for _ in range(3 + 3 * Hs)]
    want = OM.prover_first_round(ost, draws[0], draws[1], draws[2], draws[3:])
    got = zm.prover_first_round(st, ReplayRng(draws))
    alpha, etas = outside(oidx["dh"], rng, p), [rng.randrange(p) for _ in range(3)]
    want.update(OM.prover_second_round(ost, alpha, *etas))
    got += zm.prover_second_round(st, alpha, *etas)
    beta = outside(oidx["dh"], rng, p)
    want.update(OM.prover_third_round(ost, beta))
    got += zm.prover_third_round(st, beta)
    polys = {label: H.fr_ints(cid, poly) for label, poly, _, _ in got}
    for label in want:
        assert polys[label] == want[label], label
    assert OM.verifier_equality_check(oidx, cs.input[1:], polys, alpha, *etas, beta, rng.randrange(p))


def test_marlin_oracle_rejects_a_bad_witness():
    """the restated verifier_equality_check (ahp/verifier.rs:128-209) is a real test: it fails when the
    witness does not satisfy the constraints or a prover polynomial is altered"""
    cid = BLS12_381
    p = FR[cid].p
    rng = random.Random(2)
    for bad in (False, True):
        cs = OM.MarlinCS(p)
        mini(cs)
        if bad:
            cs.witness[0] = 5                       # x = 5: 5 * (3 + 2) != 10
        idx = OM.index(cs, cid)
        st = OM.prover_init(idx, cs)
        Hs = idx["dh"].size
        polys = OM.prover_first_round(st, *[rng.randrange(p) for _ in range(3)], [rng.randrange(p) for _ in range(3 * Hs)])
        alpha, etas = outside(idx["dh"], rng, p), [rng.randrange(p) for _ in range(3)]
        polys.update(OM.prover_second_round(st, alpha, *etas))
        beta = outside(idx["dh"], rng, p)
        polys.update(OM.prover_third_round(st, beta))
        ok = OM.verifier_equality_check(idx, cs.input[1:], polys, alpha, *etas, beta, rng.randrange(p))
        assert ok == (not bad)


def _csr_from_rows(cid, rows, ni):
    import numpy as np
    from ckb_zkp_b200.backend import CsrMatrix
    ptr, cols, vals = [0], [], []
    for row in rows:
        for co, v in row:
            cols.append(v[1] if v[0] == "in" else ni + v[1])
            vals.append(co)
        ptr.append(len(cols))
    return CsrMatrix(np.asarray(ptr, dtype=np.uint32), np.asarray(cols, dtype=np.uint32), H.fr_array(cid, vals))


@pytest.mark.parametrize("shape", ["a_denser", "b_denser", "more_constraints", "more_variables"])
def test_matrix_preprocessing_matches_the_oracle(shape):
    """make_matrices_square + balance_matrices + the per-row column sort (constraint_systems.rs:9-31,83-114) on the
    CSR form == the oracle's list-of-rows restatement, for every branch: A denser than B (rows swapped until the
    densities cross), B denser (nothing happens), more constraints than variables (padding variables), more
    variables than constraints (empty rows), duplicate and unsorted columns inside a row."""
    import numpy as np
    cid = BN254
    p = FR[cid].p
    rng = random.Random({"a_denser": 1, "b_denser": 2, "more_constraints": 3, "more_variables": 4}[shape])
    cs = OM.MarlinCS(p)
    n_cons = {"more_constraints": 40, "more_variables": 6}.get(shape, 20)
    n_vars = {"more_constraints": 5, "more_variables": 30}.get(shape, 18)
    vs = [cs.alloc(rng.randrange(p)) for _ in range(n_vars)] + [cs.alloc_input(rng.randrange(p))] + [("in", 0)]
    lc = lambda k: [(rng.randrange(1, p), rng.choice(vs)) for _ in range(k)]
    for i in range(n_cons):
        ka, kb = (rng.randrange(3, 7), rng.randrange(0, 3)) if shape == "a_denser" else (rng.randrange(0, 3), rng.randrange(2, 6))
        cs.enforce(lc(ka), lc(kb), lc(rng.randrange(0, 4)))
    ni, nv = len(cs.input), len(cs.input) + len(cs.witness)
    mats = [_csr_from_rows(cid, rows, ni) for rows in (cs.a, cs.b, cs.c)]
    (a, b, c), extra = zm.make_matrices_square(mats, nv)
    a, b = zm.balance_matrices(a, b)
    got = [zm.sort_rows_by_column(m) for m in (a, b, c)]
    cs.make_matrices_square()
    assert extra == len(cs.input) + len(cs.witness) - nv
    for m, want_rows in zip(got, cs.matrices()):
        assert m.n_rows == len(want_rows)
        ptr = [0]
        for row in want_rows:
            ptr.append(ptr[-1] + len(row))
        assert m.row_ptr.tolist() == ptr
        assert m.col_idx.tolist() == [j for row in want_rows for _, j in row]
        assert np.array_equal(m.coeff, H.fr_array(cid, [co for row in want_rows for co, _ in row]))
