"""Groth16 setup on the GPU (ckb_zkp_b200/generator.py, restating groth16/src/generator.rs:19-34,135-286) against the
oracle's generate_parameters on the same toxic waste, and the reference's own acceptance test
(groth16/tests/mini.rs:46-97): generate -> prove -> verify with the pairing check."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import generator as zgen
from ckb_zkp_b200 import groth16 as zg
from ckb_zkp_b200.r1cs import ONE
from oracle.pyref import groth16 as OG
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR, stream_field
from oracle.pyref.r1cs import MIMC_SEED, ConstraintSystem, mimc_circuit, mini_circuit
from tests import helpers as H

pytestmark = pytest.mark.gpu


class MiniCircuit:
    """groth16/tests/mini.rs:12-44"""

    def __init__(self, x, y, z, num):
        self.x, self.y, self.z, self.num = x, y, z, num

    def generate_constraints(self, cs):
        vx = cs.alloc(lambda: self.x)
        vy = cs.alloc(lambda: self.y)
        vz = cs.alloc_input(lambda: self.z)
        for _ in range(self.num):
            cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])


class MimcCircuit:
    """the MiMC chain of SURVEY.md 8d, same allocation order as oracle.pyref.r1cs.mimc_circuit"""

    def __init__(self, p, n_constraints, seed=MIMC_SEED):
        self.p, self.n, self.seed = p, n_constraints, seed

    def generate_constraints(self, cs):
        p = self.p
        xl_v, xr_v = stream_field(self.seed, 0, p), stream_field(self.seed, 1, p)
        xl, xr = cs.alloc(lambda: xl_v), cs.alloc(lambda: xr_v)
        rounds = self.n // 2
        for i in range(rounds):
            c = stream_field(self.seed, 2 + i, p)
            tmp_v = (xl_v + c) * (xl_v + c) % p
            tmp = cs.alloc(lambda: tmp_v)
            cs.enforce([(1, xl), (c, ONE)], [(1, xl), (c, ONE)], [(1, tmp)])
            new_v = ((xl_v + c) * tmp_v + xr_v) % p
            new = cs.alloc_input(lambda: new_v) if i == rounds - 1 else cs.alloc(lambda: new_v)
            cs.enforce([(1, tmp)], [(1, xl), (c, ONE)], [(1, new), (p - 1, xr)])
            xr, xr_v = xl, xl_v
            xl, xl_v = new, new_v


def same_points(cid, group, got, want_pts):
    xy, inf = H.points_array(cid, group, want_pts)
    gxy, ginf = got
    assert np.array_equal(np.asarray(ginf, dtype=np.uint8), inf)
    keep = inf == 0
    assert np.array_equal(np.asarray(gxy)[keep], xy[keep])


@pytest.mark.parametrize("cid,which", [(BLS12_381, "mini"), (BN254, "mini"), (BLS12_381, "mimc"), (BN254, "mimc")])
def test_generate_parameters_matches_oracle(ctx, cid, which):
    fr = FR[cid]
    if which == "mini":
        circuit, cs = MiniCircuit(2, 3, 10, 10), mini_circuit(ConstraintSystem(fr.p))
    else:
        circuit, cs = MimcCircuit(fr.p, 64), mimc_circuit(ConstraintSystem(fr.p), 64)
    alpha, beta, gamma, delta, t = [stream_field(7, i, fr.p) for i in range(5)]
    want = OG.generate_parameters(cs, cid, alpha, beta, gamma, delta, t)
    got = zgen.generate_parameters(ctx, cid, circuit, alpha, beta, gamma, delta, t)
    same_points(cid, 1, got.a_query, want.a_query)
    same_points(cid, 1, got.b_g1_query, want.b_g1_query)
    same_points(cid, 2, got.b_g2_query, want.b_g2_query)
    same_points(cid, 1, got.h_query, want.h_query)
    same_points(cid, 1, got.l_query, want.l_query)
    same_points(cid, 1, got.vk.gamma_abc_g1, want.gamma_abc_g1)
    one = lambda pt: ([pt[0]], [1 if pt[1] else 0])
    same_points(cid, 1, one(got.vk.alpha_g1), [want.alpha_g1])
    same_points(cid, 1, one(got.beta_g1), [want.beta_g1])
    same_points(cid, 1, one(got.delta_g1), [want.delta_g1])
    same_points(cid, 2, one(got.vk.beta_g2), [want.beta_g2])
    same_points(cid, 2, one(got.vk.gamma_g2), [want.gamma_g2])
    same_points(cid, 2, one(got.vk.delta_g2), [want.delta_g2])


def test_generate_prove_verify_mini(ctx):
    """groth16/tests/mini.rs:46-97 end to end on the GPU, accepted by the restated pairing verifier
    (groth16/src/verifier.rs:18-44); a wrong public input is rejected."""
    cid = BLS12_381
    fr = FR[cid]
    rng = random.Random(2024)
    circuit = MiniCircuit(2, 3, 10, 10)
    data = zgen.generate_random_parameters(ctx, cid, circuit, rng)
    params = data.upload(ctx)
    proof = zg.create_random_proof(params, circuit, rng)
    params.free()
    # the oracle's verifier works on its own Parameters object: fill it from the GPU-made key
    pk = OG.Parameters()
    pk.curve_id = cid
    pt1 = lambda x: H.array_point(cid, 1, x[0], x[1])
    pt2 = lambda x: H.array_point(cid, 2, x[0], x[1])
    pk.alpha_g1, pk.beta_g2 = pt1(data.vk.alpha_g1), pt2(data.vk.beta_g2)
    pk.gamma_g2, pk.delta_g2 = pt2(data.vk.gamma_g2), pt2(data.vk.delta_g2)
    pk.gamma_abc_g1 = H.array_points(cid, 1, *data.vk.gamma_abc_g1)
    oproof = (pt1(proof.a), pt2(proof.b), pt1(proof.c))
    assert OG.verify_proof(pk, oproof, [10])
    assert not OG.verify_proof(pk, oproof, [11])


def test_lagrange_coefficients_sum_to_one(ctx):
    """sum_i L_i(t) == 1 and sum_i w^i L_i(t) == t (interpolation of 1 and of x) at a size the oracle does not reach"""
    cid = BN254
    fr = FR[cid]
    t = stream_field(9, 0, fr.p)
    log_m = 16
    u, zt = zgen.lagrange_coefficients(ctx, cid, log_m, t)
    vals = H.fr_ints(cid, u)
    assert sum(vals) % fr.p == 1
    assert zt == (pow(t, 1 << log_m, fr.p) - 1) % fr.p
    w = pow(pow(zgen.FR_GENERATOR[cid], (fr.p - 1) >> zgen.TWO_ADICITY[cid], fr.p), 1 << (zgen.TWO_ADICITY[cid] - log_m), fr.p)
    acc, wi = 0, 1
    for v in vals:
        acc = (acc + wi * v) % fr.p
        wi = wi * w % fr.p
    assert acc == t
