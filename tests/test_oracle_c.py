"""The C++ restatement (oracle/c/zkref.cpp: arkworks-0.2 algorithms, the CPU baseline) against the
first-principles Python oracle and the committed golden vectors.  CPU only."""
import os
import random

import numpy as np
import pytest

from oracle import cref
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR
from oracle.pyref.msm import msm_naive
from oracle.pyref.ntt import Domain
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["msm_bls12_381_g1_256", "msm_bn254_g1_256", "msm_bls12_381_g2_64", "msm_bn254_g2_64"])
def test_msm_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cid, group = int(g["curve"]), int(g["group"])
    for threads in (1, 4):
        xy, inf, _ = cref.msm(cid, group, g["bases_xy"], g["bases_inf"], g["scalars"], threads)
        assert inf == bool(g["result_inf"][0])
        assert np.array_equal(xy, g["result_xy"][0])


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)])
def test_msm_edge_cases(cid, group):
    c = CURVES[(cid, group)]
    rng = random.Random(5)
    pts = H.multiples(cid, group, 40, start=2)
    pts[7] = None
    r = c.r
    sc = [0, 1, r - 1, 1, 2, (1 << 64), (1 << 128) - 1] + [rng.randrange(r) for _ in range(33)]
    for n in (0, 1, 5, 31, 32, 40):     # crosses the c = 3 / c = log-based switch at n = 32
        xy, inf = H.points_array(cid, group, pts[:n])
        got, ginf, _ = cref.msm(cid, group, xy, inf, H.ints_to_u64(sc[:n], 4), 3)
        assert H.array_point(cid, group, got, ginf) == c.to_affine(msm_naive(c, pts[:n], sc[:n]))
    same = [pts[3]] * 16
    xy, inf = H.points_array(cid, group, same)
    got, ginf, _ = cref.msm(cid, group, xy, inf, H.ints_to_u64([12345] * 16, 4))
    assert H.array_point(cid, group, got, ginf) == c.mul_affine(pts[3], 16 * 12345)


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_fixed_base_mul(cid):
    rng = random.Random(9)
    for group in (1, 2):
        c = CURVES[(cid, group)]
        ks = [0, 1, c.r - 1] + [rng.randrange(c.r) for _ in range(10)]
        gxy, _ = H.points_array(cid, group, [c.gen])
        xy, inf = cref.fixed_base_mul(cid, group, gxy[0], H.ints_to_u64(ks, 4), 2)
        assert H.array_points(cid, group, xy, inf) == [c.mul_affine(c.gen, k) for k in ks]


@pytest.mark.parametrize("name", ["ntt_bls12_381_2e8", "ntt_bn254_2e8"])
def test_ntt_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cid, log_n = int(g["curve"]), int(g["log_n"])
    for key, kw in (("fft", {}), ("ifft", {"inverse": True}), ("coset_fft", {"coset": True}),
                    ("coset_ifft", {"inverse": True, "coset": True})):
        for threads in (1, 3):
            a = g["input"].copy()
            cref.ntt(cid, a, log_n, n_threads=threads, **kw)
            assert np.array_equal(a, g[key]), key


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_ntt_sizes(cid):
    fr = FR[cid]
    rng = random.Random(1)
    for log_n in (0, 1, 2, 5, 11):
        vals = [rng.randrange(fr.p) for _ in range(1 << log_n)]
        a = H.fr_array(cid, vals)
        cref.ntt(cid, a, log_n, coset=True, n_threads=4)
        assert H.fr_ints(cid, a) == Domain(fr, 1 << log_n).coset_fft(vals)


@pytest.mark.parametrize("name", ["groth16_mini_bls12_381", "groth16_mini_bn254", "groth16_mimc_bls12_381_2e6",
                                  "groth16_mimc_bn254_2e10"])
def test_groth16_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cid = int(g["curve"])
    mats = [(g[w + "_ptr"], g[w + "_col"], g[w + "_val"]) for w in "abc"]
    ni, na = int(g["n_inputs"]), int(g["n_aux"])
    h = cref.witness_map(cid, mats[0], mats[1], mats[2], g["z"], ni, 4)
    assert np.array_equal(h, g["h"])
    pk = {"a": (g["a_query_xy"], g["a_query_inf"]), "b1": (g["b_g1_query_xy"], g["b_g1_query_inf"]),
          "b2": (g["b_g2_query_xy"], g["b_g2_query_inf"]), "h": (g["h_query_xy"], g["h_query_inf"]),
          "l": (g["l_query_xy"], g["l_query_inf"]), "g1_singles": g["g1_singles"], "g2_singles": g["g2_singles"]}
    proof = cref.groth16_prove(cid, pk, mats[0], mats[1], mats[2], g["z"], ni, na, g["r"][0], g["s"][0], 4)
    for key, got in (("proof_a", proof[0]), ("proof_b", proof[1]), ("proof_c", proof[2])):
        assert got[1] == bool(g[key + "_inf"][0])
        assert np.array_equal(got[0], g[key + "_xy"][0]), key
