"""CPU tests of the oracle itself (no GPU): what pins it, given that the reference stores no
golden vector for this path.

 * the committed golden files are reproduced bit for bit (guards against silent oracle drift);
 * Pippenger (ark-ec 0.2 restatement) == naive double-and-add; NTT == O(n^2) DFT;
 * the Groth16 prover output equals the in-the-exponent derivation from the toxic waste, which is
   an independent statement of what a valid proof is (the algebra the reference's acceptance test
   `verify_proof == true`, groth16/tests/mini.rs:89, checks through pairings).
"""
import os
import random

import numpy as np
import pytest

from oracle.pyref import groth16 as OG
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR, stream_field
from oracle.pyref.msm import ark_window, msm_naive, msm_pippenger
from oracle.pyref.ntt import Domain, dft_naive
from oracle.pyref.r1cs import ConstraintSystem, mimc_circuit, mini_circuit
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_ark_window_rule():
    # n=2^10 -> 8, 2^18 -> 14, 2^20 -> 15, 2^21 -> 16, 2^24 -> 18; n < 32 -> 3
    assert [ark_window(1 << k) for k in (10, 18, 20, 21, 24)] == [8, 14, 15, 16, 18]
    assert ark_window(31) == 3 and ark_window(32) == 5


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)])
def test_curve_sanity(cid, group):
    c = CURVES[(cid, group)]
    assert c.on_curve(c.gen)
    assert c.to_affine(c.mul(c.from_affine(c.gen), c.r)) is None
    rng = random.Random(1)
    pts = H.multiples(cid, group, 12, start=3)
    assert all(c.on_curve(P) for P in pts)
    sc = [0, 1, c.r - 1] + [rng.randrange(c.r) for _ in range(9)]
    assert c.to_affine(msm_pippenger(c, pts, sc, FR[cid].bits)) == c.to_affine(msm_naive(c, pts, sc))


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_ntt_vs_dft(cid):
    fr = FR[cid]
    rng = random.Random(2)
    vals = [rng.randrange(fr.p) for _ in range(32)]
    d = Domain(fr, 32)
    assert pow(d.group_gen, 32, fr.p) == 1 and pow(d.group_gen, 16, fr.p) != 1
    assert d.fft(vals) == dft_naive(vals, d.group_gen, fr.p)
    assert d.ifft(d.fft(vals)) == vals
    assert d.coset_ifft(d.coset_fft(vals)) == vals
    with pytest.raises(ValueError):
        Domain(fr, (1 << fr.two_adicity) + 1)


@pytest.mark.parametrize("name", ["ntt_bls12_381_2e8", "ntt_bn254_2e8"])
def test_golden_ntt(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cid = int(g["curve"])
    d = Domain(FR[cid], 1 << int(g["log_n"]))
    vals = H.fr_ints(cid, g["input"])
    assert vals[:3] == [stream_field(4, i, FR[cid].p) for i in range(3)]
    assert H.fr_ints(cid, g["fft"]) == d.fft(vals)
    assert H.fr_ints(cid, g["coset_ifft"]) == d.coset_ifft(vals)


@pytest.mark.parametrize("name", ["msm_bls12_381_g1_256", "msm_bn254_g1_256", "msm_bls12_381_g2_64", "msm_bn254_g2_64"])
def test_golden_msm(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cid, group = int(g["curve"]), int(g["group"])
    c = CURVES[(cid, group)]
    pts = H.array_points(cid, group, g["bases_xy"], g["bases_inf"])
    sc = H.u64_to_ints(g["scalars"])
    want = H.array_point(cid, group, g["result_xy"][0], g["result_inf"][0])
    assert c.to_affine(msm_naive(c, pts, sc)) == want


@pytest.mark.parametrize("name,cid,build", [
    ("groth16_mini_bls12_381", BLS12_381, lambda cs: mini_circuit(cs)),
    ("groth16_mimc_bls12_381_2e6", BLS12_381, lambda cs: mimc_circuit(cs, 64)),
])
def test_golden_groth16(name, cid, build):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fr = FR[cid]
    cs = ConstraintSystem(fr.p)
    build(cs)
    assert cs.is_satisfied()
    seed = 1 if "mini" in name else 2
    alpha, beta, gamma, delta, t, r, s = [stream_field(seed, i, fr.p) for i in range(7)]
    pk = OG.generate_parameters(cs, cid, alpha, beta, gamma, delta, t)
    proof = OG.create_proof(pk, cs, r, s)
    assert proof == OG.proof_in_exponent(pk, cs, r, s)
    for key, group, P in (("proof_a", 1, proof[0]), ("proof_b", 2, proof[1]), ("proof_c", 1, proof[2])):
        xy, inf = H.points_array(cid, group, [P])
        assert np.array_equal(xy, g[key + "_xy"]) and np.array_equal(inf, g[key + "_inf"])
    h = OG.witness_map(cs, cid)
    assert h[-1] == 0
    assert np.array_equal(H.fr_array(cid, h, mont=False), g["h"])
    # r = 0 guard (prover.rs:170) and s = 0
    for rr, ss in ((0, 5), (5, 0), (0, 0)):
        assert OG.create_proof(pk, cs, rr, ss) == OG.proof_in_exponent(pk, cs, rr, ss)


def test_kzg10_oracle_self_consistency():
    """oracle/pyref/kzg10.py: commitments and opening witnesses equal the trapdoor-side exponent
    formula, and the restated KZG10::check accepts / rejects (marlin/src/pc/kzg10.rs:158-172)."""
    from oracle.pyref import kzg10 as K
    cid = BLS12_381
    mod = FR[cid].p
    rng = random.Random(1)
    pp = K.setup(cid, 10, rng.randrange(mod), g_scalar=rng.randrange(1, mod), gamma=rng.randrange(1, mod))
    ck = K.trim(pp, 6)
    p = [0, rng.randrange(mod), 0, rng.randrange(mod), rng.randrange(mod)]
    bl = [rng.randrange(mod), rng.randrange(mod)]
    c = K.kzg_commit(cid, ck["powers_of_g"], ck["powers_of_gamma_g"], p, bl, 6)
    ce = K.commitment_exponent(ck, p, bl)
    assert c == K.exponent_point(ck, ce)
    z = rng.randrange(mod)
    w, rv = K.kzg_open(cid, ck["powers_of_g"], ck["powers_of_gamma_g"], p, z, bl)
    we = K.commitment_exponent(ck, K.poly_div_linear(p, z, mod), K.poly_div_linear(bl, z, mod))
    assert w == K.exponent_point(ck, we)
    assert K.kzg_check_in_exponent(ck, ce, z, K.poly_eval(p, z, mod), we, rv)
    assert not K.kzg_check_in_exponent(ck, ce, z, (K.poly_eval(p, z, mod) + 1) % mod, we, rv)
    q = K.poly_div_linear(p, z, mod)
    x = rng.randrange(mod)
    assert ((x - z) * K.poly_eval(q, x, mod) + K.poly_eval(p, z, mod)) % mod == K.poly_eval(p, x, mod)
    with pytest.raises(K.KzgError):
        K.kzg_commit(cid, ck["powers_of_g"], ck["powers_of_gamma_g"], [5], None, 6)
    # degree-bounded polynomial: shifted commitment exponent carries beta^(supported - bound)
    out = K.pc_commit(ck, [{"coeffs": p, "degree_bound": 4, "blinding": bl, "shifted_blinding": bl}])
    assert out[0][1] == K.exponent_point(ck, K.commitment_exponent(ck, p, bl, shift=2))
