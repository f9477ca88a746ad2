"""Parity at the sizes BASELINE.json quotes, against the C++ restatement of the reference's CPU path
(oracle/c/zkref.cpp: ark-ec 0.2 Pippenger, ark-poly 0.2 radix-2 domain, groth16/src/prover.rs:124-228), bit for bit:

  configs[1]  Groth16 prove, BLS12-381, 2^20 constraints           (CPU: one full proof, ~12 s on 16 cores)
  configs[2]  G1 MSM, BLS12-381, 2^24 bases, full-width scalars    (CPU: 15 windows in parallel)
              + the boolean-heavy scalar mix of SURVEY.md 8d at 2^22 (giant "digit = 1" bucket path)
  configs[3]  Fr NTT 2^24, both fields, forward / inverse / coset variants

and a soak: the same MSM / proof repeated hundreds of times must return the same bytes every time (the bucket
reductions contain block-level trees; a shared-memory race would show up as run-to-run differences)."""
import numpy as np
import pytest

from ckb_zkp_b200 import synth
from oracle import cref

pytestmark = pytest.mark.gpu
BLS, BN = 1, 0


def _points(ctx, curve, group, k, chunk=1 << 20):
    gen = synth.generator_mont(curve, group)
    xs, infs = [], []
    for i in range(0, len(k), chunk):
        xy, inf = ctx.fixed_base_mul(curve, group, gen, k[i:i + chunk])
        xs.append(xy)
        infs.append(inf)
    return np.concatenate(xs), np.concatenate(infs)


def test_groth16_2e20_bls12_381_proof_is_bit_exact(ctx):
    """BASELINE configs[1]: the whole prove path at full size vs the restated reference prover on the same key"""
    n = 1 << 20
    inst = synth.MimcInstance(BLS, n)
    A, B, C, z = inst.device_form(ctx)
    key = synth.SyntheticKey(inst.n_inputs + inst.n_aux, inst.n_inputs, 2 * n, b_zero_cols=np.arange(4, 4 + n, 2))
    params = key.upload(ctx, BLS, keep_host=True)
    r, s = synth.ints_to_limbs([0x1234567])[0], synth.ints_to_limbs([0x89ABCDE])[0]
    proof = ctx.groth16_prove(params.pk, A, B, C, z, inst.n_inputs, inst.n_aux, r, s)
    params.free()
    hp = key.host_points
    pk = {k: hp[k] for k in ("a", "b1", "b2", "h", "l")}
    pk["g1_singles"], pk["g2_singles"] = hp["g1_singles"], hp["g2_singles"]
    mats = [(m.row_ptr, m.col_idx, m.coeff) for m in (A, B, C)]
    ref = cref.groth16_prove(BLS, pk, mats[0], mats[1], mats[2], z, inst.n_inputs, inst.n_aux, r, s)
    for name, got, want in zip("ABC", proof, ref):
        assert got[1] == want[1] and not got[1], name
        assert np.array_equal(got[0], want[0]), "proof." + name


def test_msm_2e24_bls12_381_g1_is_bit_exact(ctx):
    """BASELINE configs[2] on one GPU: 2^24 bases k_i * G, full-width scalars, vs ark's Pippenger restated"""
    n = 1 << 24
    rng = np.random.default_rng(24)
    k = synth.random_exponents(rng, n)
    sc = synth.random_exponents(rng, n)
    sc[5] = 0
    sc[6] = [1, 0, 0, 0]
    xy, inf = _points(ctx, BLS, 1, k)
    srs = ctx.srs_upload(BLS, 1, xy, inf)
    got = ctx.msm(srs, sc)
    srs.free()
    want_xy, want_inf, _ = cref.msm(BLS, 1, xy, inf, sc)
    assert got[1] == want_inf and np.array_equal(got[0], want_xy)


def test_msm_2e22_boolean_heavy_scalars_is_bit_exact(ctx):
    """SURVEY.md 8d adversarial mix: 50 % in {0, 1}, 25 % < 2^16, 25 % full width -- the chunked big-bucket path"""
    n = 1 << 22
    rng = np.random.default_rng(22)
    k = synth.random_exponents(rng, n)
    sc = synth.random_exponents(rng, n)
    u = rng.random(n)
    sc[u < 0.5, 1:] = 0
    sc[u < 0.5, 0] &= np.uint64(1)
    mid = (u >= 0.5) & (u < 0.75)
    sc[mid, 1:] = 0
    sc[mid, 0] &= np.uint64(0xFFFF)
    xy, inf = _points(ctx, BLS, 1, k)
    inf[::1000] = 1                                  # identity bases in the key (b_g1_query of a real key)
    srs = ctx.srs_upload(BLS, 1, xy, inf)
    got = ctx.msm(srs, sc)
    srs.free()
    want_xy, want_inf, _ = cref.msm(BLS, 1, xy, inf, sc)
    assert got[1] == want_inf and np.array_equal(got[0], want_xy)


@pytest.mark.parametrize("curve", [BLS, BN])
def test_ntt_2e24_is_bit_exact(ctx, curve):
    """BASELINE configs[3], largest size: fft, ifft, coset_fft, coset_ifft vs ark-poly's radix-2 domain restated"""
    log_n = 24
    rng = np.random.default_rng(100 + curve)
    host = rng.integers(0, 1 << 62, size=(1 << log_n, 4), dtype=np.uint64)
    host[:, 3] &= np.uint64((1 << 60) - 1)
    for inverse, coset in ((False, False), (True, False), (False, True), (True, True)):
        got = ctx.ntt(curve, host.copy(), log_n, inverse=inverse, coset=coset)
        want = cref.ntt(curve, host.copy(), log_n, inverse=inverse, coset=coset)
        assert np.array_equal(got, want), (curve, inverse, coset)


def test_soak_same_bytes_every_time(ctx):
    """500 x the same 2^14 G1 MSM (full-width and boolean-heavy scalars, both bucket-reduction shapes) and 200 x the same
    2^10-constraint proof: byte-identical results every time"""
    import os
    n = 1 << 14
    rng = np.random.default_rng(5)
    xy, inf = _points(ctx, BLS, 1, synth.random_exponents(rng, n))
    sc = synth.random_exponents(rng, n)
    sb = sc.copy()
    sb[: n // 2, 1:] = 0
    sb[: n // 2, 0] &= np.uint64(1)
    for precompute in (True, False):
        srs = ctx.srs_upload(BLS, 1, xy, inf, precompute=precompute)
        for scal in (sc, sb):
            first = ctx.msm(srs, scal)
            for i in range(250):
                again = ctx.msm(srs, scal)
                assert again[1] == first[1] and np.array_equal(again[0], first[0]), (precompute, i)
        srs.free()
    from ckb_zkp_b200 import groth16 as zg
    from ckb_zkp_b200.backend import CsrMatrix
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "groth16_mimc_bn254_2e10.npz"))
    q = lambda k: (g[k + "_xy"], g[k + "_inf"])
    s1, s2 = g["g1_singles"], g["g2_singles"]
    params = zg.Parameters(ctx, BN, q("a_query"), q("b_g1_query"), q("b_g2_query"), q("h_query"), q("l_query"), s1[0], s1[1],
                           s1[2], s2[0], s2[1])
    A, B, C = [CsrMatrix(g[w + "_ptr"], g[w + "_col"], g[w + "_val"]) for w in "abc"]
    ctx.groth16_stage(params.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]))
    for i in range(200):
        ctx.groth16_prove_staged(params.pk, g["r"][0], g["s"][0])
        p = ctx.groth16_fetch_proof(params.pk)
        for key, got in (("proof_a", p[0]), ("proof_b", p[1]), ("proof_c", p[2])):
            assert np.array_equal(g[key + "_xy"][0], got[0]), (key, i)
    params.free()
