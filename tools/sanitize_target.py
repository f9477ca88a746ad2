#!/usr/bin/env python3
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), touching every
kernel family of the prove path at sizes the tools finish in minutes: both MSM modes (window tables / plain windows),
the chunked big-bucket path (block-level tree sums), the bucket reductions, NTT passes in 1, 2 and 3 pass shapes,
witness_map, a whole Groth16 proof and the sharded fold.  Results are checked against the golden vectors so a
sanitizer-clean run is also a correct run.

    compute-sanitizer --tool racecheck python tools/sanitize_target.py          (see tools/sanitize.sh)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ckb_zkp_b200 import groth16 as zg, parallel, synth  # noqa: E402
from ckb_zkp_b200.backend import Context, CsrMatrix  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ctx = Context(0)
small = "--small" in sys.argv

# MSM goldens, both modes
for name, group in (("msm_bls12_381_g1_256", 1), ("msm_bls12_381_g2_64", 2), ("msm_bn254_g1_256", 1)):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    for pre in (True, False):
        srs = ctx.srs_upload(int(g["curve"]), group, g["bases_xy"], g["bases_inf"], precompute=pre)
        xy, inf = ctx.msm(srs, g["scalars"])
        assert inf == bool(g["result_inf"][0]) and np.array_equal(xy, g["result_xy"][0]), name
        srs.free()

# big-bucket path: half the scalars are 1 -> one bucket with n/2 entries (k_big_chunks / k_big_fold / block_sum)
n = 1024 if small else 4096
rng = np.random.default_rng(1)
k = synth.random_exponents(rng, n)
xy, inf = ctx.fixed_base_mul(1, 1, synth.generator_mont(1, 1), k)
sc = synth.random_exponents(rng, n)
sc[: n // 2] = 0
sc[: n // 2, 0] = 1
p = synth.FR_MODULUS[1]
e = sum(a * b for a, b in zip(synth.limbs_to_ints(sc), synth.limbs_to_ints(k))) % p
want = ctx.fixed_base_mul(1, 1, synth.generator_mont(1, 1), synth.ints_to_limbs([e]))
for pre in (True, False):
    srs = ctx.srs_upload(1, 1, xy, inf, precompute=pre)
    got = ctx.msm(srs, sc)
    assert not got[1] and np.array_equal(got[0], want[0][0]), "big-bucket msm"
    srs.free()

# sharded halves
recs = []
for rank in range(3):
    lo, hi = parallel.shard_range(n, 3, rank)
    srs = ctx.srs_upload_shard(1, 1, xy[lo:hi], inf[lo:hi], lo, n)
    recs.append(ctx.msm_partial(srs, sc))
    srs.free()
got = ctx.msm_fold(1, 1, np.stack(recs))
assert np.array_equal(got[0], want[0][0]), "sharded msm"

# NTT: golden 2^8 (one pass), round trips at 2^12 (two passes) and 2^21 (three passes; skipped with --small)
for name in ("ntt_bls12_381_2e8", "ntt_bn254_2e8"):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    keys = set(g.files)
    inp = g["input"] if "input" in keys else None
    if inp is not None:
        for variant, kw in (("fft", {}), ("ifft", {"inverse": True}), ("coset_fft", {"coset": True}),
                            ("coset_ifft", {"inverse": True, "coset": True})):
            if variant in keys:
                out = ctx.ntt(int(g["curve"]), inp.copy(), 8, **kw)
                assert np.array_equal(out, g[variant]), (name, variant)
for curve in (0, 1):
    for log_n in ((12,) if small else (12, 21)):
        a = rng.integers(0, 1 << 62, size=(1 << log_n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        b = ctx.ntt(curve, a.copy(), log_n, coset=True)
        b = ctx.ntt(curve, b, log_n, inverse=True, coset=True)
        assert np.array_equal(a, b), ("ntt round trip", curve, log_n)

# Groth16: golden proofs (whole and from three per-rank partials)
for name in ("groth16_mimc_bls12_381_2e6", "groth16_mini_bn254") + (() if small else ("groth16_mimc_bn254_2e10",)):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    q = lambda key: (g[key + "_xy"], g[key + "_inf"])
    s1, s2 = g["g1_singles"], g["g2_singles"]
    mk = lambda shard: zg.Parameters(ctx, int(g["curve"]), q("a_query"), q("b_g1_query"), q("b_g2_query"), q("h_query"),
                                     q("l_query"), s1[0], s1[1], s1[2], s2[0], s2[1], shard=shard)
    A, B, C = [CsrMatrix(g[w + "_ptr"], g[w + "_col"], g[w + "_val"]) for w in "abc"]
    args = (A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]), g["r"][0], g["s"][0])
    whole = mk(None)
    proofs = [ctx.groth16_prove(whole.pk, *args)]
    whole.free()
    keys_, recs = [], []
    for rank in range(3):
        keys_.append(mk((3, rank)))
        recs.append(ctx.groth16_prove_partial(keys_[-1].pk, *args))
    proofs.append(ctx.groth16_fold(keys_[0].pk, np.stack(recs), g["r"][0], g["s"][0]))
    for kk in keys_:
        kk.free()
    for proof in proofs:
        for key, got in (("proof_a", proof[0]), ("proof_b", proof[1]), ("proof_c", proof[2])):
            assert np.array_equal(g[key + "_xy"][0], got[0]), (name, key)
# round-2 kernels: batched MSM entry, the batched-affine chain kernels (switch), prefix products, point decompression
g = np.load(os.path.join(GOLD, "msm_bls12_381_g1_256.npz"))
srs = ctx.srs_upload(1, 1, g["bases_xy"], g["bases_inf"])
want_xy = g["result_xy"][0]
got = ctx.msm_batch([srs, srs, srs], [g["scalars"], g["scalars"][:100], g["scalars"]], [0, 3, 0])
assert np.array_equal(got[0][0], want_xy) and np.array_equal(got[2][0], want_xy)
os.environ["ZKB_MSM_BATCH"] = "1"
xy, inf = ctx.msm(srs, g["scalars"])
assert np.array_equal(xy, want_xy), "batched-affine chains"
os.environ["ZKB_MSM_BATCH"] = "0"
srs.free()
for curve in (0, 1):
    m = 1000 if small else 5000
    a = rng.integers(1, 1 << 62, size=(m, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    z = ctx.fr_prefix_product(curve, a)
    back = ctx.fr_vec_op(curve, Context.VEC_MUL, np.ascontiguousarray(z[:-1]), np.ascontiguousarray(a[:-1]))
    assert np.array_equal(back, z[1:]), "prefix product"
gen = synth.generator_mont(1, 1)
pts, _ = ctx.fixed_base_mul(1, 1, gen, synth.random_exponents(rng, 64))
canon = ctx.debug_fp_op(3, 5, pts.reshape(-1, 6).view(np.uint32).reshape(-1, 12), np.zeros((128, 12), dtype=np.uint32))   # from_mont
comp = np.ascontiguousarray(canon.reshape(64, 2, 12)[:, 0, :]).view(np.uint8).reshape(64, 48).copy()
ys = canon.reshape(64, 2, 12)[:, 1, :]
pq = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
for i in range(64):
    y = int.from_bytes(ys[i].tobytes(), "little")
    if y > pq - y:
        comp[i, 47] |= 0x80
dxy, dinf, dst = ctx.points_decompress(1, 1, comp, check_subgroup=True)
assert not dst.any() and not dinf.any() and np.array_equal(dxy, pts), "decompress"
# pairing kernels (local-memory heavy: one thread per Miller loop / per final exponentiation) and the thread-per-term
# path of zkb_msm_batch: e(aP, Q) e(-aP, Q) == 1 on both curves; 40 two-term MSMs against single calls
from ckb_zkp_b200 import pairing as zpair  # noqa: E402
for curve in (0, 1):
    g1, g2 = synth.generator_mont(curve, 1), synth.generator_mont(curve, 2)
    ks = synth.random_exponents(rng, 4)
    aP, _ = ctx.fixed_base_mul(curve, 1, g1, ks)
    bQ, _ = ctx.fixed_base_mul(curve, 2, g2, synth.random_exponents(rng, 4))
    neg = np.stack([zpair.neg_point(curve, 1, (aP[i], False))[0] for i in range(4)])
    gt = ctx.multi_pairing(curve, (np.stack([aP, neg], axis=1).reshape(8, -1), None), (np.repeat(bQ, 2, axis=0), None), 2)
    assert (gt == zpair.gt_one(curve)[None, :]).all(), "pairing product"
    single = ctx.multi_pairing(curve, (aP, None), (bQ, None), 1)
    assert len({single[i].tobytes() for i in range(4)}) == 4 and not (single == zpair.gt_one(curve)[None, :]).all(axis=1).any()
    srs = ctx.srs_upload(curve, 1, aP, None, precompute=False)
    scs = [synth.random_exponents(rng, 2) for _ in range(40)]
    many = ctx.msm_batch([srs] * 40, scs, [i % 3 for i in range(40)])
    for i in (0, 1, 2, 39):
        one = ctx.msm(srs, scs[i], base_offset=i % 3)
        assert many[i][1] == one[1] and np.array_equal(many[i][0], one[0]), "short-MSM batch"
    srs.free()
ctx.close()
print("sanitize target ok")
