"""Variable-base MSM on the GPU vs the oracle's restatement of ark-ec 0.2 VariableBaseMSM
(call sites groth16/src/prover.rs:187,190,220; marlin/src/pc/kzg10.rs:109; curve/src/lib.rs:44)."""
import random

import numpy as np
import pytest

from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR
from oracle.pyref.msm import msm_naive, msm_pippenger
from tests import helpers as H

pytestmark = pytest.mark.gpu

ALL = [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)]


def gpu_msm(ctx, cid, group, pts, scalars, precompute=True, base_offset=0, mont=False, n=None):
    xy, inf = H.points_array(cid, group, pts)
    srs = ctx.srs_upload(cid, group, xy, inf, precompute=precompute)
    try:
        sc = H.fr_array(cid, scalars, mont=mont)
        if n is not None:
            sc = sc[:n]
        out, is_inf = ctx.msm(srs, sc, base_offset=base_offset, mont=mont)
        return H.array_point(cid, group, out, is_inf)
    finally:
        srs.free()


def random_points(c, rng, n):
    return [c.mul_affine(c.gen, rng.randrange(1, c.r)) for _ in range(n)]


@pytest.mark.parametrize("cid,group", ALL)
@pytest.mark.parametrize("precompute", [True, False])
def test_msm_small_random(ctx, cid, group, precompute):
    c = CURVES[(cid, group)]
    rng = random.Random(cid * 100 + group)
    for n in (1, 2, 7, 33, 200):
        pts = H.multiples(cid, group, n, start=rng.randrange(1, 1 << 30))
        sc = [rng.randrange(c.r) for _ in range(n)]
        want = c.to_affine(msm_pippenger(c, pts, sc, FR[cid].bits))
        assert want == c.to_affine(msm_naive(c, pts, sc))
        assert gpu_msm(ctx, cid, group, pts, sc, precompute) == want


@pytest.mark.parametrize("cid,group", ALL)
def test_msm_edge_cases(ctx, cid, group):
    c = CURVES[(cid, group)]
    r = c.r
    rng = random.Random(7)
    pts = H.multiples(cid, group, 40, start=5)
    # empty input -> identity
    assert gpu_msm(ctx, cid, group, pts, [], n=0) is None
    # all-zero scalars -> identity
    assert gpu_msm(ctx, cid, group, pts, [0] * 40) is None
    # scalars 0 / 1 / r-1 / powers of two straddling window boundaries
    sc = [0, 1, r - 1, 2, 1 << 15, 1 << 16, (1 << 16) - 1, 1 << 31, 1 << 32, (1 << 64) - 1, 1 << 64, 1 << 127, 1 << 128,
          1 << 200, r - 2, (r - 1) // 2, (r + 1) // 2] + [rng.randrange(r) for _ in range(23)]
    want = c.to_affine(msm_naive(c, pts, sc))
    assert gpu_msm(ctx, cid, group, pts, sc, True) == want
    assert gpu_msm(ctx, cid, group, pts, sc, False) == want
    # identity bases inside the SRS (b_g1_query / b_g2_query are full of them, generator.rs:218-223)
    holes = list(pts)
    for i in (0, 3, 4, 17, 39):
        holes[i] = None
    want = c.to_affine(msm_naive(c, holes, sc))
    assert gpu_msm(ctx, cid, group, holes, sc, True) == want
    assert gpu_msm(ctx, cid, group, holes, sc, False) == want
    # the same point many times with the same scalar: every bucket addition is a doubling
    same = [pts[3]] * 40
    k = rng.randrange(r)
    assert gpu_msm(ctx, cid, group, same, [k] * 40) == c.mul_affine(pts[3], 40 * k % r)
    # P and -P with equal scalars cancel to the identity inside one bucket
    pm = [pts[1], c.neg_affine(pts[1])] * 8
    assert gpu_msm(ctx, cid, group, pm, [k] * 16) is None
    # result is exactly the identity although no bucket is: k*P + (r-k)*P
    assert gpu_msm(ctx, cid, group, [pts[2], pts[2]], [k, r - k]) is None


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BLS12_381, 1)])
def test_msm_offset_truncate_mont(ctx, cid, group):
    c = CURVES[(cid, group)]
    rng = random.Random(11)
    pts = H.multiples(cid, group, 64, start=9)
    sc = [rng.randrange(c.r) for _ in range(64)]
    # calculate_coeff skips query[0] (prover.rs:220): base_offset = 1
    want = c.to_affine(msm_naive(c, pts[1:], sc[:63]))
    assert gpu_msm(ctx, cid, group, pts, sc[:63], base_offset=1) == want
    # more scalars than bases: zip truncates (prover.rs:187 passes N scalars for N-1 bases)
    assert gpu_msm(ctx, cid, group, pts[:50], sc) == c.to_affine(msm_naive(c, pts[:50], sc[:50]))
    # fewer scalars than bases
    assert gpu_msm(ctx, cid, group, pts, sc[:10]) == c.to_affine(msm_naive(c, pts[:10], sc[:10]))
    # Curve::vartime_multiscalar_mul takes Montgomery-form scalars (curve/src/lib.rs:38-45)
    assert gpu_msm(ctx, cid, group, pts, sc, mont=True) == c.to_affine(msm_naive(c, pts, sc))


@pytest.mark.parametrize("cid,group,n", [(BN254, 1, 1 << 10), (BLS12_381, 1, 3000), (BLS12_381, 2, 600), (BN254, 2, 700)])
def test_msm_medium_boolean_heavy(ctx, cid, group, n):
    """Witness-shaped scalars: 50% in {0,1}, 25% < 2^16, 25% full width -- one bucket
    receives a large share of the entries."""
    c = CURVES[(cid, group)]
    rng = random.Random(n)
    pts = H.multiples(cid, group, n, start=3)
    sc = []
    for _ in range(n):
        u = rng.random()
        sc.append(rng.randrange(2) if u < 0.5 else rng.randrange(1 << 16) if u < 0.75 else rng.randrange(c.r))
    want = c.to_affine(msm_pippenger(c, pts, sc, FR[cid].bits))
    assert gpu_msm(ctx, cid, group, pts, sc, True) == want
    assert gpu_msm(ctx, cid, group, pts, sc, False) == want


def test_msm_linearity_large(ctx):
    """Size-independent property at 2^18 (BLS12-381 G1): bases are k_i*G with known k_i, so the
    result must be (sum s_i k_i mod r) * G; also MSM(s) + MSM(t) == MSM(s + t)."""
    cid, group, n = BLS12_381, 1, 1 << 18
    c = CURVES[(cid, group)]
    r = c.r
    rng = np.random.default_rng(5)
    ks = rng.integers(1, 1 << 62, size=n, dtype=np.uint64)
    k_arr = np.zeros((n, 4), dtype=np.uint64)
    k_arr[:, 0] = ks
    gen_xy, _ = H.points_array(cid, group, [c.gen])
    xy, inf = ctx.fixed_base_mul(cid, group, gen_xy[0], k_arr)
    assert not inf.any()
    srs = ctx.srs_upload(cid, group, xy, inf)
    s = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 61) - 1)           # < 2^253 < r
    t = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    t[:, 3] &= np.uint64((1 << 61) - 1)
    s_int, t_int = H.u64_to_ints(s), H.u64_to_ints(t)
    ks_int = [int(k) for k in ks]
    out, is_inf = ctx.msm(srs, s)
    got_s = H.array_point(cid, group, out, is_inf)
    assert got_s == c.mul_affine(c.gen, sum(a * b for a, b in zip(s_int, ks_int)) % r)
    out, is_inf = ctx.msm(srs, t)
    got_t = H.array_point(cid, group, out, is_inf)
    st = H.ints_to_u64([(a + b) % r for a, b in zip(s_int, t_int)], 4)
    out, is_inf = ctx.msm(srs, st)
    got_st = H.array_point(cid, group, out, is_inf)
    assert got_st == c.to_affine(c.add(c.from_affine(got_s), c.from_affine(got_t)))
    srs.free()


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BLS12_381, 2)])
@pytest.mark.parametrize("precompute", [True, False])
def test_msm_giant_bucket(ctx, cid, group, precompute):
    """Thousands of equal scalars land in one bucket: the chunked big-bucket path with several
    2048-entry chunks (and the fold over their partials) must agree with the CPU restatement."""
    from oracle import cref
    c = CURVES[(cid, group)]
    n = 7000 if group == 1 else 6000
    rng = np.random.default_rng(n)
    gen_xy, _ = H.points_array(cid, group, [c.gen])
    ks = np.zeros((n, 4), dtype=np.uint64)
    ks[:, 0] = rng.integers(1, 1 << 40, size=n, dtype=np.uint64)
    xy, inf = ctx.fixed_base_mul(cid, group, gen_xy[0], ks)
    sc = np.zeros((n, 4), dtype=np.uint64)
    sc[:, 0] = 1                                            # digit 1 of window 0 for most entries
    sc[5000:5600, 0] = 0x0007000700070007                   # the same non-trivial digit in several windows
    sc[5600:] = rng.integers(0, 1 << 62, size=(n - 5600, 4), dtype=np.uint64)
    sc[5600:, 3] &= np.uint64((1 << 60) - 1)
    srs = ctx.srs_upload(cid, group, xy, inf, precompute=precompute)
    got, ginf = ctx.msm(srs, sc)
    want, winf, _ = cref.msm(cid, group, xy, inf, sc)
    assert ginf == winf and np.array_equal(got, want)
    srs.free()


@pytest.mark.parametrize("levels", ["0", "1", "2", "7"])
@pytest.mark.parametrize("cid,group", [(BLS12_381, 1), (BN254, 2)])
def test_msm_pair_levels_forced(ctx, cid, group, levels, monkeypatch):
    """The batched-affine pair levels (csrc/msm_affine.cuh) for every level count from "XYZZ only" to
    "until every bucket holds one point", on inputs that hit each special case of the affine addition:
    doublings (equal bases, equal scalars), P + (-P), identity partial sums that meet again one level up,
    identity bases, odd leftovers, empty buckets, and one bucket holding most entries."""
    monkeypatch.setenv("ZKB_MSM_PAIR_LEVELS", levels)
    c = CURVES[(cid, group)]
    r = c.r
    rng = random.Random(int(levels) * 10 + group)
    base = H.multiples(cid, group, 24, start=11)
    k = rng.randrange(r)
    pts, sc = [], []
    pts += [base[0]] * 9;                                sc += [k] * 9                    # 9 equal entries: doublings + leftover
    pts += [base[1], c.neg_affine(base[1])] * 5;         sc += [k] * 10                   # cancelling pairs -> identities
    pts += [base[2], c.neg_affine(base[2]), base[2]];    sc += [k] * 3                    # identity + P one level up
    pts += [None, base[3], None];                        sc += [k, k, 5]                  # identity bases
    pts += base[4:];                                     sc += [rng.randrange(r) for _ in base[4:]]
    pts += base[4:14];                                   sc += [1] * 10                   # a heavily loaded bucket
    want = c.to_affine(msm_naive(c, pts, sc))
    assert gpu_msm(ctx, cid, group, pts, sc, True) == want
    assert gpu_msm(ctx, cid, group, pts, sc, False) == want
    # larger random instance against the Pippenger restatement
    n = 1500 if group == 1 else 400
    pts = H.multiples(cid, group, n, start=77)
    sc = [rng.randrange(r) if rng.random() < 0.8 else rng.randrange(3) for _ in range(n)]
    want = c.to_affine(msm_pippenger(c, pts, sc, FR[cid].bits))
    assert gpu_msm(ctx, cid, group, pts, sc, True) == want


@pytest.mark.parametrize("batch", ["0", "1"])
@pytest.mark.parametrize("cid,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)])
def test_msm_batched_affine_forced(ctx, cid, group, batch, monkeypatch):
    """The batched-affine bucket accumulation (csrc/msm_batch.cuh, ZKB_MSM_BATCH=1) against the XYZZ loop (=0) and the
    oracle, on inputs that hit every special case of an affine running sum: the same base many times with the same
    scalar (round 1 is a doubling, later rounds generic), P then -P (the running sum becomes the identity and the next
    entry has to restart it), identity bases, lists of very different lengths inside one thread, empty buckets."""
    monkeypatch.setenv("ZKB_MSM_BATCH", batch)
    c = CURVES[(cid, group)]
    r = c.r
    rng = random.Random(int(batch) * 10 + group + 100 * cid)
    base = H.multiples(cid, group, 40, start=5)
    k = rng.randrange(r)
    pts, sc = [], []
    pts += [base[0]] * 9;                                      sc += [k] * 9              # doubling, then generic additions
    pts += [base[1], c.neg_affine(base[1])] * 4 + [base[1]];   sc += [k] * 9              # identity running sums that restart
    pts += [base[2], c.neg_affine(base[2])];                   sc += [k] * 2              # a list that ends as the identity
    pts += [None, base[3], None];                              sc += [k, k, 5]            # identity bases
    pts += base[4:];                                           sc += [rng.randrange(r) for _ in base[4:]]
    pts += base[4:30];                                         sc += [1] * 26             # one heavily loaded bucket
    want = c.to_affine(msm_naive(c, pts, sc))
    assert gpu_msm(ctx, cid, group, pts, sc, True) == want
    assert gpu_msm(ctx, cid, group, pts, sc, False) == want
    # more lists than one warp owns, against the Pippenger restatement
    n = 3000 if group == 1 else 700
    pts = H.multiples(cid, group, n, start=1234)
    sc = [rng.randrange(r) if rng.random() < 0.8 else rng.randrange(3) for _ in range(n)]
    want = c.to_affine(msm_pippenger(c, pts, sc, FR[cid].bits))
    assert gpu_msm(ctx, cid, group, pts, sc, True) == want


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BLS12_381, 2)])
def test_msm_batch_many_short_msms(ctx, cid, group):
    """zkb_msm_batch with many short MSMs (>= 32 MSMs of <= 16 terms: the thread-per-term path a batch verifier's g_ic
    takes) against the oracle: lengths 0..16, offsets, a base at infinity, zero / one / r - 1 scalars, canonical and
    Montgomery scalars, a table-backed and a plain base set, host and device-resident scalars"""
    import torch
    c = CURVES[(cid, group)]
    rng = random.Random(90 + 7 * cid + group)
    n = 40
    pts = H.multiples(cid, group, n, start=3)
    pts[5] = None
    xy, inf = H.points_array(cid, group, pts)
    srs_a = ctx.srs_upload(cid, group, xy, inf)
    srs_b = ctx.srs_upload(cid, group, xy, inf, precompute=False)
    jobs = []
    for i in range(70):
        cnt = i % 17
        off = rng.randrange(0, n - cnt + 1)
        ks = [rng.choice([0, 1, c.r - 1, rng.randrange(c.r), rng.randrange(c.r)]) for _ in range(cnt)]
        jobs.append((srs_a if i % 3 else srs_b, off, ks))
    jobs.append((srs_a, n - 2, [rng.randrange(c.r) for _ in range(9)]))              # longer than the bases left: zip semantics
    for mont in (False, True):
        scalars = [H.fr_array(cid, ks, mont=mont) for _, _, ks in jobs]
        scalars[3] = torch.from_numpy(scalars[3].view(np.int64)).cuda()                # one device-resident scalar array
        got = ctx.msm_batch([s for s, _, _ in jobs], scalars, [o for _, o, _ in jobs], mont=mont)
        for (srs, off, ks), (gxy, ginf) in zip(jobs, got):
            use = pts[off:off + len(ks)]
            want = c.to_affine(msm_naive(c, use, ks[:len(use)])) if use else None
            assert H.array_point(cid, group, gxy, ginf) == want
    srs_a.free()
    srs_b.free()


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BLS12_381, 2)])
def test_msm_batch_matches_single_calls(ctx, cid, group):
    """zkb_msm_batch: k MSMs over slices of two resident SRS (different offsets and lengths, an empty one, Montgomery
    scalars) give exactly what k single zkb_msm_mont calls give, and the first one what the oracle gives."""
    c = CURVES[(cid, group)]
    rng = random.Random(7 * cid + group)
    n = 700 if group == 1 else 260
    pts = H.multiples(cid, group, n, start=31)
    xy, inf = H.points_array(cid, group, pts)
    srs_a = ctx.srs_upload(cid, group, xy, inf)
    srs_b = ctx.srs_upload(cid, group, xy[::-1].copy(), inf[::-1].copy(), precompute=False)
    shapes = [(srs_a, 0, n), (srs_b, 5, 100), (srs_a, n - 3, 10), (srs_b, 0, 0), (srs_a, 17, 1), (srs_a, 0, 300),
              (srs_b, 100, n), (srs_a, 1, 2), (srs_a, 200, 64)]
    scalars = [H.fr_array(cid, [rng.randrange(c.r) for _ in range(cnt)], mont=True) for _, _, cnt in shapes]
    got = ctx.msm_batch([s for s, _, _ in shapes], scalars, [o for _, o, _ in shapes], mont=True)
    for (srs, off, cnt), sc, (gxy, ginf) in zip(shapes, scalars, got):
        wxy, winf = ctx.msm(srs, sc, base_offset=off, mont=True) if cnt else (None, True)
        assert ginf == winf
        if not winf:
            assert np.array_equal(gxy, wxy)
    want = c.to_affine(msm_naive(c, pts, H.fr_ints(cid, scalars[0], mont=True)))
    assert H.array_point(cid, group, *got[0]) == want
    srs_a.free()
    srs_b.free()
