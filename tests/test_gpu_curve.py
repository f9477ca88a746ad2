"""Curve::vartime_multiscalar_mul and its commitment loops (ckb_zkp_b200/curve.py, restating curve/src/lib.rs:38-45 and
spartan/src/commitments.rs:10-56) against the oracle's naive MSM."""
import random

import numpy as np
import pytest

from ckb_zkp_b200 import curve as zc
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR
from oracle.pyref.msm import msm_naive
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_vector_commitments_match_oracle(ctx, cid):
    c = CURVES[(cid, 1)]
    p = FR[cid].p
    rng = random.Random(cid)
    gens_pts = H.multiples(cid, 1, 64, start=101)
    h_pt = c.mul_affine(c.gen, 987654321)
    gxy, ginf = H.points_array(cid, 1, gens_pts)
    hxy, hinf = H.points_array(cid, 1, [h_pt])
    gens = zc.Generators(ctx, cid, (gxy, ginf), (hxy[0], bool(hinf[0])))
    values = [rng.randrange(p) for _ in range(64)]
    vm = H.fr_array(cid, values)
    # one commitment over a prefix (argument order of the reference: scalars first)
    got = zc.vartime_multiscalar_mul(gens, vm[:10])
    assert H.array_point(cid, 1, *got) == c.to_affine(msm_naive(c, gens_pts[:10], values[:10]))
    # poly_commit_vec: + blind * h
    blind = rng.randrange(p)
    got = zc.poly_commit_vec(gens, vm[:8], blind)
    want = c.to_affine(msm_naive(c, gens_pts[:8] + [h_pt], values[:8] + [blind]))
    assert H.array_point(cid, 1, *got) == want
    # packing_poly_commit: 64 values = 8 rows of 8, blinds drawn row by row
    for is_blind in (False, True):
        commits, blinds = zc.packing_poly_commit(gens, vm, random.Random(5), is_blind)
        draw = random.Random(5)
        assert blinds == [draw.randrange(p) if is_blind else 0 for _ in range(8)]
        for i, cm in enumerate(commits):
            want = c.to_affine(msm_naive(c, gens_pts[:8] + [h_pt], values[8 * i:8 * i + 8] + [blinds[i]]))
            assert H.array_point(cid, 1, *cm) == want
    with pytest.raises(AssertionError):
        zc.packing_poly_commit(gens, vm[:48], random.Random(1), False)
    gens.free()
