"""`zkp_curve::Curve::vartime_multiscalar_mul` (curve/src/lib.rs:38-45) and its commitment-loop consumers on the B200
backend.

The reference's other schemes (Spartan, Hyrax, Libra, Bulletproofs, aSVC) commit vector by vector through this one
function, many small or medium MSMs over the SAME generators.  Here the generators are made resident once
(`Generators`, a cached `zkb_srs`) and a whole loop of commitments is one `zkb_msm_batch` call whose sorts, bucket
accumulations and reductions overlap on the library's side streams.

    vartime_multiscalar_mul(gens, scalars_mont)          one commitment; note the reference's argument order is
                                                         (scalars, points)
    poly_commit_vec / packing_poly_commit                spartan/src/commitments.rs:10-56
"""
import numpy as np

from . import _lib
from .backend import point_words
from .groth16 import FR_MODULUS
from .r1cs import ints_to_limbs


class Generators:
    """`&[G::Affine]` made resident: the bases of a Pedersen-style vector commitment plus the blinding base h as the last
    base, so that `MSM(values, generators) + blind * h` (commitments.rs:48-50) is ONE sum"""

    def __init__(self, ctx, curve, generators, h, group=_lib.G1):
        """generators: (xy uint64[n, words], inf uint8[n]); h: (xy uint64[words], is_identity)"""
        self.ctx, self.curve, self.group = ctx, curve, group
        xy, inf = generators
        self.n = len(inf)
        w = point_words(curve, group)
        all_xy = np.concatenate([np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, w),
                                 np.ascontiguousarray(h[0], dtype=np.uint64).reshape(1, w)])
        all_inf = np.concatenate([np.ascontiguousarray(inf, dtype=np.uint8), np.array([1 if h[1] else 0], dtype=np.uint8)])
        # the commitments use short prefixes of a long generator vector: window tables sized for the whole vector would
        # make every small MSM pay the big bucket set, so the table is built without the window copies
        self.srs = ctx.srs_upload(curve, group, all_xy, all_inf, precompute=False)
        self.h_srs = ctx.srs_upload(curve, group, all_xy[-1:], all_inf[-1:], precompute=False)

    def free(self):
        self.srs.free()
        self.h_srs.free()


def vartime_multiscalar_mul(gens, scalars_mont):
    """Curve::vartime_multiscalar_mul(scalars, points[..scalars.len()]) -> (xy, is_identity), canonical affine
    (`into_repr` of the scalars is fused on the device: zkb_msm_mont)"""
    return gens.ctx.msm(gens.srs, scalars_mont, mont=True)


def _sum_with_blinds(gens, commits, blinds_mont):
    """commit_i + blind_i * h for all i in two batched calls (the h multiples, then the pairwise sums)"""
    ctx = gens.ctx
    k = len(commits)
    hs = ctx.msm_batch([gens.h_srs] * k, [np.ascontiguousarray(blinds_mont[i:i + 1]) for i in range(k)], mont=True)
    xy = np.stack([p[0] for pair in zip(commits, hs) for p in pair])
    inf = np.array([1 if p[1] else 0 for pair in zip(commits, hs) for p in pair], dtype=np.uint8)
    tmp = ctx.srs_upload(gens.curve, gens.group, xy, inf, precompute=False)
    try:
        ones = np.zeros((2, 4), dtype=np.uint64)
        ones[:, 0] = 1
        return ctx.msm_batch([tmp] * k, [ones] * k, [2 * i for i in range(k)])
    finally:
        tmp.free()


def poly_commit_vec(gens, values_mont, blind):
    """spartan/src/commitments.rs:42-56: MSM(values, generators) + blind * h.  blind: canonical int"""
    values = np.ascontiguousarray(values_mont, dtype=np.uint64).reshape(-1, 4)
    commit = vartime_multiscalar_mul(gens, values)
    blind_m = gens.ctx.fr_convert(gens.curve, ints_to_limbs([blind % FR_MODULUS[gens.curve]]), to_mont=True)
    return _sum_with_blinds(gens, [commit], blind_m)[0]


def packing_poly_commit(gens, values_mont, rng, is_blind):
    """spartan/src/commitments.rs:10-40: the value vector as an l_size x r_size matrix, one commitment per row; blinds
    are drawn row by row in the reference's order.  -> ([(xy, is_identity)], [blind ints]); all rows go to the device
    in one zkb_msm_batch call"""
    ctx, curve = gens.ctx, gens.curve
    p = FR_MODULUS[curve]
    values = np.ascontiguousarray(values_mont, dtype=np.uint64).reshape(-1, 4)
    n = len(values)
    size = max(n - 1, 0).bit_length()                      # ark_std::log2 = ceil(log2 n)
    l_size, r_size = 1 << (size // 2), 1 << (size - size // 2)
    if n != l_size * r_size:
        raise AssertionError("the number of values must be a power of two")
    blinds = [rng.randrange(p) if is_blind else 0 for _ in range(l_size)]
    rows = [np.ascontiguousarray(values[i * r_size:(i + 1) * r_size]) for i in range(l_size)]
    commits = ctx.msm_batch([gens.srs] * l_size, rows, mont=True)
    if is_blind:
        commits = _sum_with_blinds(gens, commits, ctx.fr_convert(curve, ints_to_limbs(blinds), to_mont=True))
    return commits, blinds
