"""Conversions between the oracle's Python ints / tuples and the array layouts of the C ABI."""
import numpy as np

from oracle.pyref.curves import CURVES
from oracle.pyref.fields import FQ, FR

M64 = (1 << 64) - 1


def ints_to_u64(vals, limbs):
    out = np.zeros((len(vals), limbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(limbs):
            out[i, j] = (v >> (64 * j)) & M64
    return out


def u64_to_int(row):
    r = 0
    for j, w in enumerate(row):
        r |= int(w) << (64 * j)
    return r


def u64_to_ints(arr):
    return [u64_to_int(row) for row in arr]


def fr_array(curve_id, vals, mont=True):
    fr = FR[curve_id]
    return ints_to_u64([fr.to_mont(v % fr.p) if mont else v % fr.p for v in vals], 4)


def fr_ints(curve_id, arr, mont=True):
    fr = FR[curve_id]
    return [fr.from_mont(v) if mont else v for v in u64_to_ints(arr)]


def _coords(group, P):
    return list(P[0]) + list(P[1]) if group == 2 else [P[0], P[1]]


def points_array(curve_id, group, pts):
    """affine points (None = identity) -> (xy uint64[n, words] Montgomery, inf uint8[n]).
    Identity rows carry ark's (0, 1) placeholder coordinates to prove the flag is what counts."""
    fq = FQ[curve_id]
    L = fq.limbs
    nc = 4 if group == 2 else 2
    xy = np.zeros((len(pts), nc * L), dtype=np.uint64)
    inf = np.zeros(len(pts), dtype=np.uint8)
    for i, P in enumerate(pts):
        if P is None:
            inf[i] = 1
            cs = [0] * nc
            cs[nc // 2] = 1                      # y = 1 (ark GroupAffine::zero())
        else:
            cs = _coords(group, P)
        for k, c in enumerate(cs):
            m = fq.to_mont(c)
            for j in range(L):
                xy[i, k * L + j] = (m >> (64 * j)) & M64
    return xy, inf


def array_point(curve_id, group, xy, is_inf):
    """one affine point from the ABI layout -> oracle tuple / None."""
    if is_inf:
        return None
    fq = FQ[curve_id]
    L = fq.limbs
    cs = [fq.from_mont(u64_to_int(xy[k * L:(k + 1) * L])) for k in range(len(xy) // L)]
    if group == 2:
        return ((cs[0], cs[1]), (cs[2], cs[3]))
    return (cs[0], cs[1])


def array_points(curve_id, group, xy, inf):
    return [array_point(curve_id, group, xy[i], inf[i]) for i in range(len(inf))]


def multiples(curve_id, group, n, start=1):
    """[start*G, (start+1)*G, ...] as affine points."""
    c = CURVES[(curve_id, group)]
    G = c.from_affine(c.gen)
    acc = c.mul(G, start)
    out = []
    for _ in range(n):
        out.append(acc)
        acc = c.add_mixed(acc, c.gen)
    return c.batch_to_affine(out)
