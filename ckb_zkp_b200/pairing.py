"""Host side of the batched pairing check (zkb_multi_pairing, csrc/pairing.cuh): what the reference's verifiers get from
ark-ec's `PairingEngine` -- groth16/src/verifier.rs:8-44 (`E::pairing`, `E::miller_loop` + `E::final_exponentiation`) and
marlin/src/pc/kzg10.rs:158-173 (`E::pairing` twice) -- for many checks in one call.

Points are (xy uint64[words], is_identity) in the layout of include/zkb.h; a GT element is uint64[12 * limbs(Fq)]
(Montgomery, ark-ff's Fq12 tower order).  GT values are a fixed power of ark-ec's (a different reduced pairing on the same
groups), so only equalities between them are meaningful -- exactly what the verifiers test."""
import numpy as np

from . import _lib
from .backend import FQ_LIMBS, point_words

FQ_MODULUS = {
    _lib.BLS12_381: 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
}


def neg_point(curve, group, pt):
    """-P: y -> q - y on every Fq component of y (the Montgomery form of -y is q - y_mont)"""
    xy, inf = pt
    if inf:
        return pt
    L, q = FQ_LIMBS[curve], FQ_MODULUS[curve]
    xy = np.array(xy, dtype=np.uint64).reshape(-1)
    half = point_words(curve, group) // 2
    for k in range(half // L):
        s = slice(half + k * L, half + (k + 1) * L)
        y = int.from_bytes(xy[s].tobytes(), "little")
        xy[s] = np.frombuffer(((q - y) % q).to_bytes(8 * L, "little"), dtype=np.uint64)
    return xy, False


def gt_one(curve):
    """1 in GT: (R mod q, 0, ..., 0)"""
    L, q = FQ_LIMBS[curve], FQ_MODULUS[curve]
    out = np.zeros(12 * L, dtype=np.uint64)
    out[:L] = np.frombuffer(((1 << (64 * L)) % q).to_bytes(8 * L, "little"), dtype=np.uint64)
    return out


def multi_pairing(ctx, curve, groups):
    """groups: equally long lists of (P in G1, Q in G2) -> one GT element per group, prod e(P, Q)
    (E::final_exponentiation(E::miller_loop(pairs)), verifier.rs:31-41), all groups in one device call"""
    if not groups:
        return []
    size = len(groups[0])
    if size == 0 or any(len(g) != size for g in groups):
        raise ValueError("multi_pairing: every group needs the same, non-zero number of pairs")
    flat = [pq for g in groups for pq in g]
    g1 = (np.stack([np.asarray(P[0], dtype=np.uint64).reshape(-1) for P, _ in flat]),
          np.array([1 if P[1] else 0 for P, _ in flat], dtype=np.uint8))
    g2 = (np.stack([np.asarray(Q[0], dtype=np.uint64).reshape(-1) for _, Q in flat]),
          np.array([1 if Q[1] else 0 for _, Q in flat], dtype=np.uint8))
    out = ctx.multi_pairing(curve, g1, g2, size)
    return [out[i] for i in range(len(groups))]


def pairing(ctx, curve, P, Q):
    """E::pairing(P, Q)"""
    return multi_pairing(ctx, curve, [[(P, Q)]])[0]
