"""ark-serialize 0.2 compressed short-Weierstrass points (TEST INFRASTRUCTURE: oracle side of zkb_points_decompress).

Layout recalled from ark-ec / ark-serialize 0.2 (`GroupAffine::serialize`, `SWFlags`; not vendored with the reference,
SURVEY.md 8c): x as canonical little-endian bytes (Fq2: c0 then c1), flags in the top bits of the last byte --
bit 7 = y is the larger of (y, -y) as canonical integers (Fq2: compare c1, then c0), bit 6 = point at infinity (x = 0).
The reference reaches this through `Parameters::<E>::deserialize` (cli/src/zkp_prove.rs:117-124, groth16/src/lib.rs:81)."""
from .curves import CURVES
from .fields import FQ


def _key(group, y):
    return (y[1], y[0]) if group == 2 else y


def compress(cid, group, P):
    """affine tuple / None -> bytes"""
    nb = 8 * FQ[cid].limbs
    c = CURVES[(cid, group)]
    if P is None:
        out = bytearray(nb * (2 if group == 2 else 1))
        out[-1] |= 0x40
        return bytes(out)
    x, y = P
    neg = c.F.neg(y)
    greatest = _key(group, y) > _key(group, neg)
    out = bytearray(b"".join(int(v).to_bytes(nb, "little") for v in (x if group == 2 else (x,))))
    if greatest:
        out[-1] |= 0x80
    return bytes(out)


def sqrt_fq(cid, a):
    p = FQ[cid].p
    r = pow(a, (p + 1) // 4, p)
    return r if r * r % p == a % p else None


def sqrt_fq2(cid, a):
    """complex method for Fq[u]/(u^2 + 1), p = 3 mod 4"""
    p = FQ[cid].p
    a0, a1 = a
    if a1 == 0:
        s = sqrt_fq(cid, a0)
        if s is not None:
            return (s, 0)
        s = sqrt_fq(cid, -a0 % p)
        return None if s is None else (0, s)
    s = sqrt_fq(cid, (a0 * a0 + a1 * a1) % p)
    if s is None:
        return None
    half = pow(2, -1, p)
    for t in ((a0 + s) * half % p, (a0 - s) * half % p):
        x0 = sqrt_fq(cid, t)
        if x0 is not None and x0 != 0:
            return (x0, a1 * pow(2 * x0, -1, p) % p)
    return None


def decompress(cid, group, data):
    """bytes -> (affine tuple / None); raises ValueError like SerializationError::InvalidData"""
    c = CURVES[(cid, group)]
    F = c.F
    p = FQ[cid].p
    nb = 8 * FQ[cid].limbs
    flags = data[-1]
    if flags & 0x40:
        return None
    raw = bytearray(data)
    raw[-1] &= 0x3F
    coords = [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(len(raw) // nb)]
    if any(v >= p for v in coords):
        raise ValueError("non-canonical coordinate")
    x = tuple(coords) if group == 2 else coords[0]
    rhs = F.add(F.mul(F.sqr(x), x), c.b)
    y = sqrt_fq2(cid, rhs) if group == 2 else sqrt_fq(cid, rhs)
    if y is None:
        raise ValueError("x is not on the curve")
    neg = F.neg(y)
    if (_key(group, neg) > _key(group, y)) == bool(flags & 0x80):
        y = neg
    return (x, y)
