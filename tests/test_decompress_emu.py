"""zkb_points_decompress's arithmetic on the CPU: the device code of csrc/serialize.cuh (Fq / Fq2 square roots, sign
selection, flag handling, subgroup check) compiled for the host with the PTX emulation of tests/host_emu, against the
oracle's Python serializer (oracle/pyref/serialize.py) on all four groups."""
import ctypes
import random

import numpy as np
import pytest

from oracle.pyref import serialize as S
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FQ
from tests import emu


@pytest.fixture(scope="module")
def lib():
    l = emu.build()
    l.emu_decompress.restype = ctypes.c_int
    return l


def run(lib, cid, group, data, check=0):
    L = FQ[cid].limbs * 2                                   # u32 limbs per Fq
    n_coords = 4 if group == 2 else 2
    buf = np.frombuffer(bytes(data), dtype=np.uint8).copy()
    out = np.zeros(n_coords * L, dtype=np.uint32)
    inf = np.zeros(1, dtype=np.uint8)
    st = lib.emu_decompress(cid, group, emu.ptr(buf), check, emu.ptr(out), emu.ptr(inf))
    fq = FQ[cid]
    cs = [emu.from_u32(out[k * L:(k + 1) * L]) * fq.Rinv % fq.p for k in range(n_coords)]
    pt = ((cs[0], cs[1]), (cs[2], cs[3])) if group == 2 else (cs[0], cs[1])
    return st, bool(inf[0]), pt


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)])
def test_decompress_round_trip(lib, cid, group):
    c = CURVES[(cid, group)]
    rng = random.Random(10 * cid + group)
    pts = [c.mul_affine(c.gen, rng.randrange(1, c.r)) for _ in range(6)] + [c.gen, c.neg_affine(c.gen)]
    for P in pts:
        data = S.compress(cid, group, P)
        assert S.decompress(cid, group, data) == P           # the oracle's own round trip
        st, inf, got = run(lib, cid, group, data, check=1)
        assert (st, inf) == (0, False) and got == P
        # the other root: flip the sign flag
        flipped = bytearray(data)
        flipped[-1] ^= 0x80
        st, inf, got = run(lib, cid, group, flipped)
        assert (st, inf) == (0, False) and got == c.neg_affine(P)
    st, inf, _ = run(lib, cid, group, S.compress(cid, group, None))
    assert (st, inf) == (0, True)


@pytest.mark.parametrize("cid,group", [(BN254, 1), (BLS12_381, 1), (BLS12_381, 2)])
def test_decompress_rejections(lib, cid, group):
    c = CURVES[(cid, group)]
    p = FQ[cid].p
    nb = 8 * FQ[cid].limbs
    rng = random.Random(99)
    # x with no point on the curve
    tried = 0
    while True:
        x = tuple(rng.randrange(p) for _ in range(2)) if group == 2 else rng.randrange(p)
        data = b"".join(int(v).to_bytes(nb, "little") for v in (x if group == 2 else (x,)))
        try:
            S.decompress(cid, group, data)
        except ValueError:
            break
        tried += 1
        assert tried < 50
    assert run(lib, cid, group, data)[0] == 2                # kDecompNotOnCurve
    # non-canonical x (>= p) where the byte width leaves room for it
    if p.bit_length() % 8 not in (0, 7) or cid == BLS12_381:
        big = (p + 5).to_bytes(nb, "little")
        if big[-1] & 0xC0 == 0:
            data = (big * 2) if group == 2 else big
            assert run(lib, cid, group, data)[0] == 1        # kDecompNotCanonical
    # a curve point outside the prime-order subgroup (cofactor != 1): only BLS12-381 G1 / G2 and BN254 G2 have them
    if (cid, group) != (BN254, 1):
        while True:
            x = tuple(rng.randrange(p) for _ in range(2)) if group == 2 else rng.randrange(p)
            data = b"".join(int(v).to_bytes(nb, "little") for v in (x if group == 2 else (x,)))
            try:
                P = S.decompress(cid, group, data)
            except ValueError:
                continue
            if c.to_affine(c.mul(c.from_affine(P), c.r)) is not None:
                break
        assert run(lib, cid, group, data, check=0)[0] == 0   # on the curve: accepted unchecked (deserialize_unchecked)
        assert run(lib, cid, group, data, check=1)[0] == 3   # kDecompNotInSubgroup
