"""Marlin's Fiat-Shamir generator (marlin/src/fs_rng.rs) on the CPU: the oracle's pure-Python restatement and the
product's host layer (C helpers in libzkb.so for Keccak-f / ChaCha20, no GPU needed) against published vectors
and against each other.

  Keccak-f[1600]   hashlib's SHA3-256 through the oracle's sponge
  ChaCha20         RFC 7539 section 2.3.2 block, and the `cryptography` package's keystream
  STROBE-128       the conformance vector of merlin's strobe.rs test-suite
  Merlin           the "test protocol" transcript vector of merlin's transcript.rs test-suite
"""
import ctypes
import hashlib
import random

import numpy as np
import pytest

from ckb_zkp_b200 import _lib, fs_rng
from oracle.pyref import transcript as T
from oracle.pyref.fields import BLS12_381, BN254, FR


def test_keccak_f_against_sha3():
    for m in (b"", b"abc", b"x" * 135, b"y" * 136, b"z" * 137, bytes(range(256)) * 3):
        assert T.sha3_256(m) == hashlib.sha3_256(m).digest()


def test_c_keccak_f_equals_the_oracle_permutation():
    lib = _lib.load()
    rng = random.Random(1)
    for _ in range(20):
        st = bytearray(rng.randrange(256) for _ in range(200))
        want = bytearray(st)
        T.keccak_f1600(want)
        arr = np.frombuffer(bytes(st), dtype=np.uint64).copy()
        lib.zkb_host_keccak_f1600(arr.ctypes.data_as(ctypes.c_void_p))
        assert arr.tobytes() == bytes(want)


def test_chacha20_rfc7539_block_and_keystream():
    key = bytes(range(32))
    blk = T.chacha20_block(key, 1, nonce_words=(0x09000000, 0x4A000000, 0), counter_words=1)
    assert blk[:4] == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3] and blk[-1] == 0x4E3C50A2
    oracle, product = T.ChaChaRng(key), fs_rng.ChaChaRng(key)
    words = [oracle.next_u32() for _ in range(200)]
    assert words == [product.next_u32() for _ in range(200)]
    o2, p2 = T.ChaChaRng(key), fs_rng.ChaChaRng(key)
    assert [o2.next_u64() for _ in range(70)] == [p2.next_u64() for _ in range(70)]
    try:
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
    except ImportError:
        return
    ks = Cipher(algorithms.ChaCha20(key, bytes(16)), mode=None).encryptor().update(bytes(800))
    assert ks == b"".join(w.to_bytes(4, "little") for w in words)


def _strobe_conformance(cls):
    s = cls(b"Conformance Test Protocol")
    s.meta_ad(b"ms", False)
    s.meta_ad(b"g", True)
    s.ad(bytes([99]) * 1024, False)
    s.meta_ad(b"prf", False)
    return s.prf(32, False).hex()


def test_strobe_and_merlin_published_vectors():
    want = "b48e645ca17c667fd5206ba57a6a228d72d8e1903814d3f17f622996d7cfefb0"
    assert _strobe_conformance(T.Strobe128) == want
    assert _strobe_conformance(fs_rng._Strobe128) == want
    for cls in (T.Transcript, fs_rng.Transcript):
        t = cls(b"test protocol")
        t.append_message(b"some label", b"some data")
        assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_fiat_shamir_rng_product_equals_oracle(cid):
    """from_seed / absorb / Fr::rand / u128::rand: same stream from both implementations, over materials that cross the
    STROBE rate (166 bytes) in every way"""
    p = FR[cid].p
    rng = random.Random(cid)
    mat = bytes(rng.randrange(256) for _ in range(777))
    o, f = T.FiatShamirRng(mat), fs_rng.FiatShamirRng(mat, cid)
    assert o.seed == f.seed
    for n in (0, 1, 165, 166, 167, 331, 332, 2000):
        more = bytes(rng.randrange(256) for _ in range(n))
        o.absorb(more)
        f.absorb(more)
        assert o.seed == f.seed
        for _ in range(5):
            a, b = o.rand_fr(p), f.rand_fr()
            assert a == b and 0 <= a < p
        assert o.rand_u128() == f.rand_u128()


def test_rand_fr_takes_the_draw_as_the_montgomery_residue():
    """ark-ff 0.2 `Fp256::rand`: four u64 from the generator, top REPR_SHAVE_BITS cleared, value < p accepted and used as
    the in-memory (Montgomery) limbs -- so the field element is draw * R^-1"""
    for cid in (BN254, BLS12_381):
        p = FR[cid].p
        o1, o2 = T.FiatShamirRng(b"seed"), T.FiatShamirRng(b"seed")
        raw = o1.rand_fr_mont(p)
        assert raw < p and o2.rand_fr(p) == raw * pow(1 << 256, -1, p) % p


def test_to_bytes_layouts():
    # identity commitment = ark's GroupAffine::zero(): x = 0, y = 1, infinity = true
    empty = (np.zeros(8, dtype=np.uint64), True)
    b = fs_rng.affine_to_bytes(BN254, empty)
    assert len(b) == 65 and b[:32] == bytes(32) and b[32:64] == (1).to_bytes(32, "little") and b[64] == 1
    # a finite BN254 G1 point (1, 2) in Montgomery limbs -> canonical little-endian coordinates
    from tests import helpers as H
    xy, inf = H.points_array(BN254, 1, [(1, 2)])
    b = fs_rng.affine_to_bytes(BN254, (xy[0], False))
    assert b == (1).to_bytes(32, "little") + (2).to_bytes(32, "little") + b"\x00"
    c = fs_rng.commitment_to_bytes(BN254, ((xy[0], False), None))
    assert len(c) == 65 + 1 + 65 and c[65] == 0 and c[66:] == fs_rng.affine_to_bytes(BN254, empty)
    c = fs_rng.commitment_to_bytes(BN254, ((xy[0], False), (xy[0], False)))
    assert c[65] == 1 and c[66:] == c[:65]
    assert fs_rng.index_info_to_bytes(3, 4, 5) == (3).to_bytes(8, "little") + (4).to_bytes(8, "little") + (5).to_bytes(8, "little")
    assert fs_rng.fr_to_bytes(7) == (7).to_bytes(32, "little")
    # G2 (BLS12-381): x.c0, x.c1, y.c0, y.c1 of 48 bytes each + flag
    xy2, _ = H.points_array(BLS12_381, 2, [((1, 2), (3, 4))])
    b2 = fs_rng.affine_to_bytes(BLS12_381, (xy2[0], False))
    assert b2 == b"".join(v.to_bytes(48, "little") for v in (1, 2, 3, 4)) + b"\x00"
