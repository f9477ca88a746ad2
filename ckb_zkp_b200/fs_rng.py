"""Marlin's Fiat-Shamir generator and the byte layouts it hashes -- the host-side half of
`zkp_marlin::create_random_proof` (marlin/src/lib.rs:97-181) that a Rust host gets from merlin, rand_chacha and ark-ff.

    FiatShamirRng          marlin/src/fs_rng.rs:9-69   seed = H(material [|| old seed]); r = ChaChaRng::from_seed(seed)
                                                        H = merlin Transcript("MARLINSEED").append_message("Seed", .)
                                                            .challenge_bytes("x", 32)
    rand_fr / rand_u128    ark-ff 0.2 UniformRand as used by ahp/verifier.rs:40-127 and lib.rs:158
    ToBytes layouts        marlin/src/data_structures.rs:22-33, pc/data_structures.rs:111-154, ahp/indexer.rs:19-26

The Keccak-f[1600] permutation and the ChaCha20 block function run in the C library's host helpers
(zkb_host_keccak_f1600 / zkb_host_chacha20_blocks: no GPU involved); STROBE-128, Merlin's framing and the sampling
rules are restated here.  The transcript is pluggable: `create_random_proof(..., fs_rng=...)` accepts any object with
absorb(bytes), rand_fr(), rand_u128() -- e.g. one fed by the Rust host.

What is pinned and what is recalled: STROBE / Merlin / ChaCha20 are checked against their published vectors
(tests/test_transcript.py).  The arkworks `ToBytes` layouts (field element = canonical integer, little-endian u64 limbs;
affine point = x, y, infinity byte; bool = one byte; Vec<T> = items back to back, no length) and `Fr::rand` (four u64,
top bits shaved, accepted value taken as the Montgomery residue) are recalled from the 0.2 crates, which are not
vendored with the reference: byte-level parity of the challenge stream with the Rust prover is UNPINNED.
"""
import ctypes

import numpy as np

from . import _lib

M64 = (1 << 64) - 1
FR_MODULUS = {
    _lib.BLS12_381: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
}
FQ_MODULUS = {
    _lib.BLS12_381: 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
}
FQ_LIMBS = {_lib.BN254: 4, _lib.BLS12_381: 6}


class _Strobe128:
    """merlin's strobe.rs: STROBE v1.0.2 restricted to AD / meta-AD / PRF, rate 166"""
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_M, FLAG_K = 1, 2, 4, 16, 32

    def __init__(self, protocol_label):
        self.lib = _lib.load()
        self.state = np.zeros(25, dtype=np.uint64)
        self.bytes = self.state.view(np.uint8)
        self.bytes[0:6] = [1, self.R + 2, 1, 0, 1, 96]
        self.bytes[6:18] = np.frombuffer(b"STROBEv1.0.2", dtype=np.uint8)
        self._f()
        self.pos, self.pos_begin, self.cur_flags = 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _f(self):
        self.lib.zkb_host_keccak_f1600(self.state.ctypes.data_as(ctypes.c_void_p))

    def _run_f(self):
        self.bytes[self.pos] ^= self.pos_begin
        self.bytes[self.pos + 1] ^= 0x04
        self.bytes[self.R + 1] ^= 0x80
        self._f()
        self.pos, self.pos_begin = 0, 0

    def _absorb(self, data):
        data = np.frombuffer(bytes(data), dtype=np.uint8)
        off = 0
        while off < len(data):
            take = min(self.R - self.pos, len(data) - off)
            self.bytes[self.pos:self.pos + take] ^= data[off:off + take]
            self.pos += take
            off += take
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        while len(out) < n:
            take = min(self.R - self.pos, n - len(out))
            out += self.bytes[self.pos:self.pos + take].tobytes()
            self.bytes[self.pos:self.pos + take] = 0
            self.pos += take
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            if self.cur_flags != flags:
                raise ValueError("continued STROBE operation with different flags")
            return
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (self.FLAG_C | self.FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more):
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    """merlin 2.0 Transcript: new / append_message / challenge_bytes"""

    def __init__(self, label):
        self.strobe = _Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, message):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)


class ChaChaRng:
    """rand_chacha 0.2 `ChaChaRng::from_seed` as a stream of little-endian u32 words (next_u64 = two words, low first)"""
    BLOCKS = 4                                   # the crate buffers four blocks too; only the word order is observable

    def __init__(self, seed):
        if len(seed) != 32:
            raise ValueError("ChaChaRng seed is 32 bytes")
        self.lib = _lib.load()
        self.key = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
        self.counter, self.buf, self.idx = 0, np.zeros(16 * self.BLOCKS, dtype=np.uint32), 16 * self.BLOCKS

    def next_u32(self):
        if self.idx == len(self.buf):
            self.lib.zkb_host_chacha20_blocks(self.key.ctypes.data_as(ctypes.c_void_p), self.counter,
                                              self.buf.ctypes.data_as(ctypes.c_void_p), self.BLOCKS)
            self.counter += self.BLOCKS
            self.idx = 0
        v = int(self.buf[self.idx])
        self.idx += 1
        return v

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


def _hash_seed(material):
    t = Transcript(b"MARLINSEED")
    t.append_message(b"Seed", material)
    return t.challenge_bytes(b"x", 32)


class FiatShamirRng:
    """marlin/src/fs_rng.rs:9-69 for one scalar field; material is already-serialised bytes (see the layouts below)"""

    def __init__(self, seed_material, curve):
        self.p = FR_MODULUS[curve]
        self.shave = 256 - self.p.bit_length()           # REPR_SHAVE_BITS
        self.rinv = pow(1 << 256, -1, self.p)
        self.seed = _hash_seed(bytes(seed_material))
        self.r = ChaChaRng(self.seed)

    def absorb(self, material):
        self.seed = _hash_seed(bytes(material) + self.seed)    # bytes.extend_from_slice(&self.seed)  (fs_rng.rs:58)
        self.r = ChaChaRng(self.seed)

    def rand_u128(self):
        lo = self.r.next_u64()
        return lo | (self.r.next_u64() << 64)

    def rand_fr(self):
        """`Fr::rand(rng)` as a canonical integer"""
        while True:
            limbs = [self.r.next_u64() for _ in range(4)]
            limbs[3] &= M64 >> self.shave
            v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
            if v < self.p:
                return v * self.rinv % self.p


# ------------------------------------------------------------------------------------------------
# ark-ff / ark-ec 0.2 `ToBytes` layouts of what create_random_proof hashes
# ------------------------------------------------------------------------------------------------
def fr_to_bytes(x):
    """Fp256::write = into_repr().write: canonical integer, 4 little-endian u64"""
    return int(x).to_bytes(32, "little")


def fr_mont_array_to_bytes(ctx, curve, a_mont):
    """Vec<Fr>::write for a Montgomery limb array: items back to back (converted on the GPU)"""
    a = np.ascontiguousarray(a_mont, dtype=np.uint64).reshape(-1, 4)
    return ctx.fr_convert(curve, a, to_mont=False).tobytes() if len(a) else b""


def _fq_canonical(curve, limbs_mont):
    L = FQ_LIMBS[curve]
    q = FQ_MODULUS[curve]
    v = int.from_bytes(np.ascontiguousarray(limbs_mont, dtype=np.uint64).tobytes(), "little")
    return (v * pow(1 << (64 * L), -1, q) % q).to_bytes(8 * L, "little")


def affine_to_bytes(curve, point):
    """GroupAffine::write: x, y (Fq2: c0 then c1) as canonical integers, then the infinity flag as one byte.
    point = (xy Montgomery limbs in the ABI layout, is_identity); the identity is written as ark's zero() = (0, 1, true)"""
    xy, inf = point
    L = FQ_LIMBS[curve]
    xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1)
    n_coords = len(xy) // L
    if inf:
        one = (1).to_bytes(8 * L, "little")
        zero = bytes(8 * L)
        coords = [zero] * (n_coords // 2) + [one] + [zero] * (n_coords // 2 - 1)
    else:
        coords = [_fq_canonical(curve, xy[k * L:(k + 1) * L]) for k in range(n_coords)]
    return b"".join(coords) + (b"\x01" if inf else b"\x00")


def commitment_to_bytes(curve, commitment):
    """pc::Commitment::write (pc/data_structures.rs:143-154): comm, shifted_exists, shifted comm or Comm::empty()"""
    comm, shifted = commitment
    n_words = len(np.asarray(comm[0]).reshape(-1))
    empty = (np.zeros(n_words, dtype=np.uint64), True)
    return (affine_to_bytes(curve, comm) + (b"\x01" if shifted is not None else b"\x00")
            + affine_to_bytes(curve, shifted if shifted is not None else empty))


def commitments_to_bytes(curve, commitments):
    """to_bytes![Vec<LabeledCommitment>]: LabeledCommitment::write is the commitment's (pc/data_structures.rs:293-298)"""
    return b"".join(commitment_to_bytes(curve, c) for c in commitments)


def index_info_to_bytes(num_variables, num_constraints, num_non_zeros):
    """ahp/indexer.rs:19-26"""
    return b"".join(int(v).to_bytes(8, "little") for v in (num_variables, num_constraints, num_non_zeros))
