"""Marlin's Fiat-Shamir generator (TEST INFRASTRUCTURE: part of the CPU oracle, never imported by the product).

Restates marlin/src/fs_rng.rs:9-69 -- `FiatShamirRng { r: ChaChaRng, seed: [u8; 32] }` with

    from_seed(s):  seed = H(bytes(s));               r = ChaChaRng::from_seed(seed)
    absorb(s):     seed = H(bytes(s) || old seed);   r = ChaChaRng::from_seed(seed)
    H(m) = Transcript::new(b"MARLINSEED"); append_message(b"Seed", m); challenge_bytes(b"x", 32 bytes)

on top of the three third-party pieces the reference pulls in (marlin/Cargo.toml:19-20; none of them vendored):

  * merlin 2.0 `Transcript` over STROBE-128/1600 (`Strobe128` of merlin's strobe.rs) over Keccak-f[1600];
  * rand_chacha 0.2 `ChaChaRng` = ChaCha20, 64-bit block counter from 0, 64-bit stream id 0, keystream consumed as
    little-endian u32 words (`next_u64` = two consecutive words, low first);
  * ark-ff 0.2 `UniformRand`: `Fr::rand` draws 4 u64, clears the top REPR_SHAVE_BITS of the last one, accepts when the
    value is below the modulus and takes it AS THE MONTGOMERY RESIDUE; `u128::rand` = low u64 then high u64.

Pinning: Keccak-f against hashlib's SHA3-256, ChaCha20 against RFC 7539 section 2.3.2 and the `cryptography` package,
STROBE against the conformance vector of merlin's own test-suite (tests/test_transcript.py).  The arkworks pieces are
recalled from the published 0.2 crates and cannot be checked in this container.
"""
M64 = (1 << 64) - 1

_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & M64 if n else x


def keccak_f1600(state):
    """in-place permutation of a 200-byte bytearray (lane (x, y) = little-endian u64 at 8 * (x + 5 y))"""
    a = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = (a[x][y] & M64).to_bytes(8, "little")


def sha3_256(msg):
    """SHA3-256 over the permutation above (rate 136, domain byte 0x06): the known-answer hook for keccak_f1600"""
    rate = 136
    st = bytearray(200)
    m = bytearray(msg) + b"\x06" + b"\x00" * ((-len(msg) - 2) % rate) + b"\x80" if (len(msg) + 1) % rate else bytearray(msg) + b"\x86"
    for off in range(0, len(m), rate):
        for i in range(rate):
            st[i] ^= m[off + i]
        keccak_f1600(st)
    return bytes(st[:32])


class Strobe128:
    """merlin's strobe.rs (a subset of STROBE v1.0.2 with R = 166)"""
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label, permutation=keccak_f1600):
        self.f = permutation
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        self.f(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.R + 1] ^= 0x80
        self.f(self.state)
        self.pos, self.pos_begin = 0, 0

    def _absorb(self, data):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _overwrite(self, data):
        for byte in data:
            self.state[self.pos] = byte
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray(n)
        for i in range(n):
            out[i] = self.state[self.pos]
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags, "continued operation with different flags"
            return
        assert not flags & self.FLAG_T, "transport operations are not part of merlin's subset"
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

    def key(self, data, more):
        self._begin_op(self.FLAG_A | self.FLAG_C, more)
        self._overwrite(data)


class Transcript:
    """merlin 2.0 `Transcript` (transcript.rs): the two operations fs_rng.rs uses"""

    def __init__(self, label, permutation=keccak_f1600):
        self.strobe = Strobe128(b"Merlin v1.0", permutation)
        self.append_message(b"dom-sep", label)

    def append_message(self, label, message):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)


def _quarter(s, a, b, c, d):
    M = 0xFFFFFFFF
    rol = lambda v, n: ((v << n) | (v >> (32 - n))) & M
    s[a] = (s[a] + s[b]) & M; s[d] = rol(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & M; s[b] = rol(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & M; s[d] = rol(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & M; s[b] = rol(s[b] ^ s[c], 7)


def chacha20_block(key, counter, nonce_words=(0, 0), counter_words=2):
    """one 64-byte ChaCha20 block as 16 u32 words.  counter_words = 2: the original layout rand_chacha uses (words 12-13
    = 64-bit block counter, 14-15 = stream id); counter_words = 1: RFC 7539 (word 12 = counter, 13-15 = nonce)."""
    k = [int.from_bytes(key[4 * i:4 * i + 4], "little") for i in range(8)]
    if counter_words == 2:
        tail = [counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF] + list(nonce_words)
    else:
        tail = [counter & 0xFFFFFFFF] + list(nonce_words)
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + k + tail
    s = list(init)
    for _ in range(10):
        _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13); _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
        _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12); _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
    return [(x + y) & 0xFFFFFFFF for x, y in zip(s, init)]


class ChaChaRng:
    """rand_chacha 0.2 ChaChaRng (ChaCha20) seen through rand_core's BlockRng: a stream of u32 words"""

    def __init__(self, seed):
        assert len(seed) == 32
        self.key, self.counter, self.buf = bytes(seed), 0, []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha20_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


def hash_seed(material, permutation=keccak_f1600):
    t = Transcript(b"MARLINSEED", permutation)
    t.append_message(b"Seed", material)
    return t.challenge_bytes(b"x", 32)


class FiatShamirRng:
    """marlin/src/fs_rng.rs:9-69 over already serialised bytes (the ToBytes layout lives with the caller)"""

    def __init__(self, seed_material, permutation=keccak_f1600):
        self.f = permutation
        self.seed = hash_seed(seed_material, permutation)
        self.r = ChaChaRng(self.seed)

    def absorb(self, material):
        self.seed = hash_seed(bytes(material) + self.seed, self.f)
        self.r = ChaChaRng(self.seed)

    def next_u64(self):
        return self.r.next_u64()

    # ark-ff 0.2 UniformRand ----------------------------------------------------------------
    def rand_u128(self):
        lo = self.next_u64()
        return lo | (self.next_u64() << 64)

    def rand_fr_mont(self, modulus):
        """Fr::rand: the accepted 256-bit value IS the Montgomery residue (limbs little-endian)"""
        shave = 256 - modulus.bit_length()
        while True:
            limbs = [self.next_u64() for _ in range(4)]
            limbs[3] &= M64 >> shave
            v = sum(l << (64 * i) for i, l in enumerate(limbs))
            if v < modulus:
                return v

    def rand_fr(self, modulus):
        """the field element Fr::rand returns, as a canonical integer"""
        return self.rand_fr_mont(modulus) * pow(1 << 256, -1, modulus) % modulus
