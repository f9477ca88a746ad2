"""Field constants and Fq2 arithmetic for BN254 and BLS12-381.

Follows the arkworks-0.2 conventions the reference relies on:
Montgomery form with R = 2^(64*limbs), little-endian u64 limbs; Fr multiplicative
generator 7 (BLS12-381) / 5 (BN254); 2-adic root of unity = g^((r-1)/2^s);
domain root for size 2^k = root^(2^(s-k)).  Field elements here are plain
Python ints in [0, p) (canonical form); helpers convert to/from Montgomery limbs.
"""

BLS12_381 = 1
BN254 = 0


class FieldParams:
    def __init__(self, name, modulus, limbs64, generator=None):
        self.name = name
        self.p = modulus
        self.limbs = limbs64
        self.bits = modulus.bit_length()
        self.R = (1 << (64 * limbs64)) % modulus
        self.R2 = self.R * self.R % modulus
        self.Rinv = pow(self.R, -1, modulus)
        self.inv64 = (-pow(modulus, -1, 1 << 64)) % (1 << 64)
        self.inv32 = (-pow(modulus, -1, 1 << 32)) % (1 << 32)
        self.generator = generator
        if generator is not None:
            s = 0
            t = modulus - 1
            while t % 2 == 0:
                t //= 2
                s += 1
            self.two_adicity = s
            self.two_adic_root = pow(generator, (modulus - 1) >> s, modulus)

    # -- Montgomery helpers -------------------------------------------------
    def to_mont(self, x):
        return x * self.R % self.p

    def from_mont(self, x):
        return x * self.Rinv % self.p

    def root_of_unity(self, log_n):
        """ark-ff get_root_of_unity: square the 2-adic root (s - log_n) times."""
        assert log_n <= self.two_adicity
        w = self.two_adic_root
        for _ in range(log_n, self.two_adicity):
            w = w * w % self.p
        return w


BLS_FR = FieldParams(
    "bls12_381_fr",
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001, 4, generator=7)
BLS_FQ = FieldParams(
    "bls12_381_fq",
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB, 6)
BN_FR = FieldParams(
    "bn254_fr",
    0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001, 4, generator=5)
BN_FQ = FieldParams(
    "bn254_fq",
    0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47, 4)

FR = {BLS12_381: BLS_FR, BN254: BN_FR}
FQ = {BLS12_381: BLS_FQ, BN254: BN_FQ}


# ---------------------------------------------------------------------------
# Generic field "ops" objects so curve code is written once for Fq and Fq2.
# ---------------------------------------------------------------------------
class FpOps:
    """Prime field ops on ints."""

    def __init__(self, p):
        self.p = p
        self.zero = 0
        self.one = 1

    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def neg(self, a):
        return (-a) % self.p

    def mul(self, a, b):
        return a * b % self.p

    def sqr(self, a):
        return a * a % self.p

    def inv(self, a):
        return pow(a, -1, self.p)

    def is_zero(self, a):
        return a == 0

    def small(self, k):
        return k % self.p


class Fp2Ops:
    """Fq[u]/(u^2+1) on tuples (c0, c1) (both BN254 and BLS12-381 use u^2 = -1)."""

    def __init__(self, p):
        self.p = p
        self.zero = (0, 0)
        self.one = (1, 0)

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def neg(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a):
        return self.mul(a, a)

    def inv(self, a):
        p = self.p
        n = pow((a[0] * a[0] + a[1] * a[1]) % p, -1, p)
        return (a[0] * n % p, (-a[1] * n) % p)

    def is_zero(self, a):
        return a[0] == 0 and a[1] == 0

    def small(self, k):
        return (k % self.p, 0)


# ---------------------------------------------------------------------------
# limb (de)serialisation helpers (little-endian u64 words, as ark BigInteger)
# ---------------------------------------------------------------------------
def int_to_limbs(x, n):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def limbs_to_int(ws):
    r = 0
    for i, w in enumerate(ws):
        r |= int(w) << (64 * i)
    return r


# ---------------------------------------------------------------------------
# SplitMix64 counter-based stream used for all synthetic data
# ---------------------------------------------------------------------------
M64 = (1 << 64) - 1


def splitmix64(seed, k):
    """k-th output (k >= 0) of SplitMix64 seeded with `seed`."""
    z = (seed + (k + 1) * 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def stream_field(seed, idx, p):
    """Field element #idx of stream `seed`: 4 outputs -> 256-bit LE integer mod p."""
    v = 0
    for j in range(4):
        v |= splitmix64(seed, 4 * idx + j) << (64 * j)
    return v % p
