"""Synthetic workloads for bench.py and the full-size tests.

* `mimc_instance`: the MiMC-chain R1CS (shaped after marlin/examples/mimc.rs:26-118 and
  gadgets/src/hashes/mimc.rs:119-157) as CSR matrices + full assignment, built directly as arrays
  (the per-constraint `enforce` path of r1cs.ProvingAssignment is too slow at 2^20 in Python).
  For n constraints: num_inputs = 2 (ONE, image), num_aux = n + 1, nnz(A) = 1.5n, nnz(B) = 2n,
  nnz(C) = 1.5n.
* `synthetic_key`: a proving key whose points are k_i * G for KNOWN pseudo-random exponents k_i.
  Proving time does not depend on the points' values, and the exponents make the proof checkable
  at full size "in the exponent" (every proof element is a known multiple of the generator).

Self-contained (the product never imports oracle/).
"""
import numpy as np

from . import _lib
from .r1cs import ints_to_limbs

M64 = (1 << 64) - 1
MIMC_SEED = 0x5ECB17

FR_MODULUS = {
    _lib.BLS12_381: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
}
FQ_MODULUS = {
    _lib.BLS12_381: 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
}
# generators (standard; x, y[, for G2: x = c0 + c1 u]) as canonical integers
G1_GEN = {
    _lib.BLS12_381: (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
                     0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1),
    _lib.BN254: (1, 2),
}
G2_GEN = {
    _lib.BLS12_381: (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
                     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
                     0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
                     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
    _lib.BN254: (10857046999023057135944570762232829481370756359578518086990519993285655852781,
                 11559732032986387107991004021392285783925812861821192530917403151452391805634,
                 8495653923123431417604973247489272438418190587263600148770280649306958101930,
                 4082367875863433681332203403145435568316851327593401208105741076214120093531),
}


def splitmix64_stream(seed, start, count):
    """outputs start .. start+count-1 of SplitMix64(seed) as uint64 (vectorised)."""
    k = np.arange(start + 1, start + count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream_field_ints(seed, start, count, p):
    """field elements start .. start+count-1 of stream `seed`: 4 outputs -> 256-bit LE integer mod p"""
    words = splitmix64_stream(seed, 4 * start, 4 * count).reshape(count, 4)
    raw = words.tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") % p for i in range(count)]


def generator_mont(curve, group):
    """generator as one affine point in the ABI layout (Montgomery u64 limbs)."""
    q = FQ_MODULUS[curve]
    L = 4 if curve == _lib.BN254 else 6
    R = (1 << (64 * L)) % q
    coords = G1_GEN[curve] if group == _lib.G1 else G2_GEN[curve]
    return ints_to_limbs([c * R % q for c in coords], L).reshape(-1)


class MimcInstance:
    """CSR matrices (coefficients canonical ints until `to_device_form`) and the full assignment."""

    def __init__(self, curve, n_constraints, seed=MIMC_SEED):
        assert n_constraints % 2 == 0 and n_constraints >= 2
        p = FR_MODULUS[curve]
        rounds = n_constraints // 2
        vals = stream_field_ints(seed, 0, rounds + 2, p)
        xl, xr, consts = vals[0], vals[1], vals[2:]
        self.curve, self.n_constraints, self.p = curve, n_constraints, p
        self.seed, self.consts = seed, consts          # the matrices below carry THESE round constants
        self.n_inputs, self.n_aux = 2, n_constraints + 1
        # variable numbering (global column index): 0 = ONE, 1 = image (last xl'), aux k -> 2 + k
        # aux 0 = xl0, aux 1 = xr0, round i: tmp_i = aux 2+2i, xl'_i = aux 3+2i (except the last: input 1)
        z = [0] * (2 + self.n_aux)
        z[0] = 1
        z[2], z[3] = xl, xr
        for i in range(rounds):
            c = consts[i]
            t = (xl + c) * (xl + c) % p
            new = ((xl + c) * t + xr) % p
            z[4 + 2 * i] = t
            if i == rounds - 1:
                z[1] = new
            else:
                z[5 + 2 * i] = new
            xr, xl = xl, new
        self.z = z
        i = np.arange(rounds, dtype=np.int64)
        col_xl = np.where(i == 0, 2, 3 + 2 * i)             # xl of round i: aux0 or xl'_{i-1} = aux 3+2(i-1) -> col 5+2(i-1)
        col_xl = np.where(i == 0, 2, 5 + 2 * (i - 1))
        col_xr = np.where(i == 0, 3, np.where(i == 1, 2, 5 + 2 * (i - 2)))
        col_tmp = 4 + 2 * i
        col_new = np.where(i == rounds - 1, 1, 5 + 2 * i)
        zero = np.zeros(rounds, dtype=np.int64)
        # row 2i  : (xl + c) * (xl + c) = tmp          A: [xl, c*ONE]  B: [xl, c*ONE]  C: [tmp]
        # row 2i+1: tmp * (xl + c) = new - xr          A: [tmp]        B: [xl, c*ONE]  C: [new, -xr]
        # coefficient codes: -1 -> round constant c_i, -2 -> p - 1, 1 -> one
        self.A = self._interleave([(col_xl, 1), (zero, -1)], [(col_tmp, 1)], consts)
        self.B = self._interleave([(col_xl, 1), (zero, -1)], [(col_xl, 1), (zero, -1)], consts)
        self.C = self._interleave([(col_tmp, 1)], [(col_new, 1), (col_xr, -2)], consts)

    def _interleave(self, even_terms, odd_terms, consts):
        rounds = len(consts)
        ne, no = len(even_terms), len(odd_terms)
        per = ne + no
        cols = np.zeros((rounds, per), dtype=np.uint32)
        codes = np.zeros((rounds, per), dtype=np.int64)
        for k, (c, code) in enumerate(even_terms + odd_terms):
            cols[:, k] = c
            codes[:, k] = code
        row_ptr = np.zeros(2 * rounds + 1, dtype=np.uint32)
        row_ptr[1::2] = np.arange(rounds, dtype=np.uint32) * per + ne
        row_ptr[2::2] = (np.arange(rounds, dtype=np.uint32) + 1) * per
        return row_ptr, cols.reshape(-1), codes.reshape(-1), per

    def coeff_ints(self, which):
        """canonical coefficient list of matrix `which` (small instances / tests)"""
        row_ptr, cols, codes, per = getattr(self, which)
        consts = self.consts
        out = []
        for idx, code in enumerate(codes):
            out.append(1 if code == 1 else self.p - 1 if code == -2 else consts[idx // per])
        return out

    def device_form(self, ctx):
        """(A, B, C CsrMatrix, z_mont) with Montgomery coefficients (converted on the GPU)."""
        from .backend import CsrMatrix
        p = self.p
        rounds = self.n_constraints // 2
        consts = self.consts
        table = ctx.fr_convert(self.curve, ints_to_limbs(consts + [1, p - 1]), to_mont=True)
        mats = []
        for which in "ABC":
            row_ptr, cols, codes, per = getattr(self, which)
            idx = np.where(codes == 1, rounds, np.where(codes == -2, rounds + 1, np.arange(len(codes)) // per))
            mats.append(CsrMatrix(row_ptr, cols, table[idx]))
        z_mont = ctx.fr_convert(self.curve, ints_to_limbs(self.z), to_mont=True)
        return mats[0], mats[1], mats[2], z_mont


def random_exponents(rng, n):
    """n pseudo-random scalars < 2^252 (below both Fr moduli) as uint64[n, 4]"""
    k = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 60) - 1)
    return k


def random_scalars(rng, n, curve):
    """n pseudo-random FULL-WIDTH scalars below the curve's Fr modulus as uint64[n, 4] (top limb drawn below the
    modulus' top limb: every window of the Pippenger decomposition is populated, including the last one)"""
    k = rng.integers(0, np.iinfo(np.uint64).max, size=(n, 4), dtype=np.uint64, endpoint=True)
    k[:, 3] = rng.integers(0, FR_MODULUS[curve] >> 192, size=n, dtype=np.uint64)
    return k


def limbs_to_ints(arr):
    raw = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    w = arr.shape[1] * 8
    return [int.from_bytes(raw[w * i:w * (i + 1)], "little") for i in range(arr.shape[0])]


class SyntheticKey:
    """Exponents of a synthetic Groth16 proving key for an instance with `n_vars` variables
    (inputs + aux), `n_inputs` inputs and domain size N; b is zero (identity point) for the
    variables listed in `b_zero_cols`, mirroring b_g1_query / b_g2_query of a real key where
    variables absent from B map to the identity (groth16/src/generator.rs:218-223)."""

    def __init__(self, n_vars, n_inputs, domain, b_zero_cols=None, seed=7):
        rng = np.random.default_rng(seed)
        self.n_vars, self.n_inputs, self.domain = n_vars, n_inputs, domain
        self.a = random_exponents(rng, n_vars)
        self.b = random_exponents(rng, n_vars)
        if b_zero_cols is not None:
            self.b[b_zero_cols] = 0
        self.h = random_exponents(rng, domain - 1)
        self.l = random_exponents(rng, n_vars - n_inputs)
        self.alpha, self.beta, self.delta = [random_exponents(rng, 1)[0] for _ in range(3)]

    def upload(self, ctx, curve, chunk=1 << 18, shard=None, keep_host=False):
        """points = exponent * generator, computed on the GPU (zkb_fixed_base_mul) -> groth16.Parameters.
        shard = (n_ranks, rank): a key sharded over the ranks (every rank computes the whole key, keeps its slice).
        keep_host: also keep the point arrays on the host as self.host_points (the CPU baseline proves with them)."""
        from .groth16 import Parameters

        def pts(group, k):
            gen = generator_mont(curve, group)
            xs, infs = [], []
            for i in range(0, len(k), chunk):
                xy, inf = ctx.fixed_base_mul(curve, group, gen, k[i:i + chunk])
                xs.append(xy)
                infs.append(inf)
            return np.concatenate(xs), np.concatenate(infs)

        singles1, _ = pts(_lib.G1, np.stack([self.alpha, self.beta, self.delta]))
        singles2, _ = pts(_lib.G2, np.stack([self.beta, self.delta]))
        q = {"a": pts(_lib.G1, self.a), "b1": pts(_lib.G1, self.b), "b2": pts(_lib.G2, self.b), "h": pts(_lib.G1, self.h),
             "l": pts(_lib.G1, self.l)}
        if keep_host:
            self.host_points = dict(q, g1_singles=singles1, g2_singles=singles2)
        return Parameters(ctx, curve, q["a"], q["b1"], q["b2"], q["h"], q["l"], singles1[0], singles1[1], singles1[2],
                          singles2[0], singles2[1], shard=shard)

    def expected_exponents(self, p, z, h, r, s):
        """(A, B, C) exponents of the proof for assignment z (ints, z[0] = 1), h (ints), r, s --
        prover.rs:164-204 evaluated in Fr."""
        ints = limbs_to_ints
        a, b, l, hq = ints(self.a), ints(self.b), ints(self.l), ints(self.h)
        alpha, beta, delta = (ints(x.reshape(1, 4))[0] for x in (self.alpha, self.beta, self.delta))
        A = (alpha + sum(x * y for x, y in zip(z, a)) + r * delta) % p
        Bv = (beta + sum(x * y for x, y in zip(z, b)) + s * delta) % p
        B1 = Bv if r != 0 else 0                                        # guard of prover.rs:170
        C = (s * A + r * B1 - r * s * delta + sum(x * y for x, y in zip(z[self.n_inputs:], l))
             + sum(x * y for x, y in zip(h, hq))) % p
        return A, Bv, C
