"""csrc/pairing.cuh compiled for the host (tests/host_emu): the exact device code of the Miller loop and of the final
exponentiation against the oracle's pairing (oracle/pyref/pairing.py), without a GPU.  The affine Miller-loop value is the
oracle's bit for bit (same steps, same line scaling); the inversion-free loop the kernels run differs from it by a factor in
Fq2* and must give the same pairing; the final exponentiation is the oracle's plain power
f^((q^12 - 1) / r) raised to m = 3 on BLS12-381 (x-chain of the hard part) and m = 1 on BN254 (exact chain).  On BN254 the
device runs the optimal ate loop (6x + 2, then the lines through pi(Q) and -pi^2(Q)), restated in the oracle as
miller_loop_optimal_bn; the plain ate pairing of the oracle's verifiers is another power of the same pairing."""
import random

import numpy as np
import pytest

from oracle.pyref import pairing as OP
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FQ
from tests import emu

W_POWER = [0, 2, 4, 1, 3, 5]


def tower_to_flat(cid, t):
    q, c = FQ[cid].p, OP._C[cid]
    flat = [0] * 12
    for s in range(6):
        x, y, k = t[2 * s], t[2 * s + 1], W_POWER[s]
        flat[k] = (flat[k] + x - c * y) % q
        flat[k + 6] = (flat[k + 6] + y) % q
    return flat


def flat_to_tower(cid, fl):
    q, c = FQ[cid].p, OP._C[cid]
    t = [0] * 12
    for s in range(6):
        k = W_POWER[s]
        t[2 * s], t[2 * s + 1] = (fl[k] + c * fl[k + 6]) % q, fl[k + 6]
    return t


def run(lib, cid, stage, P, Q, f_in=None):
    n = FQ[cid].limbs * 2
    q = FQ[cid].p
    mont = lambda v: v * (1 << (32 * n)) % q
    pa = np.concatenate([emu.to_u32(mont(v), n) for v in (P if P else (0, 0))])
    qa = np.concatenate([emu.to_u32(mont(v), n) for v in ((Q[0][0], Q[0][1], Q[1][0], Q[1][1]) if Q else (0, 0, 0, 0))])
    fa = np.concatenate([emu.to_u32(v, n) for v in (f_in if f_in else [0] * 12)])
    out = np.zeros(12 * n, dtype=np.uint32)
    lib.emu_pairing(0 if cid == BN254 else 1, stage, emu.ptr(pa), emu.ptr(qa), emu.ptr(fa), emu.ptr(out))
    return [emu.from_u32(out[i * n:(i + 1) * n]) for i in range(12)]


@pytest.fixture(scope="module")
def lib():
    return emu.build()


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_miller_loop_and_final_exponentiation_match_oracle(lib, cid):
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    rng = random.Random(9 + cid)
    m = 3 if cid == BLS12_381 else 1
    F12 = OP.Fq12(cid)
    for _ in range(3):
        P = g1.mul_affine(g1.gen, rng.randrange(1, g1.r))
        Q = g2.mul_affine(g2.gen, rng.randrange(1, g2.r))
        want = OP.device_miller_loop(cid, P, Q)                           # optimal ate on BN254, plain ate on BLS12-381
        assert tower_to_flat(cid, run(lib, cid, 0, P, Q)) == want
        fe = F12.pow(OP.final_exponentiation(cid, want), m)
        assert fe == OP.device_multi_pairing(cid, [(P, Q)])
        assert tower_to_flat(cid, run(lib, cid, 1, None, None, flat_to_tower(cid, want))) == fe
        assert tower_to_flat(cid, run(lib, cid, 2, P, Q)) == fe          # the inversion-free loop of the kernels: same pairing
        assert tower_to_flat(cid, run(lib, cid, 3, P, Q)) == fe          # affine loop + final exponentiation
    # identity on either side: the loop value is 1, and so is the pairing
    assert tower_to_flat(cid, run(lib, cid, 0, None, Q)) == F12.one
    assert tower_to_flat(cid, run(lib, cid, 2, P, None)) == F12.one


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_bilinear(lib, cid):
    g1, g2 = CURVES[(cid, 1)], CURVES[(cid, 2)]
    a, b = 0x1234567, 0x89ABCDEF1
    e1 = run(lib, cid, 2, g1.mul_affine(g1.gen, a), g2.mul_affine(g2.gen, b))
    e2 = run(lib, cid, 2, g1.mul_affine(g1.gen, a * b % g1.r), g2.gen)
    assert e1 == e2 and e1 != flat_to_tower(cid, OP.Fq12(cid).one)


def test_pairing_params_header_is_generated():
    """csrc/pairing_params.cuh (loop counts, Frobenius constants) is what tools/gen_pairing_params.py derives from the moduli"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.call([sys.executable, os.path.join(root, "tools", "gen_pairing_params.py"), "--check"]) == 0

