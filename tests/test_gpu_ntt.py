"""Radix-2 NTT on the GPU vs the oracle's restatement of ark-poly 0.2 Radix2EvaluationDomain
(call sites groth16/src/r1cs_to_qap.rs:144-169)."""
import random

import numpy as np
import pytest

from oracle.pyref.fields import BLS12_381, BN254, FR
from oracle.pyref.ntt import Domain, dft_naive
from tests import helpers as H

pytestmark = pytest.mark.gpu


def run(ctx, cid, vals, log_n, **kw):
    arr = H.fr_array(cid, vals)
    ctx.ntt(cid, arr, log_n, **kw)
    return H.fr_ints(cid, arr)


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 9, 10, 11, 12, 13, 14])
def test_ntt_matches_oracle(ctx, cid, log_n):
    fr = FR[cid]
    rng = random.Random(cid * 50 + log_n)
    n = 1 << log_n
    vals = [rng.randrange(fr.p) for _ in range(n)]
    if n >= 4:
        vals[0], vals[1], vals[2] = 0, 1, fr.p - 1
    d = Domain(fr, n)
    assert d.size == n
    assert run(ctx, cid, vals, log_n) == d.fft(vals)
    assert run(ctx, cid, vals, log_n, inverse=True) == d.ifft(vals)
    assert run(ctx, cid, vals, log_n, coset=True) == d.coset_fft(vals)
    assert run(ctx, cid, vals, log_n, inverse=True, coset=True) == d.coset_ifft(vals)


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_ntt_is_the_dft(ctx, cid):
    """first-principles O(n^2) evaluation at the domain's root of unity"""
    fr = FR[cid]
    rng = random.Random(3)
    vals = [rng.randrange(fr.p) for _ in range(64)]
    assert run(ctx, cid, vals, 6) == dft_naive(vals, fr.root_of_unity(6), fr.p)


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
@pytest.mark.parametrize("log_n", [16, 19, 20, 21, 22, 24])
def test_ntt_large_properties(ctx, cid, log_n):
    """Sizes of the BASELINE sweep: round trips, and evaluation of a sparse polynomial whose
    transform is known in closed form (a*x^j -> a*w^(ij))."""
    fr = FR[cid]
    p, n = fr.p, 1 << log_n
    rng = np.random.default_rng(log_n)
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)       # arbitrary residues < p (Montgomery form of something)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    orig = a.copy()
    ctx.ntt(cid, a, log_n)
    assert not np.array_equal(a, orig)
    ctx.ntt(cid, a, log_n, inverse=True)
    assert np.array_equal(a, orig)
    ctx.ntt(cid, a, log_n, coset=True)
    ctx.ntt(cid, a, log_n, inverse=True, coset=True)
    assert np.array_equal(a, orig)
    # sparse polynomial c0 + c1 x^j
    j = (n // 3) | 1
    c0, c1 = 12345, 67890
    poly = np.zeros((n, 4), dtype=np.uint64)
    poly[0] = H.fr_array(cid, [c0])[0]
    poly[j] = H.fr_array(cid, [c1])[0]
    ctx.ntt(cid, poly, log_n)
    w = fr.root_of_unity(log_n)
    for i in (0, 1, 2, n // 2, n - 1, 12345 % n):
        assert H.fr_ints(cid, poly[i:i + 1])[0] == (c0 + c1 * pow(w, i * j, p)) % p
    # same on the coset g*H
    poly[:] = 0
    poly[0] = H.fr_array(cid, [c0])[0]
    poly[j] = H.fr_array(cid, [c1])[0]
    ctx.ntt(cid, poly, log_n, coset=True)
    g = fr.generator
    for i in (0, 1, n // 2, n - 1):
        assert H.fr_ints(cid, poly[i:i + 1])[0] == (c0 + c1 * pow(g * pow(w, i, p) % p, j, p)) % p


def test_ntt_too_large(ctx):
    """EvaluationDomain::new -> None -> SynthesisError::PolynomialDegreeTooLarge (r1cs/src/error.rs:15)"""
    from ckb_zkp_b200.backend import ZkbError
    arr = np.zeros((2, 4), dtype=np.uint64)
    with pytest.raises((ZkbError, ValueError)):
        ctx.ntt(BN254, arr, 29)


@pytest.mark.parametrize("cid,log_n", [(BN254, 17), (BLS12_381, 21)])
def test_ntt_large_vs_cpu_restatement(ctx, cid, log_n):
    """bit-exact against the C++ restatement of ark-poly's radix-2 domain at sizes the Python oracle
    cannot reach (2^21 = the domain of the 2^20-constraint proof)"""
    from oracle import cref
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    for kw in ({}, {"inverse": True, "coset": True}):
        got = ctx.ntt(cid, a.copy(), log_n, **kw)
        want = cref.ntt(cid, a.copy(), log_n, **kw)
        assert np.array_equal(got, want), kw
