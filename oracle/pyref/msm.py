"""Variable-base MSM: restatement of ark-ec 0.2 `VariableBaseMSM::multi_scalar_mul`.

The algorithm is NOT in /root/reference (un-vendored `ark-ec = "0.2"`, see
SURVEY.md section 8c); it is anchored on the reference's call sites
groth16/src/prover.rs:187,190,220, marlin/src/pc/kzg10.rs:109,118,137,146 and
curve/src/lib.rs:44.  Published algorithm: Pippenger with window
c = 3 if n < 32 else floor(ceil(log2 n) * 69 / 100) + 2, ceil(bits/c) windows,
2^c - 1 Jacobian buckets per window, zero scalars skipped, unit scalars added once
in window 0, running-sum bucket reduction, windows combined high->low.
"""


def ark_log2(x):
    """ark_std::log2: ceil(log2 x) for x > 0."""
    if x == 0:
        return 0
    return (x - 1).bit_length() if x & (x - 1) else x.bit_length() - 1


def ark_window(n):
    return 3 if n < 32 else (ark_log2(n) * 69 // 100) + 2


def msm_naive(curve, bases, scalars):
    """sum s_i * P_i by double-and-add (first-principles check)."""
    acc = curve.identity()
    for P, s in zip(bases, scalars):
        if P is None or s == 0:
            continue
        acc = curve.add(acc, curve.mul(curve.from_affine(P), s))
    return acc


def msm_pippenger(curve, bases, scalars, modulus_bits):
    """Literal restatement of the ark-ec 0.2 bucket method.  Returns Jacobian."""
    n = min(len(bases), len(scalars))
    bases, scalars = bases[:n], scalars[:n]
    c = ark_window(n)
    pairs = [(s, P) for s, P in zip(scalars, bases) if s != 0]
    window_sums = []
    for w_start in range(0, modulus_bits, c):
        res = curve.identity()
        buckets = [curve.identity() for _ in range((1 << c) - 1)]
        for s, P in pairs:
            if s == 1:
                if w_start == 0:
                    res = curve.add_mixed(res, P)
            else:
                d = (s >> w_start) & ((1 << c) - 1)   # low 64-bit limb % 2^c, c <= 64
                if d:
                    buckets[d - 1] = curve.add_mixed(buckets[d - 1], P)
        running = curve.identity()
        for b in reversed(buckets):
            running = curve.add(running, b)
            res = curve.add(res, running)
        window_sums.append(res)
    total = curve.identity()
    for s in reversed(window_sums[1:]):
        total = curve.add(total, s)
        for _ in range(c):
            total = curve.dbl(total)
    return curve.add(window_sums[0], total)


def msm_g1_adds(n_nonzero, n, modulus_bits):
    """'G1-adds' of the reference algorithm: mixed adds + reduction adds."""
    c = ark_window(n)
    w = (modulus_bits + c - 1) // c
    return n_nonzero * w + 2 * ((1 << c) - 1) * w
