"""The reference's KZG10 polynomial commitment (`zkp_marlin::pc`) on the B200 backend: same structure, argument meaning
and error behaviour as

    KZG10::commit / KZG10::open          marlin/src/pc/kzg10.rs:100-156
    PC::commit / PC::open / batch_open   marlin/src/pc/mod.rs:34-160
    KZG10::check, PC::check / batch_check   kzg10.rs:158-173, pc/mod.rs:102-121,163-240 (pairings on the GPU, batched)
    CommitterKey, LabeledPolynomial, Randomness   marlin/src/pc/data_structures.rs:59-100,146-263

The committer key lives in HBM (`powers_of_g`, `powers_of_gamma_g` uploaded once); commitments and
opening witnesses are MSMs over it (zkb_msm_mont: Montgomery coefficients in, `into_repr` fused), the
witness polynomial p / (x - z), the opening linear combination and the evaluations run on the GPU too.
Polynomials are uint64[n, 4] Montgomery coefficient arrays, low degree first.
"""
import numpy as np

from . import _lib
from .backend import is_dev, point_words, torch
from .r1cs import ints_to_limbs, limbs_to_int

FR_MODULUS = {
    _lib.BLS12_381: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
}


class KzgError(Exception):
    """marlin/src/pc/error.rs"""


class DegreeIsZero(KzgError):
    pass


class DegreeOutOfBound(KzgError):
    pass


class HidingBoundIsZero(KzgError):
    pass


class HidingBoundTooLarge(KzgError):
    pass


class MissingRng(KzgError):
    pass


class MissingPolynomial(KzgError):
    pass


def _nonzero_rows(p):
    """indices of the non-zero coefficients of a host (numpy) or resident (CUDA tensor) polynomial"""
    if is_dev(p):
        return torch.nonzero((p != 0).any(dim=1)).reshape(-1)
    return np.flatnonzero(p.any(axis=1))


def _degree(p):
    nz = _nonzero_rows(p)
    return int(nz[-1]) if len(nz) else 0


def _leading_zeros(p):
    nz = _nonzero_rows(p)
    return int(nz[0]) if len(nz) else len(p)


class CommitterKey:
    """data_structures.rs:59-100 resident on one GPU -- or, with shard = (n_ranks, rank), one contiguous slice of
    powers_of_g per GPU: every rank then runs the same prover (same polynomials, same randomness) and each commitment /
    opening MSM is computed from the ranks' partial sums with one ncclAllGather inside the library (zkb_msm_sharded);
    all ranks obtain the identical commitment, so the Fiat-Shamir transcripts stay in step without further exchange.
    The hiding key powers_of_gamma_g is only ever used for hiding_bound + 1 <= 2 terms and stays whole on every rank."""

    def __init__(self, ctx, curve, powers_of_g, powers_of_gamma_g, supported_degree=None, shard=None):
        """powers_*: (xy uint64[n, words], inf uint8[n]) in the layout of include/zkb.h (the WHOLE key on every rank)"""
        self.ctx, self.curve, self.shard = ctx, curve, shard
        self.n = len(powers_of_g[1])
        self.supported_degree = self.n - 1 if supported_degree is None else supported_degree
        if shard is None:
            self.g = ctx.srs_upload(curve, _lib.G1, powers_of_g[0], powers_of_g[1])
        else:
            from .parallel import shard_range
            lo, hi = shard_range(self.n, shard[0], shard[1])
            self.g = ctx.srs_upload_shard(curve, _lib.G1, powers_of_g[0][lo:hi], powers_of_g[1][lo:hi], lo, self.n)
        n_hiding = self.n_hiding = min(self.n, 1024)
        self.gamma_g = ctx.srs_upload(curve, _lib.G1, powers_of_gamma_g[0][:n_hiding], powers_of_gamma_g[1][:n_hiding])
        self.to_mont = lambda ints: ctx.fr_convert(curve, ints_to_limbs(ints), to_mont=True)

    def msm_g(self, scalars_mont, base_offset=0):
        """multi_scalar_mul(&powers_of_g[base_offset..], scalars) with Montgomery scalars (into_repr fused)"""
        if self.shard is None:
            return self.ctx.msm(self.g, scalars_mont, base_offset=base_offset, mont=True)
        return self.ctx.msm_sharded(self.g, scalars_mont, base_offset=base_offset, mont=True)

    def free(self):
        self.g.free()
        self.gamma_g.free()


class Randomness:
    """kzg10 `Rand` (data_structures.rs:146-200): the blinding polynomial, empty = not hiding"""

    def __init__(self, blinding=None):
        self.blinding = np.zeros((0, 4), dtype=np.uint64) if blinding is None else blinding

    def is_hiding(self):
        return bool(self.blinding.any())

    @staticmethod
    def rand(ck, hiding_bound, rng):
        """DensePolynomial::rand(hiding_bound, rng): hiding_bound + 1 coefficients, drawn in order"""
        p = FR_MODULUS[ck.curve]
        return Randomness(ck.to_mont([rng.randrange(p) for _ in range(hiding_bound + 1)]))


class LabeledPolynomial:
    """data_structures.rs:218-263"""

    def __init__(self, label, coeffs_mont, degree_bound=None, hiding_bound=None):
        self.label = label
        self.coeffs = (coeffs_mont.contiguous() if is_dev(coeffs_mont)
                       else np.ascontiguousarray(coeffs_mont, dtype=np.uint64).reshape(-1, 4))
        self.degree_bound, self.hiding_bound = degree_bound, hiding_bound


def _add_points(ctx, curve, pts):
    """sum of a few affine points on the device (unit-scalar MSM over a throw-away SRS)"""
    xy = np.stack([p[0] for p in pts])
    inf = np.array([1 if p[1] else 0 for p in pts], dtype=np.uint8)
    srs = ctx.srs_upload(curve, _lib.G1, xy, inf, precompute=False)
    try:
        ones = np.zeros((len(pts), 4), dtype=np.uint64)
        ones[:, 0] = 1
        return ctx.msm(srs, ones)
    finally:
        srs.free()


def kzg_commit(ck, p, hiding_bound=None, rng=None, base_offset=0, supported_degree=None):
    """KZG10::commit (kzg10.rs:100-123) over powers_of_g[base_offset..] -> ((xy, is_identity), Randomness)"""
    ctx = ck.ctx
    sup = (ck.n - 1 - base_offset) if supported_degree is None else supported_degree
    deg = _degree(p)
    if deg < 1:
        raise DegreeIsZero()
    if deg > sup:
        raise DegreeOutOfBound()
    nz = _leading_zeros(p)                                       # skip_leading_zeros_and_convert_to_bigints
    comm = ck.msm_g(p[nz:], base_offset + nz)
    rand = Randomness()
    if hiding_bound is not None:
        if rng is None:
            raise MissingRng()
        if hiding_bound == 0:
            raise HidingBoundIsZero()
        if hiding_bound > ck.n - base_offset or hiding_bound + 1 > ck.n_hiding:
            raise HidingBoundTooLarge()
        rand = Randomness.rand(ck, hiding_bound, rng)
        rc = ctx.msm(ck.gamma_g, rand.blinding, mont=True)
        comm = _add_points(ctx, ck.curve, [comm, rc])
    return comm, rand


def kzg_open(ck, p, point_mont, rand):
    """KZG10::open (kzg10.rs:125-156) -> ((w xy, is_identity), rand_v Montgomery or None)"""
    ctx = ck.ctx
    deg = _degree(p)
    if deg < 1:
        raise DegreeIsZero()
    if deg > ck.n:
        raise DegreeOutOfBound()
    witness, _ = ctx.poly_div_linear(ck.curve, p, point_mont)    # compute_witness_polynomial :211-226
    nz = _leading_zeros(witness)
    w = ck.msm_g(witness[nz:], nz)
    rand_v = None
    if rand.is_hiding():
        rq, rand_v = ctx.poly_div_linear(ck.curve, rand.blinding, point_mont)
        w = _add_points(ctx, ck.curve, [w, ctx.msm(ck.gamma_g, rq, mont=True)])
    return w, rand_v


def _open_job(ck, p, point_mont, rand):
    """the checks and divisions of KZG10::open (kzg10.rs:125-156) without the MSMs: -> (jobs, rand_v); the witness is
    the sum of the jobs' MSMs"""
    ctx = ck.ctx
    deg = _degree(p)
    if deg < 1:
        raise DegreeIsZero()
    if deg > ck.n:
        raise DegreeOutOfBound()
    witness, _ = ctx.poly_div_linear(ck.curve, p, point_mont)    # compute_witness_polynomial :211-226
    nz = _leading_zeros(witness)
    jobs = [(ck.g, nz, witness[nz:])]
    rand_v = None
    if rand.is_hiding():
        rq, rand_v = ctx.poly_div_linear(ck.curve, rand.blinding, point_mont)
        jobs.append((ck.gamma_g, 0, rq))
    return jobs, rand_v


def _commit_job(ck, p, hiding_bound, rng, base_offset=0, supported_degree=None):
    """the checks and rng draws of KZG10::commit (kzg10.rs:100-123) without the MSMs: -> (jobs, Randomness) where jobs
    is [(srs, base_offset, scalars)] -- the commitment is the sum of the jobs' MSMs"""
    sup = (ck.n - 1 - base_offset) if supported_degree is None else supported_degree
    deg = _degree(p)
    if deg < 1:
        raise DegreeIsZero()
    if deg > sup:
        raise DegreeOutOfBound()
    nz = _leading_zeros(p)                                       # skip_leading_zeros_and_convert_to_bigints
    jobs = [(ck.g, base_offset + nz, p[nz:])]
    rand = Randomness()
    if hiding_bound is not None:
        if rng is None:
            raise MissingRng()
        if hiding_bound == 0:
            raise HidingBoundIsZero()
        if hiding_bound > ck.n - base_offset or hiding_bound + 1 > ck.n_hiding:
            raise HidingBoundTooLarge()
        rand = Randomness.rand(ck, hiding_bound, rng)
        jobs.append((ck.gamma_g, 0, rand.blinding))
    return jobs, rand


def _add_point_groups(ctx, curve, groups):
    """[[points]] -> [sum of each group] with one upload and one batched call (unit-scalar MSMs over a throw-away SRS)"""
    flat = [p for g in groups for p in g]
    xy = np.stack([p[0] for p in flat])
    inf = np.array([1 if p[1] else 0 for p in flat], dtype=np.uint8)
    srs = ctx.srs_upload(curve, _lib.G1, xy, inf, precompute=False)
    try:
        ones = np.zeros((max(len(g) for g in groups), 4), dtype=np.uint64)
        ones[:, 0] = 1
        offs, pos = [], 0
        for g in groups:
            offs.append(pos)
            pos += len(g)
        return ctx.msm_batch([srs] * len(groups), [ones[:len(g)] for g in groups], offs)
    finally:
        srs.free()


def pc_commit(ck, polynomials, rng=None):
    """PC::commit (pc/mod.rs:34-71) -> ([(comm, shifted_comm or None)], [(rand, shifted_rand or None)]).
    The reference commits polynomial by polynomial; the rng draws happen in that order here too, while the MSMs of all
    the commitments of the call go to the device together (zkb_msm_batch: their sorts, accumulations and reductions
    overlap on the side streams).  With a sharded committer key every MSM is a collective and they run one by one."""
    if ck.shard is not None:
        return _pc_commit_serial(ck, polynomials, rng)
    ctx = ck.ctx
    jobs, slots, rands = [], [], []            # slots[i] = (job indices of comm, job indices of shifted comm or None)
    for P in polynomials:
        j, rand = _commit_job(ck, P.coeffs, P.hiding_bound, rng, supported_degree=ck.supported_degree)
        main = list(range(len(jobs), len(jobs) + len(j)))
        jobs += j
        shifted, shifted_rand = None, None
        if P.degree_bound is not None:
            if P.degree_bound > ck.supported_degree:
                raise DegreeOutOfBound()
            off = ck.supported_degree - P.degree_bound              # shifted_powers (data_structures.rs:87-99)
            j, shifted_rand = _commit_job(ck, P.coeffs, P.hiding_bound, rng, base_offset=off, supported_degree=P.degree_bound)
            shifted = list(range(len(jobs), len(jobs) + len(j)))
            jobs += j
        slots.append((main, shifted))
        rands.append((rand, shifted_rand))
    pts = ctx.msm_batch([j[0] for j in jobs], [j[2] for j in jobs], [j[1] for j in jobs], mont=True) if jobs else []
    groups = [idx for pair in slots for idx in pair if idx is not None and len(idx) > 1]
    sums = iter(_add_point_groups(ctx, ck.curve, [[pts[i] for i in idx] for idx in groups]) if groups else [])
    value = lambda idx: None if idx is None else (pts[idx[0]] if len(idx) == 1 else next(sums))
    comms = []
    for main, shifted in slots:                 # same order as `groups` was built in
        c = value(main)
        comms.append((c, value(shifted)))
    return comms, rands


def _pc_commit_serial(ck, polynomials, rng=None):
    comms, rands = [], []
    for P in polynomials:
        comm, rand = kzg_commit(ck, P.coeffs, P.hiding_bound, rng, supported_degree=ck.supported_degree)
        shifted, shifted_rand = None, None
        if P.degree_bound is not None:
            if P.degree_bound > ck.supported_degree:
                raise DegreeOutOfBound()
            off = ck.supported_degree - P.degree_bound              # shifted_powers (data_structures.rs:87-99)
            shifted, shifted_rand = kzg_commit(ck, P.coeffs, P.hiding_bound, rng, base_offset=off,
                                               supported_degree=P.degree_bound)
        comms.append((comm, shifted))
        rands.append((rand, shifted_rand))
    return comms, rands


def _open_combination(ck, polynomials, opening_challenge, randomnesses):
    """the linear combination PC::open builds before KZG10::open (pc/mod.rs:81-98) -> (p, Randomness)"""
    ctx = ck.ctx
    mod = FR_MODULUS[ck.curve]
    polys, shifts, coeffs = [], [], []
    rpolys, rcoeffs = [], []
    challenge = 1
    for P, (rand, shifted_rand) in zip(polynomials, randomnesses):
        polys.append(P.coeffs); shifts.append(0); coeffs.append(challenge)
        if len(rand.blinding):
            rpolys.append(rand.blinding); rcoeffs.append(challenge)
        if P.degree_bound is not None:
            sc = challenge * opening_challenge % mod
            if bool(P.coeffs.any()):                                 # shift_polynomial (:241-250)
                polys.append(P.coeffs); shifts.append(ck.supported_degree - P.degree_bound); coeffs.append(sc)
            if shifted_rand is not None and len(shifted_rand.blinding):
                rpolys.append(shifted_rand.blinding); rcoeffs.append(sc)
        challenge = challenge * opening_challenge % mod * opening_challenge % mod
    p = ctx.poly_lincomb(ck.curve, polys, ck.to_mont(coeffs), shifts)
    r = Randomness(ctx.poly_lincomb(ck.curve, rpolys, ck.to_mont(rcoeffs)) if rpolys else None)
    return p, r


def pc_open(ck, polynomials, point_mont, opening_challenge, randomnesses):
    """PC::open (pc/mod.rs:73-100); opening_challenge is a canonical int"""
    p, r = _open_combination(ck, polynomials, opening_challenge, randomnesses)
    return kzg_open(ck, p, point_mont, r)


def pc_batch_open(ck, polynomials, query_set, opening_challenge, randomnesses):
    """PC::batch_open (pc/mod.rs:122-160).  query_set: iterable of (label, point canonical int); proofs are
    returned in increasing order of the point (BTreeMap order), labels sorted inside a point."""
    by_label = {P.label: (P, r) for P, r in zip(polynomials, randomnesses)}
    point_to_labels = {}
    for label, point in query_set:
        point_to_labels.setdefault(point, set()).add(label)
    groups = []
    for point in sorted(point_to_labels):
        polys, rands = [], []
        for label in sorted(point_to_labels[point]):
            if label not in by_label:
                raise MissingPolynomial(label)
            polys.append(by_label[label][0])
            rands.append(by_label[label][1])
        groups.append((polys, rands, point))
    if ck.shard is not None:                                        # every MSM is a collective: one by one
        return [pc_open(ck, polys, ck.to_mont([point])[0], opening_challenge, rands) for polys, rands, point in groups]
    # the witness MSMs of all query points go to the device together (zkb_msm_batch), like the commitments of a round
    ctx = ck.ctx
    jobs, slots, rand_vs = [], [], []
    for polys, rands, point in groups:
        p, r = _open_combination(ck, polys, opening_challenge, rands)
        j, rand_v = _open_job(ck, p, ck.to_mont([point])[0], r)
        slots.append(list(range(len(jobs), len(jobs) + len(j))))
        jobs += j
        rand_vs.append(rand_v)
    pts = ctx.msm_batch([j[0] for j in jobs], [j[2] for j in jobs], [j[1] for j in jobs], mont=True) if jobs else []
    multi = [idx for idx in slots if len(idx) > 1]
    sums = iter(_add_point_groups(ctx, ck.curve, [[pts[i] for i in idx] for idx in multi]) if multi else [])
    return [(pts[idx[0]] if len(idx) == 1 else next(sums), rv) for idx, rv in zip(slots, rand_vs)]


# ------------------------------------------------------------------------------------------------
# verifier side: KZG10::check (kzg10.rs:158-173), PC::check / batch_check (pc/mod.rs:102-121,163-202)
# ------------------------------------------------------------------------------------------------
class MissingEvaluation(KzgError):
    pass


def _lincomb_points(ctx, curve, group, jobs):
    """[sum_i k_i * P_i] for jobs = [([(xy, is_identity)], [canonical int])]: throw-away base sets, one batched MSM call"""
    if not jobs:
        return []
    srs = []
    try:
        for pts, _ in jobs:
            srs.append(ctx.srs_upload(curve, group, np.stack([np.asarray(p[0], dtype=np.uint64).reshape(-1) for p in pts]),
                                      np.array([1 if p[1] else 0 for p in pts], dtype=np.uint8), precompute=False))
        mod = FR_MODULUS[curve]
        return ctx.msm_batch(srs, [ints_to_limbs([k % mod for k in ks]) for _, ks in jobs])
    finally:
        for s in srs:
            s.free()


def _check_many(ctx, vk, items):
    """items: [(terms, point, proof)] with terms = [(G1 point, canonical int)] summing to comm - value * g (before the
    hiding term) -> [bool].  KZG10::check (kzg10.rs:158-173) tests e(u, h) == e(w, beta_h - point * h); here every item
    is one group (u, h), (-w, beta_h - point * h) of a single zkb_multi_pairing call, compared with 1."""
    from . import pairing as _pairing
    curve = vk.curve
    mod = FR_MODULUS[curve]
    rinv = pow(1 << 256, -1, mod)
    g1_jobs, g2_jobs = [], []
    for terms, point, (w, rand_v) in items:
        pts, ks = [t[0] for t in terms], [t[1] for t in terms]
        if rand_v is not None:                                         # u -= gamma_g * rand_v (:166-168)
            pts.append(vk.gamma_g)
            ks.append(-(limbs_to_int(rand_v) * rinv % mod))
        g1_jobs.append((pts, ks))
        g2_jobs.append(([vk.beta_h, vk.h], [1, -point]))               # :169
    us = _lincomb_points(ctx, curve, _lib.G1, g1_jobs)
    vs = _lincomb_points(ctx, curve, _lib.G2, g2_jobs)
    groups = [[(u, vk.h), (_pairing.neg_point(curve, _lib.G1, it[2][0]), v)] for u, v, it in zip(us, vs, items)]
    one = _pairing.gt_one(curve)
    return [bool(np.array_equal(t, one)) for t in _pairing.multi_pairing(ctx, curve, groups)]


def kzg_check(ctx, vk, comm, point, value, proof):
    """KZG10::check (kzg10.rs:158-173).  point, value: canonical ints; proof = (w, rand_v Montgomery limbs or None)"""
    return _check_many(ctx, vk, [([(comm, 1), (vk.g, -value)], point, proof)])[0]


def _accumulate(vk, commitments, degree_bounds, point, values, opening_challenge):
    """accumulate_commitments_and_values (pc/mod.rs:204-240) as the terms of one linear combination"""
    mod = FR_MODULUS[vk.curve]
    terms, acc_v, ch = [], 0, 1
    for (comm, shifted), db, v in zip(commitments, degree_bounds, values):
        assert (db is not None) == (shifted is not None)
        terms.append((comm, ch))
        acc_v = (acc_v + v * ch) % mod
        if db is not None:
            sc = ch * opening_challenge % mod
            terms.append((shifted, sc))
            acc_v = (acc_v + pow(point, vk.supported_degree - db, mod) * v % mod * sc) % mod
        ch = ch * opening_challenge % mod * opening_challenge % mod
    terms.append((vk.g, -acc_v))
    return terms


def pc_check(ctx, vk, commitments, degree_bounds, point, values, proof, opening_challenge):
    """PC::check (pc/mod.rs:102-121): commitments = [(comm, shifted or None)], values / point / challenge canonical ints"""
    return _check_many(ctx, vk, [(_accumulate(vk, commitments, degree_bounds, point, values, opening_challenge), point, proof)])[0]


def pc_batch_check(ctx, vk, commitments, query_set, values, proofs, opening_challenge):
    """PC::batch_check (pc/mod.rs:163-202).  commitments: {label: ((comm, shifted or None), degree_bound)},
    query_set: iterable of (label, point), values: {(label, point): canonical int}, proofs in increasing point order.
    All query points' pairings run in one device call; the result is the conjunction (result &= ..., :199)."""
    point_to_labels = {}
    for label, point in query_set:
        point_to_labels.setdefault(point, set()).add(label)
    assert len(point_to_labels) == len(proofs)
    items = []
    for point, proof in zip(sorted(point_to_labels), proofs):
        cs, dbs, vs = [], [], []
        for label in sorted(point_to_labels[point]):
            if label not in commitments:
                raise MissingPolynomial(label)
            if (label, point) not in values:
                raise MissingEvaluation(label)
            cs.append(commitments[label][0])
            dbs.append(commitments[label][1])
            vs.append(values[(label, point)])
        items.append((_accumulate(vk, cs, dbs, point, vs, opening_challenge), point, proof))
    return all(_check_many(ctx, vk, items))
