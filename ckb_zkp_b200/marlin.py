"""Marlin's AHP prover rounds (`zkp_marlin::ahp`, marlin/src/ahp/prover.rs:86-427) on the B200 backend.

Same round structure, argument meaning and outputs as the reference:

    prover_init          prover.rs:86-147     z_A = A z, z_B = B z (device SpMV)
    prover_first_round   prover.rs:150-222    w, z_a, z_b, mask
    prover_second_round  prover.rs:230-321    t, g_1, h_1
    prover_third_round   prover.rs:331-427    g_2, h_2

Every transform (interpolate / evaluate_over_domain = zkb_ntt), pointwise loop (zkb_fr_vec_op), batch
inversion (zkb_fr_batch_inverse), sparse accumulation (zkb_spmv) and domain-element table (zkb_fr_powers)
runs on the GPU; commitments and openings go through `kzg10.py`.  Data stays in host arrays between
calls in this round of the build (one H2D/D2H per primitive) -- correct, not yet fast; keeping the round
state resident is listed as next work in DESIGN.md.

The index (indexer.rs:71-116) is an input, as in the reference where `index()` runs once per circuit;
randomness and verifier challenges are explicit arguments (the Fiat-Shamir byte stream of
marlin/src/fs_rng.rs is produced by the Rust host).  Polynomials are uint64[n, 4] Montgomery coefficient
arrays, low degree first, trimmed like ark-poly's DensePolynomial.
"""
import numpy as np

from . import _lib
from .backend import CsrMatrix, is_dev, torch
from .r1cs import ints_to_limbs

FR_MODULUS = {
    _lib.BLS12_381: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    _lib.BN254: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
}
FR_GENERATOR = {_lib.BLS12_381: 7, _lib.BN254: 5}
FR_TWO_ADICITY = {_lib.BLS12_381: 32, _lib.BN254: 28}
R = 1 << 256


class PolynomialDegreeTooLarge(Exception):
    """SynthesisError::PolynomialDegreeTooLarge (EvaluationDomain::new -> None)"""


class InstanceDoesNotMatchIndex(Exception):
    pass


def domain_size(n):
    """GeneralEvaluationDomain::compute_size_of_domain: next power of two"""
    return 1 << max(n - 1, 0).bit_length()


def reindex_by_subdomain(h_size, x_size, index):
    """ark-poly EvaluationDomain::reindex_by_subdomain (vectorised over numpy index arrays)"""
    index = np.asarray(index, dtype=np.int64)
    period = h_size // x_size
    i = index - x_size
    return np.where(index < x_size, index * period, i + i // max(period - 1, 1) + 1)


class Field:
    """scalar-side helper: canonical ints <-> Montgomery limbs for a handful of values"""

    def __init__(self, curve):
        self.curve, self.p = curve, FR_MODULUS[curve]
        self.rinv = pow(R, -1, self.p)

    def mont(self, x):
        return ints_to_limbs([x % self.p * R % self.p])[0]

    def mont_arr(self, xs):
        return ints_to_limbs([x % self.p * R % self.p for x in xs])

    def to_int(self, limbs):
        return int.from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes(), "little") * self.rinv % self.p

    def root_of_unity(self, size):
        log = size.bit_length() - 1
        if log > FR_TWO_ADICITY[self.curve]:
            raise PolynomialDegreeTooLarge()
        w = pow(FR_GENERATOR[self.curve], (self.p - 1) >> FR_TWO_ADICITY[self.curve], self.p)
        for _ in range(log, FR_TWO_ADICITY[self.curve]):
            w = w * w % self.p
        return w

    def vanishing_at(self, size, x):
        return (pow(x, size, self.p) - 1) % self.p


def trim(p):
    """drop leading (high-degree) zero coefficients, like DensePolynomial::from_coefficients_vec"""
    if is_dev(p):
        nz = torch.nonzero((p != 0).any(dim=1))
        return p[:int(nz[-1]) + 1] if len(nz) else p[:0]
    nz = np.flatnonzero(p.any(axis=1))
    return p[:int(nz[-1]) + 1] if len(nz) else p[:0]


def pad(p, n):
    """a fresh array of exactly n coefficients (always a copy: the transforms work in place)"""
    if is_dev(p):
        if len(p) >= n:
            return p[:n].clone()
        out = torch.zeros((n, 4), dtype=torch.int64, device=p.device)
        out[:len(p)] = p
        return out
    if len(p) >= n:
        return p[:n].copy()
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:len(p)] = p
    return out


def contig(a):
    return a.contiguous() if is_dev(a) else np.ascontiguousarray(a)


def to_host(a):
    """uint64[n, 4] numpy copy of a host or device Fr vector"""
    return a.cpu().numpy().view(np.uint64) if is_dev(a) else np.asarray(a)


class Ops:
    """device-backed polynomial / vector arithmetic for one (ctx, curve)"""

    def __init__(self, ctx, curve, resident=False):
        """resident=True keeps every vector in HBM (CUDA int64[n, 4] tensors): the primitives then exchange device
        pointers and only the few scalars that feed the transcript travel to the host.  torch's current stream is
        pointed at the library's stream so tensor glue (slicing, concatenation, zero fill) and kernels stay ordered."""
        self.ctx, self.curve, self.f = ctx, curve, Field(curve)
        self.device, self._prev_stream = None, None
        if resident:
            self.device = torch.device("cuda", torch.cuda.current_device())
            self._prev_stream = torch.cuda.current_stream(self.device)
            torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream, device=self.device))

    def release(self):
        """give torch's current stream back to the caller (create_random_proof does this when the proof is done; callers of
        the round-level API with resident=True call it themselves once they are finished with the round state)"""
        if self._prev_stream is not None:
            self.ctx.sync()
            torch.cuda.set_stream(self._prev_stream)
            self._prev_stream = None

    # array plumbing on whichever side the vectors live
    def zeros(self, n):
        if self.device is not None:
            return torch.zeros((n, 4), dtype=torch.int64, device=self.device)
        return np.zeros((n, 4), dtype=np.uint64)

    def put(self, a):
        """host uint64[n, 4] -> the side this Ops works on"""
        if self.device is None or is_dev(a):
            return a
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(self.device)

    def cat(self, parts):
        return torch.cat(parts) if is_dev(parts[0]) else np.concatenate(parts)

    def take(self, a, idx):
        """a[idx] for a host index array"""
        if is_dev(a):
            return a[torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(a.device)]
        return np.ascontiguousarray(a[idx])

    def zero_rows(self, a, mask):
        """a[mask] = 0 for a host boolean mask"""
        if is_dev(a):
            a[torch.from_numpy(mask).to(a.device)] = 0
        else:
            a[mask] = 0
        return a

    def _v(self, op, a, b=None, s=None):
        return self.ctx.fr_vec_op(self.curve, op, a, b, None if s is None else self.f.mont(s))

    def add(self, a, b): return self._v(self.ctx.VEC_ADD, a, b)
    def sub(self, a, b): return self._v(self.ctx.VEC_SUB, a, b)
    def mul(self, a, b): return self._v(self.ctx.VEC_MUL, a, b)
    def scale(self, a, s): return self._v(self.ctx.VEC_SCALE, a, s=s)
    def axpy(self, a, s, b): return self._v(self.ctx.VEC_AXPY, a, b, s)          # a + s * b
    def rsub(self, s, a): return self._v(self.ctx.VEC_RSUB, a, s=s)              # s - a
    def addc(self, a, s): return self._v(self.ctx.VEC_ADDC, a, s=s)              # a + s
    def inv(self, a): return self.ctx.fr_batch_inverse(self.curve, a)

    def fft(self, coeffs, size):
        """domain.fft(&coeffs): evaluations over the size-`size` domain"""
        a = pad(coeffs, size)
        return self.ctx.ntt(self.curve, a, size.bit_length() - 1)

    def ifft(self, evals, size):
        """Evaluations::interpolate"""
        a = pad(evals, size)
        return self.ctx.ntt(self.curve, a, size.bit_length() - 1, inverse=True)

    def elements(self, size):
        """domain.elements(): w^i"""
        return self.ctx.fr_powers(self.curve, self.f.mont(self.f.root_of_unity(size)), size, device=self.device)

    def poly_add(self, a, b):
        n = max(len(a), len(b))
        return trim(self.add(pad(a, n), pad(b, n))) if n else a[:0]

    def poly_sub(self, a, b):
        n = max(len(a), len(b))
        return trim(self.sub(pad(a, n), pad(b, n))) if n else a[:0]

    def poly_mul(self, a, b):
        a, b = trim(a), trim(b)
        if not len(a) or not len(b):
            return a[:0]
        size = domain_size(len(a) + len(b) - 1)
        return trim(self.ifft(self.mul(self.fft(a, size), self.fft(b, size)), size))

    def divide_by_vanishing_poly(self, p, n):
        """DensePolynomial::divide_by_vanishing_poly for x^n - 1 -> (quotient, remainder)"""
        if len(p) < n:
            return p[:0], trim(p)
        # q[j] = sum_{i >= 1} p[j + i n]: a suffix sum with stride n, by doubling (log2(len / n) vector additions
        # instead of len / n -- the division of w by v_X has n = |X| = 2 and len = |H|)
        q = pad(p[n:], len(p) - n)
        s = n
        while s < len(q):
            m = len(q) - s
            q[:m] = self.add(contig(q[:m]), contig(q[s:]))
            s *= 2
        r = pad(p[:n], n)
        k = min(n, len(q))
        if k:
            r[:k] = self.add(contig(r[:k]), contig(q[:k]))
        return trim(q), trim(r)

    def mul_by_vanishing_poly(self, p, n):
        out = self.zeros(len(p) + n)
        out[n:] = p
        if len(p):
            out[:len(p)] = self.sub(contig(out[:len(p)]), contig(p))
        return trim(out)

    def add_const_terms(self, p, terms):
        """p + sum c * x^k for a few (k, c) pairs with canonical int c (host-side scalar arithmetic)"""
        n = max([len(p)] + [k + 1 for k, _ in terms])
        out = pad(p, n)
        for k, c in terms:
            cur = self.f.to_int(to_host(out[k:k + 1])[0])
            out[k:k + 1] = self.put(self.f.mont(cur + c).reshape(1, 4))
        return trim(out)

    def batch_evals(self, size, x):
        """arithmetic.rs:28-34: v_H(x) / (x - w^i)"""
        den = self.inv(self.rsub(x, self.elements(size)))
        return self.scale(den, self.f.vanishing_at(size, x))


class Index:
    """What the prover reads from `Index` (indexer.rs:29-69): the three square matrices and, per
    matrix, the arithmetisation's evaluations on K and on B (arithmetic.rs:97-172)."""

    def __init__(self, curve, num_constraints, num_variables, num_non_zeros, num_inputs, matrices, stars):
        """matrices: {'a'|'b'|'c': CsrMatrix over the formatted variable numbering}
        stars: {'a'|'b'|'c': {'row_evals_on_k', 'col_evals_on_k', 'val_evals_on_k', 'row_evals_on_b',
                'col_evals_on_b', 'val_evals_on_b', 'row_col_evals_on_b'}} Montgomery arrays"""
        self.curve = curve
        self.num_constraints, self.num_variables, self.num_non_zeros = num_constraints, num_variables, num_non_zeros
        self.x_size, self.h_size, self.k_size = domain_size(num_inputs), domain_size(num_variables), domain_size(num_non_zeros)
        self.b_size = domain_size(3 * self.k_size - 3)
        self.matrices, self.stars = matrices, stars
        for s in (self.x_size, self.h_size, self.k_size, self.b_size):
            if s.bit_length() - 1 > FR_TWO_ADICITY[curve]:
                raise PolynomialDegreeTooLarge()
        # transposed, H-reindexed matrices for the t accumulation of round 2 (prover.rs:259-269)
        self.transposed = {}
        for name, m in matrices.items():
            rows = np.repeat(np.arange(m.n_rows, dtype=np.int64), np.diff(m.row_ptr.astype(np.int64)))
            k = reindex_by_subdomain(self.h_size, self.x_size, m.col_idx.astype(np.int64))
            order = np.argsort(k, kind="stable")
            ptr = np.zeros(self.h_size + 1, dtype=np.uint32)
            np.cumsum(np.bincount(k, minlength=self.h_size), out=ptr[1:])
            self.transposed[name] = CsrMatrix(ptr, rows[order].astype(np.uint32), m.coeff[order])
        self.resident = False

    def make_resident(self, ops):
        """move the matrices and the evaluation tables into HBM once (the index outlives many proofs)"""
        if self.resident:
            return
        for m in list(self.matrices.values()) + list(self.transposed.values()):
            m.to_device(ops.device)
        for star in self.stars.values():
            for key in ("row_evals_on_k", "col_evals_on_k", "val_evals_on_k", "row_evals_on_b", "col_evals_on_b",
                        "val_evals_on_b", "row_col_evals_on_b", "row", "col", "val", "row_col"):
                if key in star:
                    star[key] = ops.put(star[key])
        self.resident = True


def make_matrices_square(mats, num_variables):
    """constraint_systems.rs:9-31 on CSR matrices: pad with empty constraints (or report how many padding
    variables the caller must append to the witness) so that #constraints == #variables."""
    nc = mats[0].n_rows
    extra_vars = max(nc - num_variables, 0)
    if num_variables > nc:
        pad_rows = num_variables - nc
        mats = [CsrMatrix(np.concatenate([m.row_ptr, np.full(pad_rows, m.row_ptr[-1], dtype=np.uint32)]), m.col_idx, m.coeff)
                for m in mats]
    return mats, extra_vars


def balance_matrices(a, b):
    """constraint_systems.rs:100-114: while A is the denser matrix, swap row i of A and B.  `denser` is only
    re-evaluated after a swap, so nothing happens unless A starts out denser (the reference's behaviour)."""
    da, db = a.nnz, b.nnz
    if not da > db:
        return a, b
    la, lb = np.diff(a.row_ptr.astype(np.int64)), np.diff(b.row_ptr.astype(np.int64))
    swap = np.zeros(len(la), dtype=bool)
    denser = True
    for i in range(len(la)):
        if not denser:
            break
        swap[i] = True
        da += lb[i] - la[i]
        db += la[i] - lb[i]
        denser = da > db

    def merge(first, second, take_second):
        lens = np.where(take_second, np.diff(second.row_ptr.astype(np.int64)), np.diff(first.row_ptr.astype(np.int64)))
        ptr = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        cols = np.zeros(ptr[-1], dtype=np.uint32)
        coeff = np.zeros((ptr[-1], 4), dtype=np.uint64)
        for src, mask in ((first, ~take_second), (second, take_second)):
            sp = src.row_ptr.astype(np.int64)
            rows = np.flatnonzero(mask)
            cnt = (sp[rows + 1] - sp[rows])
            src_pos = np.repeat(sp[rows], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
            dst_pos = np.repeat(ptr[rows], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
            cols[dst_pos] = src.col_idx[src_pos]
            coeff[dst_pos] = src.coeff[src_pos]
        return CsrMatrix(ptr.astype(np.uint32), cols, coeff)

    return merge(a, b, swap), merge(b, a, swap)


def sort_rows_by_column(m):
    """process_matrices' per-row stable sort by variable index (constraint_systems.rs:92-96)"""
    rows = np.repeat(np.arange(m.n_rows, dtype=np.int64), np.diff(m.row_ptr.astype(np.int64)))
    order = np.lexsort((np.arange(len(rows)), m.col_idx.astype(np.int64), rows))
    return CsrMatrix(m.row_ptr, m.col_idx[order], m.coeff[order])


def compose_matrix_polynomials(ops, m, x_size, h_size, k_size, b_size, h_elems, inv_diag):
    """arithmetic.rs:97-172 with every field operation on the GPU: row / col / val / row_col evaluations on K,
    interpolated over K and evaluated over B.  h_elems = w_H^i, inv_diag[j] = 1 / u_H(w^j, w^j)."""
    nnz = m.nnz
    rows = np.repeat(np.arange(m.n_rows, dtype=np.int64), np.diff(m.row_ptr.astype(np.int64)))
    j = reindex_by_subdomain(h_size, x_size, m.col_idx.astype(np.int64))
    row_vec = np.empty((k_size, 4), dtype=np.uint64)
    col_vec = np.empty((k_size, 4), dtype=np.uint64)
    val_vec = np.zeros((k_size, 4), dtype=np.uint64)
    row_vec[:nnz] = h_elems[j]                   # note the reference's naming: "row" holds the column's element
    col_vec[:nnz] = h_elems[rows]
    row_vec[nnz:] = h_elems[0]
    col_vec[nnz:] = h_elems[0]
    if nnz:
        val_vec[:nnz] = ops.mul(np.ascontiguousarray(m.coeff), np.ascontiguousarray(inv_diag[j]))
    row_col_vec = ops.mul(row_vec, col_vec)
    out = {"row_evals_on_k": row_vec, "col_evals_on_k": col_vec, "val_evals_on_k": val_vec}
    for name, vec in (("row", row_vec), ("col", col_vec), ("val", val_vec), ("row_col", row_col_vec)):
        poly = trim(ops.ifft(vec, k_size))
        out[name] = poly
        out[name + "_evals_on_b"] = ops.fft(poly, b_size)
    return out


def index(ctx, curve, a, b, c, num_inputs, num_variables):
    """AHP::index (indexer.rs:71-116) for already synthesised matrices (CsrMatrix over the formatted variable
    numbering: inputs first, then witness).  Returns (Index, number of padding witness variables)."""
    (a, b, c), extra_vars = make_matrices_square([a, b, c], num_variables)
    num_variables += extra_vars
    a, b = balance_matrices(a, b)
    a, b, c = sort_rows_by_column(a), sort_rows_by_column(b), sort_rows_by_column(c)
    nnz = max(a.nnz, b.nnz, c.nnz)
    ops = Ops(ctx, curve)
    x_size, h_size, k_size = domain_size(num_inputs), domain_size(num_variables), domain_size(nnz)
    b_size = domain_size(3 * k_size - 3)
    for s_ in (x_size, h_size, k_size, b_size):
        if s_.bit_length() - 1 > FR_TWO_ADICITY[curve]:
            raise PolynomialDegreeTooLarge()
    h_elems = ops.elements(h_size)
    # u_H(w^j, w^j) = |H| * w^(-j) (arithmetic.rs:19-26 on the diagonal; :105-111 builds it reversed)
    inv_diag = ops.scale(h_elems, pow(h_size, -1, ops.f.p))
    stars = {name: compose_matrix_polynomials(ops, m, x_size, h_size, k_size, b_size, h_elems, inv_diag)
             for name, m in (("a", a), ("b", b), ("c", c))}
    return Index(curve, a.n_rows, num_variables, nnz, num_inputs, {"a": a, "b": b, "c": c}, stars), extra_vars


class ProverState:
    pass


def prover_init(ctx, index, formatted_input_mont, witness_mont, resident=False):
    """prover.rs:86-147 after synthesis and make_matrices_square: formatted_input = [one, inputs..],
    witness = aux assignment (+ padding variables).  resident=True keeps the whole round state in HBM
    (see Ops); the oracles returned by the rounds are then CUDA tensors."""
    ni, nw = len(formatted_input_mont), len(witness_mont)
    if index.num_constraints != index.matrices["a"].n_rows or index.num_constraints != ni + nw:
        raise InstanceDoesNotMatchIndex()
    st = ProverState()
    st.ctx, st.index, st.ops = ctx, index, Ops(ctx, index.curve, resident)
    o = st.ops
    if resident:
        index.make_resident(o)
    st.x, st.w = o.put(np.ascontiguousarray(formatted_input_mont)), o.put(np.ascontiguousarray(witness_mont))
    z = o.cat([st.x, st.w])
    st.z_a = ctx.spmv(index.curve, index.matrices["a"], z)
    st.z_b = ctx.spmv(index.curve, index.matrices["b"], z)
    st.zk_bound = 1
    return st


def prover_first_round(st, rng):
    """prover.rs:150-222.  rng.randrange(p) is called in the reference's order: the blinding constants of
    w, z_a, z_b (DensePolynomial::rand(zk_bound - 1)), then the 3|H| + 2 zk_bound - 2 mask coefficients."""
    o, idx = st.ops, st.index
    p = o.f.p
    H, X = idx.h_size, idx.x_size
    x_poly = trim(o.ifft(st.x, X))
    x_evals_on_h = o.fft(x_poly, H)
    ratio = H // X
    i = np.arange(H)
    if H > X:
        w_ext = pad(st.w, H - X)
        src = np.where(i % ratio == 0, 0, i - i // ratio - 1)
        gathered = o.take(w_ext, src)
    else:
        gathered = o.zeros(H)
    w_minus_x = o.zero_rows(o.sub(gathered, x_evals_on_h), i % ratio == 0)

    def blind(poly):                                                 # + DensePolynomial::rand(zk_bound - 1) * v_H
        c = rng.randrange(p)
        return o.add_const_terms(poly, [(0, -c), (H, c)])

    w_poly = blind(trim(o.ifft(w_minus_x, H)))
    w_poly, rem = o.divide_by_vanishing_poly(w_poly, X)
    assert not len(rem), "w is not divisible by v_X"                 # prover.rs:192
    z_a_poly = blind(trim(o.ifft(st.z_a, H)))
    z_b_poly = blind(trim(o.ifft(st.z_b, H)))
    mask_degree = 3 * H + 2 * st.zk_bound - 3
    if hasattr(rng, "field_array"):
        # bulk draw for large instances: rng.field_array(n) -> canonical uint64[n, 4] residues (the host RNG of
        # DensePolynomial::rand, prover.rs:202, without a Python-level loop)
        mask_canon = np.ascontiguousarray(rng.field_array(mask_degree + 1))
        heads = np.ascontiguousarray(mask_canon[0::H])
        sigma = sum(int.from_bytes(h.tobytes(), "little") for h in heads) % p
        mask_canon[0] = ints_to_limbs([(int.from_bytes(mask_canon[0].tobytes(), "little") - sigma) % p])[0]
        mask_poly = trim(st.ctx.fr_convert(idx.curve, o.put(mask_canon), to_mont=True))
    else:
        mask_ints = [rng.randrange(p) for _ in range(mask_degree + 1)]
        sigma = sum(mask_ints[k] for k in range(0, len(mask_ints), H)) % p     # remainder coefficient 0 mod (x^H - 1)
        mask_ints[0] = (mask_ints[0] - sigma) % p
        mask_poly = trim(st.ctx.fr_convert(idx.curve, o.put(ints_to_limbs(mask_ints)), to_mont=True))
    st.x_poly, st.w_poly, st.z_a_poly, st.z_b_poly, st.mask_poly = x_poly, w_poly, z_a_poly, z_b_poly, mask_poly
    # (label, polynomial, degree_bound, hiding_bound) as in ProverFirstOracles
    return [("w", w_poly, None, 1), ("z_a", z_a_poly, None, 1), ("z_b", z_b_poly, None, 1), ("mask", mask_poly, None, None)]


def prover_second_round(st, alpha, eta_a, eta_b, eta_c):
    """prover.rs:230-321; the verifier's first message as canonical ints"""
    o, idx = st.ops, st.index
    H, X = idx.h_size, idx.x_size
    za, zb = st.z_a_poly, st.z_b_poly
    m = o.scale(o.poly_mul(za, zb), eta_c)
    k = min(len(m), len(za), len(zb))
    if k:
        low = o.axpy(o.axpy(contig(m[:k]), eta_a, contig(za[:k])), eta_b, contig(zb[:k]))
        m = o.cat([low, m[k:]])
    m_poly = trim(m)
    r_alpha_evals = o.batch_evals(H, alpha)
    r_alpha_poly = trim(o.ifft(r_alpha_evals, H))
    t_evals = o.zeros(H)
    for name, eta in (("a", eta_a), ("b", eta_b), ("c", eta_c)):
        t_evals = o.axpy(t_evals, eta, st.ctx.spmv(idx.curve, idx.transposed[name], r_alpha_evals))
    t_poly = trim(o.ifft(t_evals, H))
    z_poly = o.mul_by_vanishing_poly(st.w_poly, X)
    k = min(len(z_poly), len(st.x_poly))
    if k:
        z_poly = o.cat([o.add(contig(z_poly[:k]), contig(st.x_poly[:k])), z_poly[k:]])
    size = domain_size(max(len(st.mask_poly), len(r_alpha_poly) + len(m_poly), len(t_poly) + len(z_poly)))
    ev = o.sub(o.mul(o.fft(r_alpha_poly, size), o.fft(m_poly, size)), o.mul(o.fft(t_poly, size), o.fft(z_poly, size)))
    q1 = o.poly_add(st.mask_poly, trim(o.ifft(ev, size)))
    h_1, x_g_1 = o.divide_by_vanishing_poly(q1, H)
    g_1 = trim(x_g_1[1:])
    st.t_poly, st.first_msg = t_poly, (alpha, eta_a, eta_b, eta_c)
    return [("t", t_poly, None, None), ("g_1", g_1, H - 2, st.zk_bound), ("h_1", h_1, None, None)]


def prover_third_round(st, beta):
    """prover.rs:331-427"""
    o, idx = st.ops, st.index
    p = o.f.p
    H, K, B = idx.h_size, idx.k_size, idx.b_size
    alpha, eta_a, eta_b, eta_c = st.first_msg
    vv = o.f.vanishing_at(H, alpha) * o.f.vanishing_at(H, beta) % p
    stars = [idx.stars[n] for n in "abc"]
    etas = [eta_a, eta_b, eta_c]
    t_k = o.zeros(K)
    for s, eta in zip(stars, etas):
        inv = o.inv(o.mul(o.rsub(beta, s["row_evals_on_k"]), o.rsub(alpha, s["col_evals_on_k"])))
        t_k = o.axpy(t_k, eta, o.mul(s["val_evals_on_k"], inv))
    t_poly = trim(o.ifft(o.scale(t_k, vv), K))
    g_2 = trim(t_poly[1:])
    # denom = beta * alpha - alpha * row - beta * col + row_col on B
    den = []
    for s in stars:
        d = o.axpy(s["row_col_evals_on_b"], -alpha % p, s["row_evals_on_b"])
        d = o.axpy(d, -beta % p, s["col_evals_on_b"])
        den.append(o.addc(d, alpha * beta % p))
    pairs = [(1, 2), (2, 0), (0, 1)]
    a_evals = o.zeros(B)
    for s, eta, (u, v) in zip(stars, etas, pairs):
        a_evals = o.axpy(a_evals, eta, o.mul(s["val_evals_on_b"], o.mul(den[u], den[v])))
    a_poly = trim(o.ifft(o.scale(a_evals, vv), B))
    b_poly = trim(o.ifft(o.mul(den[0], o.mul(den[1], den[2])), B))
    h_2 = o.divide_by_vanishing_poly(o.poly_sub(a_poly, o.poly_mul(b_poly, t_poly)), K)[0]
    return [("g_2", g_2, K - 2, None), ("h_2", h_2, None, None)]


# ------------------------------------------------------------------------------------------------
# The crate-level API: universal_setup / index / create_random_proof  (marlin/src/lib.rs:57-181)
# ------------------------------------------------------------------------------------------------
from . import fs_rng as _fs          # noqa: E402  (host-side Fiat-Shamir generator and ToBytes layouts)
from . import kzg10 as _kzg          # noqa: E402

INDEXER_POLYNOMIALS = ["a_row", "a_col", "a_val", "a_row_col", "b_row", "b_col", "b_val", "b_row_col",
                       "c_row", "c_col", "c_val", "c_row_col"]                      # ahp/mod.rs:34-50
PROVER_POLYNOMIALS = ["w", "z_a", "z_b", "mask", "t", "g_1", "h_1", "g_2", "h_2"]   # ahp/mod.rs:52-57


class IndexTooLarge(Exception):
    """Error::IndexTooLarge (lib.rs:72-74)"""


class UniversalParams:
    """pc::UniversalParams (pc/data_structures.rs:20-57): powers of beta in G1 (plain and gamma-scaled) and (h, beta h)
    in G2.  Point arrays in the layout of include/zkb.h."""

    def __init__(self, curve, powers_of_g, powers_of_gamma_g, h, beta_h):
        self.curve, self.powers_of_g, self.powers_of_gamma_g, self.h, self.beta_h = curve, powers_of_g, powers_of_gamma_g, h, beta_h

    def max_degree(self):
        return len(self.powers_of_g[1]) - 1


def universal_setup(ctx, curve, max_degree, rng):
    """lib.rs:57-65 -> PC::setup -> KZG10::setup (pc/kzg10.rs:27-72) with the fixed-base multiplications on the GPU
    (zkb_fixed_base_mul == FixedBaseMSM::multi_scalar_mul + batch_normalization_into_affine).  `rng.randrange` supplies
    beta and the discrete logs of g, gamma_g, h with respect to the standard generators (the reference draws random group
    elements through ark's `rand`, a byte stream only a Rust host reproduces; any generators give a valid SRS)."""
    from . import synth
    f = Field(curve)
    p = f.p
    max_degree = domain_size(max_degree)                       # compute_size_of_domain (lib.rs:61-62)
    beta = rng.randrange(1, p)
    kg, kgamma, kh = (rng.randrange(1, p) for _ in range(3))
    g1, g2 = synth.generator_mont(curve, _lib.G1), synth.generator_mont(curve, _lib.G2)

    def powers(scale):
        pw = ctx.fr_convert(curve, ctx.fr_powers(curve, f.mont(beta), max_degree + 1, scale_mont=f.mont(scale)), to_mont=False)
        xs, infs = [], []
        for i in range(0, len(pw), 1 << 18):
            xy, inf = ctx.fixed_base_mul(curve, _lib.G1, g1, np.ascontiguousarray(pw[i:i + (1 << 18)]))
            xs.append(xy)
            infs.append(inf)
        return np.concatenate(xs), np.concatenate(infs)

    hs, _ = ctx.fixed_base_mul(curve, _lib.G2, g2, ints_to_limbs([kh, kh * beta % p]))
    return UniversalParams(curve, powers(kg), powers(kgamma), (hs[0], False), (hs[1], False))


class VerifierKey:
    """pc::VerifierKey (pc/data_structures.rs:102-119)"""

    def __init__(self, curve, g, gamma_g, h, beta_h, supported_degree):
        self.curve, self.g, self.gamma_g, self.h, self.beta_h, self.supported_degree = curve, g, gamma_g, h, beta_h, supported_degree

    def to_bytes(self):
        return (b"".join(_fs.affine_to_bytes(self.curve, pt) for pt in (self.g, self.gamma_g, self.h, self.beta_h))
                + int(self.supported_degree).to_bytes(8, "little"))


def pc_trim(ctx, pp, supported_degree, shard=None):
    """PC::trim -> KZG10::trim (pc/kzg10.rs:74-98) -> (CommitterKey resident in HBM, VerifierKey).
    shard = (n_ranks, rank): powers_of_g sliced over the ranks (kzg10.CommitterKey)"""
    if supported_degree > pp.max_degree():
        raise _kzg.KzgError("TrimmingDegreeTooLarge")
    n = supported_degree + 1
    pg = (pp.powers_of_g[0][:n], pp.powers_of_g[1][:n])
    pgg = (pp.powers_of_gamma_g[0][:n], pp.powers_of_gamma_g[1][:n])
    ck = _kzg.CommitterKey(ctx, pp.curve, pg, pgg, supported_degree, shard=shard)
    vk = VerifierKey(pp.curve, (pg[0][0], bool(pg[1][0])), (pgg[0][0], bool(pgg[1][0])), pp.h, pp.beta_h, supported_degree)
    return ck, vk


class IndexVerifierKey:
    """data_structures.rs:10-33"""

    def __init__(self, curve, index_info, index_comms, verifier_key):
        self.curve, self.index_info, self.index_comms, self.verifier_key = curve, index_info, index_comms, verifier_key

    def to_bytes(self):
        """ToBytes (data_structures.rs:22-33): index_info, u32 count, the commitments, the verifier key"""
        return (_fs.index_info_to_bytes(*self.index_info) + len(self.index_comms).to_bytes(4, "little")
                + _fs.commitments_to_bytes(self.curve, self.index_comms) + self.verifier_key.to_bytes())


class IndexProverKey:
    """data_structures.rs:35-41"""

    def __init__(self, index, index_rands, index_verifier_key, committer_key, extra_vars=0):
        self.index, self.index_rands, self.index_verifier_key, self.committer_key = index, index_rands, index_verifier_key, committer_key
        self.extra_vars = extra_vars           # padding witness variables make_matrices_square appends (constraint_systems.rs:9-31)
        self.refresh_polys()

    def refresh_polys(self):
        """Index::iter (indexer.rs:51-67): the twelve labeled index polynomials, on whichever side the index lives"""
        self.index_polys = [_kzg.LabeledPolynomial(label, self.index.stars[label[0]][label[2:]], None, None)
                            for label in INDEXER_POLYNOMIALS]


class Proof:
    """data_structures.rs:43-48: commitments per round, the evaluations in query-set order, the opening proofs"""

    def __init__(self, commitments, evaluations, opening_proofs):
        self.commitments, self.evaluations, self.opening_proofs = commitments, evaluations, opening_proofs


def ahp_max_degree(num_constraints, num_variables, num_non_zeros):
    """AHP::max_degree (ahp/mod.rs:66-84)"""
    h, k = domain_size(max(num_constraints, num_variables)), domain_size(num_non_zeros)
    return max(3 * h + 2 * 1 - 1, 3 * k - 3)


def _synthesize(curve, circuit):
    """ConstraintSynthesizer -> (A, B, C CsrMatrix, formatted input ints or Montgomery array, witness, n_inputs).
    `circuit` either implements generate_constraints(cs) against the zkp_r1cs interface (r1cs.py), or -- for large
    synthetic instances built as arrays -- exposes marlin_arrays(ctx) -> (A, B, C, x_mont, w_mont)."""
    from .r1cs import ProvingAssignment
    pa = ProvingAssignment(FR_MODULUS[curve])
    pa.alloc_input(1)
    circuit.generate_constraints(pa)
    return pa


def _matrices_and_assignment(ctx, curve, circuit):
    if hasattr(circuit, "marlin_arrays"):
        return circuit.marlin_arrays(ctx)
    pa = _synthesize(curve, circuit)
    mats = []
    for which in "abc":
        ptr, cols, coeffs = pa.csr(which)
        mats.append(CsrMatrix(ptr, cols, ctx.fr_convert(curve, ints_to_limbs(coeffs), to_mont=True)))
    x = ctx.fr_convert(curve, ints_to_limbs(pa.input_assignment), to_mont=True)
    w = ctx.fr_convert(curve, ints_to_limbs(pa.aux_assignment), to_mont=True)
    return mats[0], mats[1], mats[2], x, w


def index_keys(ctx, srs, circuit, shard=None):
    """zkp_marlin::index (lib.rs:67-95): AHP::index, PC::trim, PC::commit of the twelve index polynomials (not hiding)
    -> (IndexProverKey, IndexVerifierKey).  shard = (n_ranks, rank): one process per GPU, the committer key sliced over
    the ranks; index_keys and create_random_proof are then collective calls with identical arguments on every rank."""
    curve = srs.curve
    a, b, c, x, w = _matrices_and_assignment(ctx, curve, circuit)
    idx, extra = index(ctx, curve, a, b, c, len(x), len(x) + len(w))
    max_degree = ahp_max_degree(idx.num_constraints, idx.num_variables, idx.num_non_zeros)
    if srs.max_degree() < max_degree:
        raise IndexTooLarge()
    ck, vk = pc_trim(ctx, srs, max_degree, shard=shard)
    ipk = IndexProverKey(idx, None, None, ck, extra)
    comms, rands = _kzg.pc_commit(ck, ipk.index_polys, None)
    ivk = IndexVerifierKey(curve, (idx.num_variables, idx.num_constraints, idx.num_non_zeros), comms, vk)
    ipk.index_rands, ipk.index_verifier_key = rands, ivk
    return ipk, ivk


def sample_element_outside_domain(f, size, rng):
    """ahp/verifier.rs:117-126"""
    t = rng.rand_fr()
    while f.vanishing_at(size, t) == 0:
        t = rng.rand_fr()
    return t


def verifier_query_set(beta, gamma):
    """ahp/verifier.rs:90-115 in BTreeSet order: sorted by label (every label occurs once)"""
    at_beta = ["w", "z_a", "z_b", "mask", "t", "g_1", "h_1"]
    return sorted([(l, beta) for l in at_beta] + [(l, gamma) for l in ["g_2", "h_2"] + INDEXER_POLYNOMIALS])


def create_random_proof(ctx, ipk, circuit, zk_rng, fs_rng=None, resident=True):
    """zkp_marlin::create_random_proof (lib.rs:97-181).  zk_rng: the prover's randomness (randrange(p); optional
    field_array(n) for bulk draws); fs_rng: the Fiat-Shamir generator -- fs_rng.FiatShamirRng (marlin/src/fs_rng.rs
    restated) seeded as lib.rs:105-106 when None, or any object with absorb(bytes) / rand_fr() / rand_u128()."""
    idx, ck, ivk = ipk.index, ipk.committer_key, ipk.index_verifier_key
    curve = idx.curve
    f = Field(curve)
    a, b, c, x, w = _matrices_and_assignment(ctx, curve, circuit)
    if ipk.extra_vars:                  # make_matrices_square's padding variables carry F::one() (constraint_systems.rs:15-19)
        w = np.concatenate([np.ascontiguousarray(w), np.tile(f.mont(1), (ipk.extra_vars, 1))])
    st = prover_init(ctx, idx, x, w, resident=resident)
    try:
        return _create_random_proof(ctx, ipk, st, x, zk_rng, fs_rng, resident)
    finally:
        st.ops.release()                # torch's current stream goes back to the caller


def _create_random_proof(ctx, ipk, st, x, zk_rng, fs_rng, resident):
    idx, ck, ivk = ipk.index, ipk.committer_key, ipk.index_verifier_key
    curve = idx.curve
    f = Field(curve)
    if resident:                        # the index polynomials moved into HBM with the rest of the index
        ipk.refresh_polys()
    public_input = np.ascontiguousarray(x)[1:]
    if fs_rng is None:
        fs_rng = _fs.FiatShamirRng(ivk.to_bytes() + _fs.fr_mont_array_to_bytes(ctx, curve, public_input), curve)
    labeled, comms_by_round, rands = list(ipk.index_polys), [], list(ipk.index_rands)

    def commit(round_polys):
        polys = [_kzg.LabeledPolynomial(label, poly, db, hb) for label, poly, db, hb in round_polys]
        cs, rs = _kzg.pc_commit(ck, polys, zk_rng)                      # lib.rs:109-110,117-118,124-125
        labeled.extend(polys)
        rands.extend(rs)
        comms_by_round.append(cs)
        fs_rng.absorb(_fs.commitments_to_bytes(curve, cs))             # lib.rs:112,120,127

    commit(prover_first_round(st, zk_rng))
    alpha = sample_element_outside_domain(f, idx.h_size, fs_rng)       # verifier_first_round (ahp/verifier.rs:40-70)
    eta_a, eta_b, eta_c = fs_rng.rand_fr(), fs_rng.rand_fr(), fs_rng.rand_fr()
    commit(prover_second_round(st, alpha, eta_a, eta_b, eta_c))
    beta = sample_element_outside_domain(f, idx.h_size, fs_rng)        # verifier_second_round (:72-80)
    commit(prover_third_round(st, beta))
    gamma = fs_rng.rand_fr()                                           # verifier_third_round (:82-88)

    query_set = verifier_query_set(beta, gamma)
    by_label = {P.label: P for P in labeled}
    points = {beta: f.mont(beta), gamma: f.mont(gamma)}
    evaluations = list(ctx.poly_eval_batch(curve, [by_label[label].coeffs for label, _ in query_set],
                                           np.stack([points[pt] for _, pt in query_set])))               # lib.rs:147-156
    fs_rng.absorb(_fs.fr_mont_array_to_bytes(ctx, curve, np.stack(evaluations)))                          # lib.rs:157
    opening_challenge = fs_rng.rand_u128()                             # u128::rand(&mut fs_rng).into() (lib.rs:158)
    opening_proofs = _kzg.pc_batch_open(ck, labeled, query_set, opening_challenge, rands)                 # lib.rs:160-166
    proof = Proof(comms_by_round, evaluations, opening_proofs)
    proof.challenges = {"alpha": alpha, "eta_a": eta_a, "eta_b": eta_b, "eta_c": eta_c, "beta": beta, "gamma": gamma,
                        "opening_challenge": opening_challenge}       # not part of the reference's Proof: kept for the tests
    return proof


# ------------------------------------------------------------------------------------------------
# verifier (lib.rs:183-260; ahp/verifier.rs): scalar work on the host as in the reference, the pairings of
# PC::batch_check on the GPU in one call
# ------------------------------------------------------------------------------------------------
class NonSquareMatrix(Exception):
    """ahp::Error::NonSquareMatrix (ahp/verifier.rs:44-46)"""


def _bivariate_eval(f, size, x, y):
    """arithmetic.rs:19-26: (v_H(x) - v_H(y)) / (x - y)"""
    p = f.p
    if x != y:
        return (f.vanishing_at(size, x) - f.vanishing_at(size, y)) * pow((x - y) % p, -1, p) % p
    return size * pow(x, size - 1, p) % p


def verifier_equality_check(curve, index_info, public_input, ev, alpha, eta_a, eta_b, eta_c, beta, gamma):
    """AHP::verifier_equality_check (ahp/verifier.rs:128-209).  public_input and the evaluations ev[(label, point)] are
    canonical ints."""
    f = Field(curve)
    p = f.p
    nv, nc, nnz = index_info
    h_size, k_size = domain_size(nc), domain_size(nnz)
    vha, vhb = f.vanishing_at(h_size, alpha), f.vanishing_at(h_size, beta)
    r_alpha_at_beta = _bivariate_eval(f, h_size, alpha, beta)
    formatted = [1] + [int(v) % p for v in public_input]
    x_size = domain_size(len(formatted))
    formatted += [0] * (x_size - len(formatted))
    vxb = f.vanishing_at(x_size, beta)
    # the interpolant of the formatted input over the input domain, at beta: sum_i x_i * v_X(beta) w^i / (|X| (beta - w^i))
    w, wi, x_at_beta = f.root_of_unity(x_size), 1, 0
    n_inv = pow(x_size, -1, p)
    for xi in formatted:
        if (beta - wi) % p == 0:
            x_at_beta = xi
            break
        x_at_beta = (x_at_beta + xi * vxb % p * wi % p * n_inv % p * pow((beta - wi) % p, -1, p)) % p
        wi = wi * w % p
    za, zb = ev[("z_a", beta)], ev[("z_b", beta)]
    lhs = (ev[("mask", beta)] + r_alpha_at_beta * (eta_a * za + eta_b * zb + eta_c * za * zb)
           - ev[("t", beta)] * (vxb * ev[("w", beta)] + x_at_beta)) % p
    if lhs != (ev[("h_1", beta)] * vhb + beta * ev[("g_1", beta)]) % p:                     # :163-166
        return False
    ab = alpha * beta % p
    den, val = [], []
    for name in "abc":
        e = {k: ev[("%s_%s" % (name, k), gamma)] for k in ("row", "col", "val", "row_col")}
        den.append((ab - alpha * e["row"] - beta * e["col"] + e["row_col"]) % p)
        val.append(e["val"])
    a_at = (eta_a * val[0] * den[1] * den[2] + eta_b * val[1] * den[2] * den[0] + eta_c * val[2] * den[0] * den[1]) % p
    a_at = a_at * vha % p * vhb % p
    b_at = den[0] * den[1] * den[2] % p
    lhs = ev[("h_2", gamma)] * f.vanishing_at(k_size, gamma) % p
    return lhs == (a_at - b_at * (gamma * ev[("g_2", gamma)] + ev[("t", beta)] * pow(k_size, -1, p))) % p   # :204-208


def verify_proof(ctx, ivk, proof, public_input, fs_rng=None):
    """zkp_marlin::verify_proof (lib.rs:183-260).  public_input: the instance without the leading one, as a Montgomery
    uint64[n, 4] array (what create_random_proof absorbs) or canonical ints."""
    curve = ivk.curve
    f = Field(curve)
    nv, nc, nnz = ivk.index_info
    if nc != nv:
        raise NonSquareMatrix()
    if isinstance(public_input, np.ndarray) and public_input.ndim == 2:
        x_mont = np.ascontiguousarray(public_input, dtype=np.uint64)
        x_int = [f.to_int(r) for r in x_mont]
    else:
        x_int = [int(v) % f.p for v in public_input]
        x_mont = f.mont_arr(x_int)
    if fs_rng is None:
        fs_rng = _fs.FiatShamirRng(ivk.to_bytes() + _fs.fr_mont_array_to_bytes(ctx, curve, x_mont), curve)
    first, second, third = proof.commitments
    h_size, k_size = domain_size(nc), domain_size(nnz)
    fs_rng.absorb(_fs.commitments_to_bytes(curve, first))
    alpha = sample_element_outside_domain(f, h_size, fs_rng)
    eta_a, eta_b, eta_c = fs_rng.rand_fr(), fs_rng.rand_fr(), fs_rng.rand_fr()
    fs_rng.absorb(_fs.commitments_to_bytes(curve, second))
    beta = sample_element_outside_domain(f, h_size, fs_rng)
    fs_rng.absorb(_fs.commitments_to_bytes(curve, third))
    gamma = fs_rng.rand_fr()
    query_set = verifier_query_set(beta, gamma)
    fs_rng.absorb(_fs.fr_mont_array_to_bytes(ctx, curve, np.stack([to_host(e) for e in proof.evaluations])))
    opening_challenge = fs_rng.rand_u128()
    labels = INDEXER_POLYNOMIALS + ["w", "z_a", "z_b", "mask", "t", "g_1", "h_1", "g_2", "h_2"]      # AHP::polynomial_labels
    bounds = dict.fromkeys(labels)
    bounds["g_1"], bounds["g_2"] = h_size - 2, k_size - 2
    comms = list(ivk.index_comms) + list(first) + list(second) + list(third)
    commitments = {l: (c, bounds[l]) for l, c in zip(labels, comms)}
    ev = {(l, pt): f.to_int(to_host(e)) for (l, pt), e in zip(query_set, proof.evaluations)}
    if not verifier_equality_check(curve, ivk.index_info, x_int, ev, alpha, eta_a, eta_b, eta_c, beta, gamma):
        return False
    return _kzg.pc_batch_check(ctx, ivk.verifier_key, commitments, query_set, ev, proof.opening_proofs, opening_challenge)
