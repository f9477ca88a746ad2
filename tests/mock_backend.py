"""A stand-in for ckb_zkp_b200.backend.Context whose primitives are computed by the Python oracle.
TEST INFRASTRUCTURE: lets the CPU suite exercise the host-side orchestration of ckb_zkp_b200/marlin.py
(which primitive is called with which operands) without a GPU.  It is never used by the product."""
import numpy as np

from oracle.pyref.fields import FR
from oracle.pyref.ntt import Domain
from tests import helpers as H


class MockContext:
    VEC_ADD, VEC_SUB, VEC_MUL, VEC_SCALE, VEC_AXPY, VEC_RSUB, VEC_ADDC = range(7)

    def _ints(self, curve, a):
        return H.fr_ints(curve, np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4))

    def fr_vec_op(self, curve, op, a, b=None, s=None):
        p = FR[curve].p
        x = self._ints(curve, a)
        y = self._ints(curve, b) if b is not None else [0] * len(x)
        sv = self._ints(curve, s)[0] if s is not None else 0
        f = [lambda u, v: u + v, lambda u, v: u - v, lambda u, v: u * v, lambda u, v: u * sv, lambda u, v: u + sv * v,
             lambda u, v: sv - u, lambda u, v: u + sv][op]
        return H.fr_array(curve, [f(u, v) % p for u, v in zip(x, y)])

    def fr_batch_inverse(self, curve, a):
        p = FR[curve].p
        return H.fr_array(curve, [pow(x, -1, p) if x else 0 for x in self._ints(curve, a)])

    def ntt(self, curve, data, log_n, inverse=False, coset=False):
        d = Domain(FR[curve], 1 << log_n)
        vals = self._ints(curve, data)
        fn = {(False, False): d.fft, (True, False): d.ifft, (False, True): d.coset_fft, (True, True): d.coset_ifft}[(inverse, coset)]
        data[:] = H.fr_array(curve, fn(vals))
        return data

    def fr_powers(self, curve, base_mont, n, scale_mont=None, device=None):
        p = FR[curve].p
        b = self._ints(curve, base_mont)[0]
        s = self._ints(curve, scale_mont)[0] if scale_mont is not None else 1
        out, cur = [], s
        for _ in range(n):
            out.append(cur)
            cur = cur * b % p
        return H.fr_array(curve, out)

    def spmv(self, curve, m, x_mont):
        p = FR[curve].p
        x = self._ints(curve, x_mont)
        co = self._ints(curve, m.coeff)
        y = []
        for i in range(m.n_rows):
            y.append(sum(co[k] * x[int(m.col_idx[k])] for k in range(int(m.row_ptr[i]), int(m.row_ptr[i + 1]))) % p)
        return H.fr_array(curve, y)

    def fr_convert(self, curve, a, to_mont):
        fr = FR[curve]
        vals = H.u64_to_ints(np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4))
        return H.ints_to_u64([fr.to_mont(v) if to_mont else fr.from_mont(v) for v in vals], 4)

    def fr_prefix_product(self, curve, a):
        p = FR[curve].p
        out, acc = [], 1
        for x in self._ints(curve, a):
            out.append(acc)
            acc = acc * x % p
        return H.fr_array(curve, out)

    def poly_eval(self, curve, p_mont, z_mont):
        p = FR[curve].p
        z = self._ints(curve, z_mont)[0]
        acc = 0
        for c in reversed(self._ints(curve, p_mont)):
            acc = (acc * z + c) % p
        return H.fr_array(curve, [acc])[0]

    def poly_lincomb(self, curve, polys, coeffs_mont, shifts=None, out_len=None):
        p = FR[curve].p
        cs = self._ints(curve, coeffs_mont)
        shifts = shifts or [0] * len(polys)
        n = out_len if out_len is not None else max([len(q) + s for q, s in zip(polys, shifts)] + [0])
        out = [0] * n
        for q, c, s in zip(polys, cs, shifts):
            for i, v in enumerate(self._ints(curve, q)):
                if i + s < n:
                    out[i + s] = (out[i + s] + c * v) % p
        return H.fr_array(curve, out) if n else np.zeros((0, 4), dtype=np.uint64)


class _MockSrs:
    def __init__(self, curve, group, pts):
        self.curve, self.group, self.pts, self.n = curve, group, pts, len(pts)

    def free(self):
        pass


class MockProverVerifierContext(MockContext):
    """MockContext plus the group side: base sets, MSMs (the oracle's double-and-add), fixed-base multiples, the remaining
    polynomial helpers, and pairings through the DEVICE code compiled for the host (tests/host_emu, csrc/pairing.cuh).
    Enough of backend.Context for whole proofs -- universal_setup, index_keys, create_random_proof, verify_proof of Marlin;
    keygen, prove, verify of PLONK -- to run on the CPU at toy sizes."""

    def __init__(self):
        from tests import emu
        self.lib_emu = emu.build()

    def sync(self):
        pass

    # ---- groups -------------------------------------------------------------------------------------------------
    def srs_upload(self, curve, group, xy, inf=None, precompute=True):
        inf = np.zeros(len(xy), dtype=np.uint8) if inf is None else inf
        return _MockSrs(curve, group, H.array_points(curve, group, np.asarray(xy), np.asarray(inf)))

    def msm(self, srs, scalars, base_offset=0, mont=False):
        from oracle.pyref.curves import CURVES
        from oracle.pyref.msm import msm_naive
        c = CURVES[(srs.curve, srs.group)]
        sc = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        ks = H.fr_ints(srs.curve, sc) if mont else H.u64_to_ints(sc)
        pt = c.to_affine(msm_naive(c, srs.pts[base_offset:], ks))
        xy, inf = H.points_array(srs.curve, srs.group, [pt])
        return xy[0], bool(inf[0])

    def msm_batch(self, srs_list, scalars_list, base_offsets=None, mont=False):
        offs = [0] * len(srs_list) if base_offsets is None else list(base_offsets)
        return [self.msm(s, sc, base_offset=o, mont=mont) for s, sc, o in zip(srs_list, scalars_list, offs)]

    def msm_many(self, srs, scalars, base_offsets=None, mont=False):
        k = len(scalars)
        offs = [0] * k if base_offsets is None else [int(o) for o in base_offsets]
        res = [self.msm(srs, scalars[i], base_offset=offs[i], mont=mont) for i in range(k)]
        w = len(res[0][0]) if res else 0
        return (np.stack([r[0] for r in res]) if res else np.zeros((0, w), dtype=np.uint64),
                np.array([1 if r[1] else 0 for r in res], dtype=np.uint8))

    def fixed_base_mul(self, curve, group, base_xy, scalars):
        from oracle.pyref.curves import CURVES
        c = CURVES[(curve, group)]
        base = H.array_point(curve, group, np.asarray(base_xy).reshape(-1), 0)
        ks = H.u64_to_ints(np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4))
        return H.points_array(curve, group, [c.mul_affine(base, k % c.r) if k % c.r else None for k in ks])

    def multi_pairing(self, curve, g1, g2, group_size):
        from oracle.pyref import pairing as OP
        from oracle.pyref.fields import FQ
        from tests.test_pairing_emu import flat_to_tower, run, tower_to_flat
        (xy1, inf1), (xy2, inf2) = g1, g2
        n = len(xy1)
        inf1 = np.zeros(n, dtype=np.uint8) if inf1 is None else inf1
        inf2 = np.zeros(n, dtype=np.uint8) if inf2 is None else inf2
        L, q = FQ[curve].limbs, FQ[curve].p
        F12 = OP.Fq12(curve)
        out = np.zeros((n // group_size, 12 * L), dtype=np.uint64)
        for g in range(n // group_size):
            f = F12.one
            for i in range(g * group_size, (g + 1) * group_size):
                P = H.array_point(curve, 1, xy1[i], inf1[i])
                Q = H.array_point(curve, 2, xy2[i], inf2[i])
                f = F12.mul(f, tower_to_flat(curve, run(self.lib_emu, curve, 0, P, Q)))
            gt = run(self.lib_emu, curve, 1, None, None, flat_to_tower(curve, f))
            out[g] = H.ints_to_u64([v * (1 << (64 * L)) % q for v in gt], L).reshape(-1)
        return out

    # ---- the remaining polynomial helpers ---------------------------------------------------------------------
    def poly_div_linear(self, curve, p_mont, z_mont, want_quotient=True):
        p = FR[curve].p
        z = self._ints(curve, z_mont)[0]
        coeffs = self._ints(curve, p_mont)
        q, acc = [0] * max(len(coeffs) - 1, 0), 0
        for i in reversed(range(len(coeffs))):                    # synthetic division by (x - z)
            acc = (acc * z + coeffs[i]) % p
            if i:
                q[i - 1] = acc
        quotient = (H.fr_array(curve, q) if q else np.zeros((0, 4), dtype=np.uint64)) if want_quotient else None
        return quotient, H.fr_array(curve, [acc])[0]

    def poly_eval_batch(self, curve, polys, points_mont):
        pts = np.ascontiguousarray(points_mont, dtype=np.uint64).reshape(-1, 4)
        if not len(polys):
            return np.zeros((0, 4), dtype=np.uint64)
        return np.stack([self.poly_eval(curve, q, pts[j]) if len(q) else np.zeros(4, dtype=np.uint64) for j, q in enumerate(polys)])
