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
