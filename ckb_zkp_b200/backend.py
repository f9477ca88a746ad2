"""Thin object layer over the C ABI: one `Context` per GPU, numpy arrays in and out.

Array conventions (identical to the bytes ark-ff 0.2 keeps in memory, see include/zkb.h):
  Fr vector      uint64[n, 4]                      Montgomery unless a function says canonical
  G1 points      uint64[n, 2 * L]  (x || y)        L = 4 (BN254) or 6 (BLS12-381), Montgomery
  G2 points      uint64[n, 4 * L]  (x.c0 || x.c1 || y.c0 || y.c1)
  infinity       uint8[n]                          1 = identity (coordinates ignored)
"""
import ctypes

import numpy as np

try:                      # device-resident operands are torch tensors (PyTorch = device memory plumbing only)
    import torch
except Exception:         # pragma: no cover
    torch = None

from . import _lib
from ._lib import BLS12_381, BN254, G1, G2, Csr

FQ_LIMBS = {BN254: 4, BLS12_381: 6}


def point_words(curve, group):
    """u64 words of one affine point."""
    return FQ_LIMBS[curve] * 2 * (2 if group == G2 else 1)


class ZkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("zkb error %d: %s" % (code, msg))
        self.code = code


def _ptr(a):
    return ctypes.c_void_p(a.__array_interface__["data"][0]) if a is not None else None


def is_dev(a):
    """device-resident Fr vector: a CUDA torch tensor int64[n, 4] holding the same bytes as the uint64[n, 4] host form"""
    return torch is not None and isinstance(a, torch.Tensor)


def _fr(a, what="scalars"):
    if is_dev(a):
        if a.dtype != torch.int64 or a.dim() != 2 or a.shape[1] != 4 or not a.is_cuda:
            raise ValueError("%s must be a CUDA int64[n, 4] tensor" % what)
        return a.contiguous()
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("%s must be uint64[n, 4]" % what)
    return a


def _out_like(a, n):
    """uninitialised Fr vector of n elements on the same side as `a`"""
    if is_dev(a):
        return torch.empty((n, 4), dtype=torch.int64, device=a.device)
    return np.zeros((n, 4), dtype=np.uint64)


class CsrMatrix:
    """Row-major sparse matrix = flattened ProvingAssignment::{at,bt,ct} (groth16/src/prover.rs:16-25)."""

    def __init__(self, row_ptr, col_idx, coeff_mont):
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        self.col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
        self.coeff = np.ascontiguousarray(coeff_mont, dtype=np.uint64).reshape(-1, 4)
        if len(self.row_ptr) == 0:
            raise ValueError("row_ptr needs n_rows + 1 entries")
        if len(self.col_idx) != len(self.coeff) or int(self.row_ptr[-1]) != len(self.col_idx):
            raise ValueError("inconsistent CSR arrays")
        self.max_col = int(self.col_idx.max()) if len(self.col_idx) else -1     # checked against the vector length by spmv
        self.c = Csr(len(self.row_ptr) - 1, len(self.col_idx), self.row_ptr.ctypes.data, self.col_idx.ctypes.data,
                     self.coeff.ctypes.data)

    @property
    def n_rows(self):
        return len(self.row_ptr) - 1

    @property
    def nnz(self):
        return len(self.col_idx)

    def to_device(self, device):
        """resident copy for repeated zkb_spmv calls: same object, the C struct then points at device memory"""
        t = lambda a, dt: torch.from_numpy(a.view(dt)).to(device)
        self._dev = (t(self.row_ptr, np.int32), t(self.col_idx, np.int32), t(self.coeff, np.int64))
        self.c = Csr(len(self.row_ptr) - 1, len(self.col_idx), self._dev[0].data_ptr(), self._dev[1].data_ptr(),
                     self._dev[2].data_ptr())
        return self


class Srs:
    """Bases resident in HBM (the `&[G::Affine]` of VariableBaseMSM::multi_scalar_mul)."""

    def __init__(self, ctx, handle, curve, group, n):
        self.ctx, self.handle, self.curve, self.group, self.n = ctx, handle, curve, group, n

    def free(self):
        if self.handle:
            self.ctx.lib.zkb_srs_free(self.handle)
            self.handle = None

    def __len__(self):
        return self.n


class ProvingKey:
    """groth16 Parameters<E> (groth16/src/lib.rs:81-91) resident in HBM."""

    def __init__(self, ctx, handle, curve):
        self.ctx, self.handle, self.curve = ctx, handle, curve

    def free(self):
        if self.handle:
            self.ctx.lib.zkb_groth16_pk_free(self.handle)
            self.handle = None


class Context:
    def __init__(self, device=0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.zkb_init(device, ctypes.byref(h))
        if rc != 0:
            raise ZkbError(rc, "zkb_init(device=%d) failed: no usable B200 (sm_100) device -- this backend has no "
                               "CPU fallback" % device)
        self.handle = h
        self.device = device
        self._stream_ptr = self.lib.zkb_stream(h)
        self._lib_stream, self._order_cur = None, None

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, rc):
        cur, self._order_cur = self._order_cur, None
        if cur is not None:                      # the caller's torch stream continues after the library's work
            cur.wait_stream(self._lib_stream)
        if rc != 0:
            raise ZkbError(rc, self.lib.zkb_last_error(self.handle).decode())

    def _order_before(self):
        """A device buffer is about to be handed to the library: if the caller's current torch stream is not the library's
        stream (marlin.Ops makes it so for the resident prover), the library stream first waits for what the caller has
        enqueued, and _check makes the caller's stream wait for the library afterwards -- tensors produced or consumed
        by torch around a call need no manual synchronisation."""
        if self._order_cur is not None or torch is None:
            return
        dev = torch.device("cuda", self.device)          # the context's device, whatever torch's current device is
        cur = torch.cuda.current_stream(dev)
        if cur.cuda_stream == self._stream_ptr:
            return
        if self._lib_stream is None:
            self._lib_stream = torch.cuda.ExternalStream(self._stream_ptr, device=dev)
        self._lib_stream.wait_stream(cur)
        self._order_cur = cur

    def _addr(self, a):
        """host or device address: the library copies with cudaMemcpyDefault (unified addressing), so every buffer
        argument of the vector / polynomial / MSM entry points may live on either side"""
        if a is None:
            return None
        if is_dev(a):
            self._order_before()
            return ctypes.c_void_p(a.data_ptr())
        return _ptr(a)

    def close(self):
        if self.handle:
            self.lib.zkb_destroy(self.handle)
            self.handle = None

    def sync(self):
        self._check(self.lib.zkb_sync(self.handle))

    @property
    def stream(self):
        return self.lib.zkb_stream(self.handle)

    @property
    def launch_count(self):
        return int(self.lib.zkb_launch_count(self.handle))

    def set_serial(self, on):
        self._check(self.lib.zkb_set_serial(self.handle, 1 if on else 0))

    def prof_enable(self, on):
        self._check(self.lib.zkb_prof_enable(self.handle, 1 if on else 0))

    def prof_read(self):
        """{'ms', 'launches', 'alg_bytes'} of the bucket-accumulation kernel since prof_enable(True)"""
        ms, n, b = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_double()
        self._check(self.lib.zkb_prof_read(self.handle, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(b)))
        return {"ms": ms.value, "launches": int(n.value), "alg_bytes": b.value}

    # -- SRS / MSM ----------------------------------------------------------------------------
    def srs_upload(self, curve, group, xy, inf=None, precompute=True):
        xy = np.ascontiguousarray(xy, dtype=np.uint64)
        w = point_words(curve, group)
        if xy.ndim != 2 or xy.shape[1] != w:
            raise ValueError("bases must be uint64[n, %d]" % w)
        n = xy.shape[0]
        inf = np.zeros(n, dtype=np.uint8) if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
        if inf.shape != (n,):
            raise ValueError("infinity flags must be uint8[n]")
        h = ctypes.c_void_p()
        self._check(self.lib.zkb_srs_upload(self.handle, curve, group, _ptr(xy), _ptr(inf), n,
                                            _lib.SRS_PRECOMPUTE if precompute else 0, ctypes.byref(h)))
        return Srs(self, h, curve, group, n)

    def msm(self, srs, scalars, base_offset=0, mont=False):
        """sum scalars[i] * bases[base_offset + i]; returns (xy uint64[words], is_identity)."""
        scalars = _fr(scalars)
        out = np.zeros(point_words(srs.curve, srs.group), dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        fn = self.lib.zkb_msm_mont if mont else self.lib.zkb_msm
        self._check(fn(self.handle, srs.handle, base_offset, self._addr(scalars), scalars.shape[0], _ptr(out), _ptr(oinf)))
        return out, bool(oinf[0])

    def msm_batch(self, srs_list, scalars_list, base_offsets=None, mont=False):
        """k independent MSMs in one call (zkb_msm_batch): [(xy, is_identity)] in order.  scalars: host or device arrays"""
        k = len(srs_list)
        if k == 0:
            return []
        if len(scalars_list) != k:
            raise ValueError("one scalar array per SRS")
        base_offsets = [0] * k if base_offsets is None else list(base_offsets)
        sc = [_fr(s) for s in scalars_list]
        w = point_words(srs_list[0].curve, srs_list[0].group)
        handles = (ctypes.c_void_p * k)(*[s.handle for s in srs_list])
        offs = (ctypes.c_size_t * k)(*base_offsets)
        if any(is_dev(a) for a in sc):
            self._order_before()
        ptrs = (ctypes.c_void_p * k)(*[(a.data_ptr() if is_dev(a) else a.ctypes.data) if a.shape[0] else None for a in sc])
        lens = (ctypes.c_size_t * k)(*[a.shape[0] for a in sc])
        out = np.zeros((k, w), dtype=np.uint64)
        oinf = np.zeros(k, dtype=np.uint8)
        self._check(self.lib.zkb_msm_batch(self.handle, k, handles, offs, ptrs, lens, 1 if mont else 0, _ptr(out), _ptr(oinf)))
        return [(out[i], bool(oinf[i])) for i in range(k)]

    def msm_many(self, srs, scalars, base_offsets=None, mont=False):
        """k MSMs of n terms each over ONE base set (zkb_msm_batch): scalars uint64[k, n, 4] on the host, MSM i reads
        bases [base_offsets[i], base_offsets[i] + n) -> (xy uint64[k, words], inf uint8[k]).  The argument arrays are
        built without a Python-level loop: the shape of a batch verifier's calls (thousands of short MSMs)."""
        sc = np.ascontiguousarray(scalars, dtype=np.uint64)
        if sc.ndim != 3 or sc.shape[2] != 4:
            raise ValueError("scalars must be uint64[k, n, 4]")
        k, n = sc.shape[0], sc.shape[1]
        w = point_words(srs.curve, srs.group)
        out = np.zeros((k, w), dtype=np.uint64)
        oinf = np.zeros(k, dtype=np.uint8)
        if k == 0:
            return out, oinf
        handles = np.full(k, srs.handle.value, dtype=np.uint64)
        offs = np.zeros(k, dtype=np.uint64) if base_offsets is None else np.ascontiguousarray(base_offsets, dtype=np.uint64)
        ptrs = np.uint64(sc.ctypes.data) + np.arange(k, dtype=np.uint64) * np.uint64(n * 32)
        lens = np.full(k, n, dtype=np.uint64)
        self._check(self.lib.zkb_msm_batch(self.handle, k, _ptr(handles), _ptr(offs), _ptr(ptrs), _ptr(lens), 1 if mont else 0,
                                           _ptr(out), _ptr(oinf)))
        return out, oinf

    def msm_dev(self, srs, d_scalars_ptr, n, base_offset=0):
        """Same with canonical scalars already in device memory (raw device pointer)."""
        out = np.zeros(point_words(srs.curve, srs.group), dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        self._check(self.lib.zkb_msm_dev(self.handle, srs.handle, base_offset, ctypes.c_void_p(d_scalars_ptr), n,
                                         _ptr(out), _ptr(oinf)))
        return out, bool(oinf[0])

    def fixed_base_mul(self, curve, group, base_xy, scalars):
        scalars = _fr(scalars)
        base_xy = np.ascontiguousarray(base_xy, dtype=np.uint64).reshape(-1)
        w = point_words(curve, group)
        if base_xy.shape[0] != w:
            raise ValueError("base must be uint64[%d]" % w)
        n = scalars.shape[0]
        out = np.zeros((n, w), dtype=np.uint64)
        oinf = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.zkb_fixed_base_mul(self.handle, curve, group, _ptr(base_xy), _ptr(scalars), n, _ptr(out),
                                                _ptr(oinf)))
        return out, oinf

    # -- multi-GPU (one process per GPU; NCCL inside the library) -----------------------------------
    def comm_unique_id(self):
        """rank 0: the 128-byte NCCL id the host then distributes (torch.distributed broadcast, MPI, a file)"""
        buf = np.zeros(_lib.COMM_ID_BYTES, dtype=np.uint8)
        self._check(self.lib.zkb_comm_unique_id(self.handle, _ptr(buf)))
        return buf

    def comm_init(self, n_ranks, rank, unique_id=None):
        """collective: every rank calls it with the same id (None only for n_ranks == 1)"""
        uid = None if unique_id is None else np.ascontiguousarray(unique_id, dtype=np.uint8).reshape(_lib.COMM_ID_BYTES)
        self._check(self.lib.zkb_comm_init(self.handle, n_ranks, rank, _ptr(uid)))

    def comm_init_torch(self):
        """convenience for hosts that already run torch.distributed: broadcast the id over the default process
        group (the rendezvous only -- the data path is the library's own ncclAllGather)"""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            self.comm_init(1, 0)
            return 1, 0
        world, rank = dist.get_world_size(), dist.get_rank()
        on_gpu = dist.get_backend() == "nccl"
        t = torch.zeros(_lib.COMM_ID_BYTES, dtype=torch.uint8, device="cuda" if on_gpu else "cpu")
        if rank == 0:
            t.copy_(torch.from_numpy(self.comm_unique_id()))
        dist.broadcast(t, 0)
        self.comm_init(world, rank, t.cpu().numpy())
        return world, rank

    def comm_destroy(self):
        self.lib.zkb_comm_destroy(self.handle)

    @property
    def comm_size(self):
        return int(self.lib.zkb_comm_size(self.handle))

    @property
    def comm_rank(self):
        return int(self.lib.zkb_comm_rank(self.handle))

    @property
    def collective_count(self):
        return int(self.lib.zkb_comm_collectives(self.handle))

    def srs_upload_shard(self, curve, group, xy_local, inf_local, global_lo, global_n, precompute=True):
        xy = np.ascontiguousarray(xy_local, dtype=np.uint64)
        w = point_words(curve, group)
        if xy.ndim != 2 or xy.shape[1] != w:
            raise ValueError("bases must be uint64[n, %d]" % w)
        n = xy.shape[0]
        inf = np.zeros(n, dtype=np.uint8) if inf_local is None else np.ascontiguousarray(inf_local, dtype=np.uint8)
        if inf.shape != (n,):
            raise ValueError("infinity flags must be uint8[n]")
        h = ctypes.c_void_p()
        self._check(self.lib.zkb_srs_upload_shard(self.handle, curve, group, _ptr(xy), _ptr(inf), n, global_lo, global_n,
                                                  _lib.SRS_PRECOMPUTE if precompute else 0, ctypes.byref(h)))
        return Srs(self, h, curve, group, n)

    def msm_sharded(self, srs_shard, scalars, base_offset=0, mont=False):
        """collective; every rank passes the same full-length scalars and gets the same (xy, is_identity)"""
        scalars = _fr(scalars)
        out = np.zeros(point_words(srs_shard.curve, srs_shard.group), dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        self._check(self.lib.zkb_msm_sharded(self.handle, srs_shard.handle, base_offset, self._addr(scalars), scalars.shape[0],
                                             1 if mont else 0, _ptr(out), _ptr(oinf)))
        return out, bool(oinf[0])

    def msm_sharded_local(self, srs_shard, d_scalars_ptr, n_local):
        """collective; canonical scalars of this rank's pairs already in device memory"""
        out = np.zeros(point_words(srs_shard.curve, srs_shard.group), dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        self._check(self.lib.zkb_msm_sharded_local(self.handle, srs_shard.handle, ctypes.c_void_p(d_scalars_ptr), n_local,
                                                   _ptr(out), _ptr(oinf)))
        return out, bool(oinf[0])

    def msm_partial(self, srs_shard, scalars, base_offset=0, mont=False):
        """this rank's partial point as opaque bytes (for a caller-supplied transport)"""
        scalars = _fr(scalars)
        out = np.zeros(int(self.lib.zkb_partial_bytes(srs_shard.curve, srs_shard.group)), dtype=np.uint8)
        self._check(self.lib.zkb_msm_partial(self.handle, srs_shard.handle, base_offset, self._addr(scalars), scalars.shape[0],
                                             1 if mont else 0, _ptr(out)))
        return out

    def msm_fold(self, curve, group, partials):
        """fold of the ranks' partials (uint8[count, partial_bytes]) in index order -> (xy, is_identity)"""
        partials = np.ascontiguousarray(partials, dtype=np.uint8)
        out = np.zeros(point_words(curve, group), dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        self._check(self.lib.zkb_msm_fold(self.handle, curve, group, _ptr(partials), partials.shape[0], _ptr(out), _ptr(oinf)))
        return out, bool(oinf[0])

    def points_decompress(self, curve, group, compressed, check_subgroup=False):
        """ark-serialize compressed points (uint8[n, bytes]) -> (xy uint64[n, words], inf uint8[n], status uint8[n])"""
        w = point_words(curve, group)
        data = np.ascontiguousarray(compressed, dtype=np.uint8).reshape(-1, w * 4)
        n = data.shape[0]
        xy = np.zeros((n, w), dtype=np.uint64)
        inf = np.zeros(n, dtype=np.uint8)
        status = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.zkb_points_decompress(self.handle, curve, group, _ptr(data), n,
                                                   _lib.DECOMPRESS_CHECK_SUBGROUP if check_subgroup else 0, _ptr(xy), _ptr(inf),
                                                   _ptr(status)))
        return xy, inf, status

    def multi_pairing(self, curve, g1, g2, group_size):
        """products of pairings (zkb_multi_pairing): g1 = (xy uint64[n, words], inf uint8[n] or None), g2 likewise,
        n = n_groups * group_size -> uint64[n_groups, 12 * limbs] (Fq12, Montgomery, ark-ff tower order)"""
        (xy1, inf1), (xy2, inf2) = g1, g2
        w1, w2 = point_words(curve, 1), point_words(curve, 2)
        xy1 = np.ascontiguousarray(xy1, dtype=np.uint64).reshape(-1, w1)
        xy2 = np.ascontiguousarray(xy2, dtype=np.uint64).reshape(-1, w2)
        n = xy1.shape[0]
        if xy2.shape[0] != n or group_size < 1 or n % group_size:
            raise ValueError("multi_pairing: %d G1 points, %d G2 points, groups of %d" % (n, xy2.shape[0], group_size))
        inf1 = None if inf1 is None else np.ascontiguousarray(inf1, dtype=np.uint8)
        inf2 = None if inf2 is None else np.ascontiguousarray(inf2, dtype=np.uint8)
        if (inf1 is not None and inf1.shape != (n,)) or (inf2 is not None and inf2.shape != (n,)):
            raise ValueError("multi_pairing: one identity flag per point")
        out = np.zeros((n // group_size, 6 * w1), dtype=np.uint64)
        self._check(self.lib.zkb_multi_pairing(self.handle, curve, _ptr(xy1), None if inf1 is None else _ptr(inf1), _ptr(xy2),
                                               None if inf2 is None else _ptr(inf2), n // group_size, group_size, _ptr(out)))
        return out

    # -- NTT ----------------------------------------------------------------------------------
    def ntt(self, curve, data, log_n, inverse=False, coset=False):
        """In-place transform of uint64[2^log_n, 4] (Montgomery), natural order in and out."""
        flags = (_lib.NTT_INVERSE if inverse else 0) | (_lib.NTT_COSET if coset else 0)
        if is_dev(data):
            if not (data.dtype == torch.int64 and data.is_contiguous() and tuple(data.shape) == (1 << log_n, 4)):
                raise ValueError("data must be a contiguous CUDA int64[2^log_n, 4] tensor")
            self._order_before()
            self._check(self.lib.zkb_ntt_dev(self.handle, curve, ctypes.c_void_p(data.data_ptr()), log_n, flags))
            return data
        if not (isinstance(data, np.ndarray) and data.dtype == np.uint64 and data.flags.c_contiguous
                and data.shape == (1 << log_n, 4)):
            raise ValueError("data must be a contiguous uint64[2^log_n, 4] array")
        self._check(self.lib.zkb_ntt(self.handle, curve, _ptr(data), log_n, flags))
        return data

    def ntt_dev(self, curve, d_ptr, log_n, inverse=False, coset=False):
        flags = (_lib.NTT_INVERSE if inverse else 0) | (_lib.NTT_COSET if coset else 0)
        self._check(self.lib.zkb_ntt_dev(self.handle, curve, ctypes.c_void_p(d_ptr), log_n, flags))

    def fr_convert(self, curve, a, to_mont):
        a = _fr(a, "elements")
        out = _out_like(a, a.shape[0])
        self._check(self.lib.zkb_fr_convert(self.handle, curve, self._addr(a), self._addr(out), a.shape[0], 1 if to_mont else 0))
        return out

    # -- polynomial helpers (Marlin) ---------------------------------------------------------------
    def poly_div_linear(self, curve, p_mont, z_mont, want_quotient=True):
        """(q, rem): q = p / (x - z) (n - 1 coefficients) and rem = p(z); all Montgomery uint64[., 4]"""
        p = _fr(p_mont, "polynomial")
        z = np.ascontiguousarray(z_mont, dtype=np.uint64).reshape(4)
        n = p.shape[0]
        q = _out_like(p, max(n - 1, 0)) if want_quotient else None
        rem = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.zkb_poly_div_linear(self.handle, curve, self._addr(p), n, _ptr(z),
                                                 self._addr(q) if (want_quotient and n > 1) else None, _ptr(rem)))
        return q, rem

    def poly_eval(self, curve, p_mont, z_mont):
        return self.poly_div_linear(curve, p_mont, z_mont, want_quotient=False)[1]

    def poly_eval_batch(self, curve, polys, points_mont):
        """[polys[j](points[j])] as uint64[k, 4] (Montgomery) with one device synchronisation (zkb_poly_eval_batch)"""
        polys = [_fr(p, "polynomial") for p in polys]
        k = len(polys)
        pts = np.ascontiguousarray(points_mont, dtype=np.uint64).reshape(-1, 4)
        if pts.shape[0] != k:
            raise ValueError("one point per polynomial")
        out = np.zeros((k, 4), dtype=np.uint64)
        if k == 0:
            return out
        if any(is_dev(p) for p in polys):
            self._order_before()
        ptrs = (ctypes.c_void_p * k)(*[(p.data_ptr() if is_dev(p) else p.ctypes.data) if len(p) else None for p in polys])
        lens = (ctypes.c_size_t * k)(*[len(p) for p in polys])
        self._check(self.lib.zkb_poly_eval_batch(self.handle, curve, k, ptrs, lens, _ptr(pts), _ptr(out)))
        return out

    def poly_lincomb(self, curve, polys, coeffs_mont, shifts=None, out_len=None):
        """sum_j coeffs[j] * x^shifts[j] * polys[j] as uint64[out_len, 4]"""
        polys = [_fr(p, "polynomial") for p in polys]
        k = len(polys)
        shifts = [0] * k if shifts is None else list(shifts)
        if out_len is None:
            out_len = max([len(p) + s for p, s in zip(polys, shifts)] + [0])
        coeffs = _fr(coeffs_mont, "coefficients")
        if coeffs.shape[0] != k:
            raise ValueError("one coefficient per polynomial")
        if any(is_dev(p) for p in polys):
            self._order_before()
        ptrs = (ctypes.c_void_p * max(k, 1))(*[p.data_ptr() if is_dev(p) else p.ctypes.data for p in polys])
        lens = (ctypes.c_size_t * max(k, 1))(*[len(p) for p in polys])
        shs = (ctypes.c_size_t * max(k, 1))(*shifts)
        dev = [p for p in polys if is_dev(p)]
        out = _out_like(dev[0], out_len) if dev else np.zeros((out_len, 4), dtype=np.uint64)
        self._check(self.lib.zkb_poly_lincomb(self.handle, curve, k, ptrs, lens, shs, _ptr(coeffs), self._addr(out), out_len))
        return out

    def fr_prefix_product(self, curve, a_mont):
        """out[i] = a[0] * ... * a[i - 1] (out[0] = 1)"""
        a = _fr(a_mont, "elements")
        out = _out_like(a, a.shape[0])
        self._check(self.lib.zkb_fr_prefix_product(self.handle, curve, self._addr(a), self._addr(out), a.shape[0]))
        return out

    def fr_batch_inverse(self, curve, a_mont):
        a = _fr(a_mont, "elements")
        out = _out_like(a, a.shape[0])
        self._check(self.lib.zkb_fr_batch_inverse(self.handle, curve, self._addr(a), self._addr(out), a.shape[0]))
        return out

    VEC_ADD, VEC_SUB, VEC_MUL, VEC_SCALE, VEC_AXPY, VEC_RSUB, VEC_ADDC = range(7)

    def fr_vec_op(self, curve, op, a, b=None, s=None):
        """elementwise op on uint64[n, 4] Montgomery arrays (see zkb_fr_vec_op); s = uint64[4] scalar"""
        a = _fr(a, "a")
        b = None if b is None else _fr(b, "b")
        if b is not None and b.shape != a.shape:
            raise ValueError("operand shapes differ")
        s = None if s is None else np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        out = _out_like(a, a.shape[0])
        self._check(self.lib.zkb_fr_vec_op(self.handle, curve, op, self._addr(a), self._addr(b), _ptr(s), self._addr(out), a.shape[0]))
        return out

    def fr_powers(self, curve, base_mont, n, scale_mont=None, device=None):
        base = np.ascontiguousarray(base_mont, dtype=np.uint64).reshape(4)
        sc = None if scale_mont is None else np.ascontiguousarray(scale_mont, dtype=np.uint64).reshape(4)
        out = np.zeros((n, 4), dtype=np.uint64) if device is None else torch.empty((n, 4), dtype=torch.int64, device=device)
        self._check(self.lib.zkb_fr_powers(self.handle, curve, _ptr(base), _ptr(sc), self._addr(out), n))
        return out

    def spmv(self, curve, m, x_mont):
        x = _fr(x_mont, "x")
        if m.max_col >= x.shape[0]:          # the kernel reads x[col] unchecked: a malformed matrix must not reach the device
            raise ValueError("CSR column index %d out of range for a vector of %d elements" % (m.max_col, x.shape[0]))
        y = _out_like(x, m.n_rows)
        self._check(self.lib.zkb_spmv(self.handle, curve, ctypes.byref(m.c), self._addr(x), x.shape[0], self._addr(y)))
        return y

    # -- Groth16 ------------------------------------------------------------------------------
    def groth16_h(self, curve, A, B, C, z_mont, n_inputs, n_aux):
        z = _fr(z_mont, "assignment")
        if z.shape[0] != n_inputs + n_aux:
            raise ValueError("assignment length != n_inputs + n_aux")
        need = A.n_rows + n_inputs
        log_n = max(need - 1, 0).bit_length()
        h = np.zeros((1 << log_n, 4), dtype=np.uint64)
        self._check(self.lib.zkb_groth16_h(self.handle, curve, ctypes.byref(A.c), ctypes.byref(B.c), ctypes.byref(C.c),
                                           _ptr(z), n_inputs, n_aux, _ptr(h)))
        return h

    def groth16_pk(self, curve, a, b_g1, b_g2, h, l, g1_singles, g2_singles, shard=None):
        """a, b_g1, b_g2, h, l: (xy, inf) pairs; g1_singles = [alpha, beta, delta], g2_singles = [beta, delta].
        shard = (n_ranks, rank): keep only this rank's slice of the MSM pairs resident (zkb_groth16_pk_create_sharded)."""
        args = []
        for (xy, inf), group in ((a, G1), (b_g1, G1), (b_g2, G2), (h, G1), (l, G1)):
            xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, point_words(curve, group))
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
            if inf.shape != (xy.shape[0],):
                raise ValueError("infinity flags / points mismatch")
            args.append((xy, inf))
        s1 = np.ascontiguousarray(g1_singles, dtype=np.uint64).reshape(3, point_words(curve, G1))
        s2 = np.ascontiguousarray(g2_singles, dtype=np.uint64).reshape(2, point_words(curve, G2))
        flat = []
        for xy, inf in args:
            flat += [_ptr(xy), _ptr(inf), xy.shape[0]]
        hdl = ctypes.c_void_p()
        if shard is None:
            self._check(self.lib.zkb_groth16_pk_create(self.handle, curve, *flat, _ptr(s1), _ptr(s2), ctypes.byref(hdl)))
        else:
            self._check(self.lib.zkb_groth16_pk_create_sharded(self.handle, curve, *flat, _ptr(s1), _ptr(s2), int(shard[0]),
                                                               int(shard[1]), ctypes.byref(hdl)))
        return ProvingKey(self, hdl, curve)

    def _proof_arrays(self, curve):
        w1, w2 = point_words(curve, G1), point_words(curve, G2)
        return np.zeros(2 * w1 + w2, dtype=np.uint64), np.zeros(3, dtype=np.uint8), w1, w2

    @staticmethod
    def _split_proof(buf, inf, w1, w2):
        return (buf[:w1].copy(), bool(inf[0])), (buf[w1:w1 + w2].copy(), bool(inf[1])), (buf[w1 + w2:].copy(), bool(inf[2]))

    def groth16_prove(self, pk, A, B, C, z_mont, n_inputs, n_aux, r, s):
        """zkb_groth16_prove: host buffers in, proof (A, B, C) out as ((xy, is_identity), ...)."""
        z = _fr(z_mont, "assignment")
        if z.shape[0] != n_inputs + n_aux:
            raise ValueError("assignment length != n_inputs + n_aux")
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        s = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        buf, inf, w1, w2 = self._proof_arrays(pk.curve)
        self._check(self.lib.zkb_groth16_prove(self.handle, pk.handle, ctypes.byref(A.c), ctypes.byref(B.c),
                                               ctypes.byref(C.c), _ptr(z), n_inputs, n_aux, _ptr(r), _ptr(s), _ptr(buf),
                                               _ptr(inf)))
        return self._split_proof(buf, inf, w1, w2)

    def groth16_stage(self, pk, A, B, C, z_mont, n_inputs, n_aux):
        z = _fr(z_mont, "assignment")
        self._check(self.lib.zkb_groth16_stage(self.handle, pk.handle, ctypes.byref(A.c), ctypes.byref(B.c),
                                               ctypes.byref(C.c), _ptr(z), n_inputs, n_aux))

    def groth16_prove_staged(self, pk, r, s):
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        s = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        self._check(self.lib.zkb_groth16_prove_staged(self.handle, pk.handle, _ptr(r), _ptr(s)))

    def groth16_prove_sharded(self, pk, A, B, C, z_mont, n_inputs, n_aux, r, s):
        """collective: ONE proof by all ranks (same arguments and result on every rank as groth16_prove)"""
        z = _fr(z_mont, "assignment")
        if z.shape[0] != n_inputs + n_aux:
            raise ValueError("assignment length != n_inputs + n_aux")
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        s = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        buf, inf, w1, w2 = self._proof_arrays(pk.curve)
        self._check(self.lib.zkb_groth16_prove_sharded(self.handle, pk.handle, ctypes.byref(A.c), ctypes.byref(B.c),
                                                       ctypes.byref(C.c), _ptr(z), n_inputs, n_aux, _ptr(r), _ptr(s),
                                                       _ptr(buf), _ptr(inf)))
        return self._split_proof(buf, inf, w1, w2)

    def groth16_prove_sharded_staged(self, pk, r, s):
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        s = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        self._check(self.lib.zkb_groth16_prove_sharded_staged(self.handle, pk.handle, _ptr(r), _ptr(s)))

    def groth16_prove_partial(self, pk, A, B, C, z_mont, n_inputs, n_aux, r, s):
        """this rank's (A_k, C_k, B2_k) as opaque bytes (caller-supplied transport / single-GPU tests)"""
        z = _fr(z_mont, "assignment")
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        s = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        out = np.zeros(int(self.lib.zkb_groth16_partial_bytes(pk.curve)), dtype=np.uint8)
        self._check(self.lib.zkb_groth16_prove_partial(self.handle, pk.handle, ctypes.byref(A.c), ctypes.byref(B.c),
                                                       ctypes.byref(C.c), _ptr(z), n_inputs, n_aux, _ptr(r), _ptr(s),
                                                       _ptr(out)))
        return out

    def groth16_fold(self, pk, partials, r, s):
        partials = np.ascontiguousarray(partials, dtype=np.uint8)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        s = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        buf, inf, w1, w2 = self._proof_arrays(pk.curve)
        self._check(self.lib.zkb_groth16_fold(self.handle, pk.handle, _ptr(partials), partials.shape[0], _ptr(r), _ptr(s),
                                              _ptr(buf), _ptr(inf)))
        return self._split_proof(buf, inf, w1, w2)

    def groth16_fetch_proof(self, pk):
        buf, inf, w1, w2 = self._proof_arrays(pk.curve)
        self._check(self.lib.zkb_groth16_fetch_proof(self.handle, pk.handle, _ptr(buf), _ptr(inf)))
        return self._split_proof(buf, inf, w1, w2)

    # -- diagnostics --------------------------------------------------------------------------
    def debug_fp_op(self, field, op, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        out = np.zeros_like(a)
        self._check(self.lib.zkb_debug_fp_op(self.handle, field, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out

    def debug_pt_op(self, curve, group, op, acc, q, neg, out_words):
        acc = np.ascontiguousarray(acc, dtype=np.uint32)
        q = None if q is None else np.ascontiguousarray(q, dtype=np.uint32)
        out = np.zeros((acc.shape[0], out_words), dtype=np.uint32)
        self._check(self.lib.zkb_debug_pt_op(self.handle, curve, group, op, _ptr(acc), _ptr(q), 1 if neg else 0,
                                             _ptr(out), acc.shape[0]))
        return out
