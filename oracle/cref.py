"""ctypes wrapper of oracle/c/libzkref.so -- the C++ restatement of the reference's CPU prove path
(see the header of oracle/c/zkref.cpp).  TEST INFRASTRUCTURE / CPU BASELINE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "libzkref.so")
LIB_V3 = os.path.join(HERE, "c", "libzkref_v3.so")     # -march=x86-64-v3 build of the same source
SRC = os.path.join(HERE, "c", "zkref.cpp")

c_void_p, c_int, c_uint, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t


class Csr(ctypes.Structure):
    _fields_ = [("n_rows", c_size_t), ("nnz", c_size_t), ("row_ptr", c_void_p), ("col_idx", c_void_p),
                ("coeff_mont", c_void_p)]


class G16Key(ctypes.Structure):
    _fields_ = ([(k + "_xy", c_void_p) for k in ("a", "b1", "b2", "h", "l")]
                + [(k + "_inf", c_void_p) for k in ("a", "b1", "b2", "h", "l")]
                + [(k + "_len", c_size_t) for k in ("a", "b1", "b2", "h", "l")]
                + [("g1_singles", c_void_p), ("g2_singles", c_void_p)])


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        subprocess.check_call(["make", "-C", HERE])


_lib = None
_variant = None


def _host_is_v3():
    """x86-64-v3 needs (among others) bmi2 (mulx), avx2, fma, movbe; adx is what arkworks' `asm` feature also asks for"""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return {"bmi1", "bmi2", "avx2", "fma", "movbe", "adx", "abm"} <= flags
    except OSError:
        pass
    return False


def lib():
    global _lib, _variant
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        use_v3 = os.environ.get("ZKREF_GENERIC") != "1" and os.path.exists(LIB_V3) and _host_is_v3()
        _lib = ctypes.CDLL(LIB_V3 if use_v3 else LIB)
        _variant = "x86-64-v3 (mulx/adx/avx2)" if use_v3 else "generic x86-64"
        _lib.zkref_threads.restype = c_int
    return _lib


def variant():
    """which build of zkref.cpp is loaded (stated next to every CPU baseline number)"""
    lib()
    return _variant


def threads():
    return lib().zkref_threads()


def _p(a):
    return a.ctypes.data_as(c_void_p)


FQ_LIMBS = {0: 4, 1: 6}


def point_words(curve, group):
    return FQ_LIMBS[curve] * 2 * group


def msm(curve, group, bases_xy, inf, scalars, n_threads=None):
    """ark VariableBaseMSM::multi_scalar_mul -> (xy, is_identity, threads actually used)"""
    bases_xy = np.ascontiguousarray(bases_xy, dtype=np.uint64)
    inf = np.ascontiguousarray(inf, dtype=np.uint8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = min(len(inf), len(scalars))
    out = np.zeros(point_words(curve, group), dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    used = c_int(0)
    rc = lib().zkref_msm(c_int(curve), c_int(group), _p(bases_xy), _p(inf), _p(scalars), c_size_t(n),
                         c_int(n_threads or threads()), _p(out), _p(oinf), ctypes.byref(used))
    assert rc == 0
    return out, bool(oinf[0]), used.value


def fixed_base_mul(curve, group, base_xy, scalars, n_threads=None):
    base_xy = np.ascontiguousarray(base_xy, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = len(scalars)
    out = np.zeros((n, point_words(curve, group)), dtype=np.uint64)
    oinf = np.zeros(n, dtype=np.uint8)
    rc = lib().zkref_fixed_base_mul(c_int(curve), c_int(group), _p(base_xy), _p(scalars), c_size_t(n),
                                    c_int(n_threads or threads()), _p(out), _p(oinf))
    assert rc == 0
    return out, oinf


def ntt(curve, data, log_n, inverse=False, coset=False, n_threads=None):
    assert data.dtype == np.uint64 and data.flags.c_contiguous and data.shape == (1 << log_n, 4)
    rc = lib().zkref_ntt(c_int(curve), _p(data), c_uint(log_n), c_uint((1 if inverse else 0) | (2 if coset else 0)),
                         c_int(n_threads or threads()))
    if rc == -3:
        raise ValueError("PolynomialDegreeTooLarge")
    assert rc == 0
    return data


def fr_convert(curve, a, to_mont):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().zkref_fr_convert(c_int(curve), _p(a), _p(out), c_size_t(len(a)), c_int(1 if to_mont else 0))
    return out


def _csr(m):
    """m = (row_ptr uint32, col_idx uint32, coeff uint64[nnz,4]); returns (struct, keepalive)"""
    ptr = np.ascontiguousarray(m[0], dtype=np.uint32)
    col = np.ascontiguousarray(m[1], dtype=np.uint32)
    val = np.ascontiguousarray(m[2], dtype=np.uint64)
    return Csr(len(ptr) - 1, len(col), ptr.ctypes.data, col.ctypes.data, val.ctypes.data), (ptr, col, val)


def witness_map(curve, A, B, C, z_mont, n_inputs, n_threads=None):
    ca, ka = _csr(A)
    cb, kb = _csr(B)
    cc, kc = _csr(C)
    z = np.ascontiguousarray(z_mont, dtype=np.uint64)
    need = ca.n_rows + n_inputs
    log_n = max(need - 1, 0).bit_length()
    h = np.zeros((1 << log_n, 4), dtype=np.uint64)
    rc = lib().zkref_witness_map(c_int(curve), ctypes.byref(ca), ctypes.byref(cb), ctypes.byref(cc), _p(z),
                                 c_size_t(n_inputs), c_int(n_threads or threads()), _p(h))
    if rc == -3:
        raise ValueError("PolynomialDegreeTooLarge")
    assert rc == 0
    return h


def groth16_prove(curve, pk, A, B, C, z_mont, n_inputs, n_aux, r, s, n_threads=None):
    """pk: dict with a/b1/b2/h/l -> (xy, inf), g1_singles [alpha,beta,delta], g2_singles [beta,delta].
    Returns ((xy, inf) for A, B, C)."""
    keep = []
    key = G16Key()
    for k in ("a", "b1", "b2", "h", "l"):
        xy = np.ascontiguousarray(pk[k][0], dtype=np.uint64)
        inf = np.ascontiguousarray(pk[k][1], dtype=np.uint8)
        keep += [xy, inf]
        setattr(key, k + "_xy", xy.ctypes.data)
        setattr(key, k + "_inf", inf.ctypes.data)
        setattr(key, k + "_len", len(inf))
    s1 = np.ascontiguousarray(pk["g1_singles"], dtype=np.uint64)
    s2 = np.ascontiguousarray(pk["g2_singles"], dtype=np.uint64)
    key.g1_singles, key.g2_singles = s1.ctypes.data, s2.ctypes.data
    ca, ka = _csr(A)
    cb, kb = _csr(B)
    cc, kc = _csr(C)
    z = np.ascontiguousarray(z_mont, dtype=np.uint64)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    w1, w2 = point_words(curve, 1), point_words(curve, 2)
    out = np.zeros(2 * w1 + w2, dtype=np.uint64)
    oinf = np.zeros(3, dtype=np.uint8)
    rc = lib().zkref_groth16_prove(c_int(curve), ctypes.byref(key), ctypes.byref(ca), ctypes.byref(cb), ctypes.byref(cc),
                                   _p(z), c_size_t(n_inputs), c_size_t(n_aux), _p(r), _p(s),
                                   c_int(n_threads or threads()), _p(out), _p(oinf))
    if rc == -3:
        raise ValueError("PolynomialDegreeTooLarge")
    assert rc == 0
    return (out[:w1], bool(oinf[0])), (out[w1:w1 + w2], bool(oinf[1])), (out[w1 + w2:], bool(oinf[2]))
