"""ctypes binding of libzkb.so (C ABI: include/zkb.h).

The library holds device code only.  Importing this module never touches a GPU; creating a
`Context` does, and fails loudly (RuntimeError) when the extension is not built or no B200 is
visible -- there is no CPU fallback anywhere in this package.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libzkb.so")

BN254, BLS12_381 = 0, 1
G1, G2 = 1, 2
NTT_INVERSE, NTT_COSET = 1, 2
SRS_PRECOMPUTE = 1
COMM_ID_BYTES = 128
DECOMPRESS_CHECK_SUBGROUP = 1

OK, E_INVALID, E_CUDA, E_TOO_LARGE, E_NO_DEVICE = 0, -1, -2, -3, -4

c_void_p, c_int, c_uint, c_size_t, c_u64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t,
                                            ctypes.c_uint64)


class Csr(ctypes.Structure):
    """struct zkb_csr"""
    _fields_ = [("n_rows", c_size_t), ("nnz", c_size_t), ("row_ptr", c_void_p), ("col_idx", c_void_p),
                ("coeff_mont", c_void_p)]


# name -> (restype, argtypes); every symbol include/zkb.h declares
SIGNATURES = {
    "zkb_init": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "zkb_destroy": (None, [c_void_p]),
    "zkb_last_error": (ctypes.c_char_p, [c_void_p]),
    "zkb_stream": (c_void_p, [c_void_p]),
    "zkb_sync": (c_int, [c_void_p]),
    "zkb_launch_count": (c_u64, [c_void_p]),
    "zkb_set_serial": (c_int, [c_void_p, c_int]),
    "zkb_prof_enable": (c_int, [c_void_p, c_int]),
    "zkb_prof_read": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_u64),
                              ctypes.POINTER(ctypes.c_double)]),
    "zkb_srs_upload": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_uint, ctypes.POINTER(c_void_p)]),
    "zkb_srs_free": (None, [c_void_p]),
    "zkb_srs_len": (c_size_t, [c_void_p]),
    "zkb_msm": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zkb_msm_mont": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zkb_msm_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zkb_ntt": (c_int, [c_void_p, c_int, c_void_p, c_uint, c_uint]),
    "zkb_ntt_dev": (c_int, [c_void_p, c_int, c_void_p, c_uint, c_uint]),
    "zkb_groth16_h": (c_int, [c_void_p, c_int, ctypes.POINTER(Csr), ctypes.POINTER(Csr), ctypes.POINTER(Csr), c_void_p,
                              c_size_t, c_size_t, c_void_p]),
    "zkb_groth16_pk_create": (c_int, [c_void_p, c_int] + [c_void_p, c_void_p, c_size_t] * 5 + [c_void_p, c_void_p,
                                                                                           ctypes.POINTER(c_void_p)]),
    "zkb_groth16_pk_free": (None, [c_void_p]),
    "zkb_groth16_prove": (c_int, [c_void_p, c_void_p, ctypes.POINTER(Csr), ctypes.POINTER(Csr), ctypes.POINTER(Csr),
                                  c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zkb_groth16_stage": (c_int, [c_void_p, c_void_p, ctypes.POINTER(Csr), ctypes.POINTER(Csr), ctypes.POINTER(Csr),
                                  c_void_p, c_size_t, c_size_t]),
    "zkb_groth16_prove_staged": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "zkb_groth16_fetch_proof": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "zkb_comm_unique_id": (c_int, [c_void_p, c_void_p]),
    "zkb_comm_init": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "zkb_comm_destroy": (None, [c_void_p]),
    "zkb_comm_rank": (c_int, [c_void_p]),
    "zkb_comm_size": (c_int, [c_void_p]),
    "zkb_comm_collectives": (c_u64, [c_void_p]),
    "zkb_srs_upload_shard": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_uint,
                                     ctypes.POINTER(c_void_p)]),
    "zkb_msm_sharded": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "zkb_msm_sharded_local": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zkb_partial_bytes": (c_size_t, [c_int, c_int]),
    "zkb_msm_partial": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_void_p]),
    "zkb_msm_fold": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zkb_groth16_pk_create_sharded": (c_int, [c_void_p, c_int] + [c_void_p, c_void_p, c_size_t] * 5
                                      + [c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "zkb_groth16_prove_sharded": (c_int, [c_void_p, c_void_p, ctypes.POINTER(Csr), ctypes.POINTER(Csr), ctypes.POINTER(Csr),
                                          c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zkb_groth16_prove_sharded_staged": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "zkb_groth16_partial_bytes": (c_size_t, [c_int]),
    "zkb_groth16_prove_partial": (c_int, [c_void_p, c_void_p, ctypes.POINTER(Csr), ctypes.POINTER(Csr), ctypes.POINTER(Csr),
                                          c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "zkb_groth16_fold": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zkb_msm_batch": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "zkb_fixed_base_mul": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zkb_points_decompress": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_uint, c_void_p, c_void_p, c_void_p]),
    "zkb_multi_pairing": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "zkb_fr_convert": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_int]),
    "zkb_poly_div_linear": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "zkb_poly_eval_batch": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zkb_poly_lincomb": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "zkb_fr_prefix_product": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t]),
    "zkb_fr_batch_inverse": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t]),
    "zkb_fr_vec_op": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "zkb_fr_powers": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t]),
    "zkb_spmv": (c_int, [c_void_p, c_int, ctypes.POINTER(Csr), c_void_p, c_size_t, c_void_p]),
    "zkb_host_keccak_f1600": (None, [c_void_p]),
    "zkb_host_chacha20_blocks": (None, [c_void_p, c_u64, c_void_p, c_size_t]),
    "zkb_debug_fp_op": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t]),
    "zkb_debug_pt_op": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t]),
}

_lib = None


def load():
    """dlopen libzkb.so and bind every symbol; raises when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("ckb_zkp_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (make -C ckb_zkp_b200/csrc); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
