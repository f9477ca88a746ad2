"""B200-native proving backend for ckb-zkp's Groth16 / Marlin prove path.

`ckb_zkp_b200.backend.Context` wraps the C ABI (include/zkb.h); `ckb_zkp_b200.groth16` mirrors
the reference's `zkp_groth16` prover API on top of it.  The CUDA extension (`libzkb.so`, built
by `__graft_entry__.build()`) is mandatory: there is no CPU code path in this package.
"""
from ._lib import BLS12_381, BN254, G1, G2  # noqa: F401
