#!/usr/bin/env python3
"""Throughput of the batched pairing check (zkb_multi_pairing): B groups of 3 pairs -- the shape of B Groth16
verifications (groth16/src/verifier.rs:31-41) -- per curve, host buffers in, GT elements out (copies inside the timed
region).  One JSON line per (curve, B)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from ckb_zkp_b200 import _lib, synth  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402
from ckb_zkp_b200.r1cs import ints_to_limbs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, nargs="+", default=[1, 64, 1024, 8192, 32768])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
ctx = Context(0)
rng = np.random.default_rng(3)
for curve, name in ((_lib.BN254, "bn254"), (_lib.BLS12_381, "bls12_381")):
    r = synth.FR_MODULUS[curve]
    g1, g2 = synth.generator_mont(curve, _lib.G1), synth.generator_mont(curve, _lib.G2)
    for B in a.batches:
        n = 3 * B
        ks = ints_to_limbs([int.from_bytes(rng.bytes(31), "little") % r for _ in range(2 * n)])
        P = ctx.fixed_base_mul(curve, _lib.G1, g1, ks[:n])
        Q = ctx.fixed_base_mul(curve, _lib.G2, g2, ks[n:])
        ctx.multi_pairing(curve, P, Q, 3)                       # warm-up (module load, allocator)
        best = None
        for _ in range(a.reps):
            t = time.perf_counter()
            out = ctx.multi_pairing(curve, P, Q, 3)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        print(json.dumps({"curve": name, "groups": B, "pairs": n, "ms": round(best * 1e3, 3),
                          "verifications_per_s": round(B / best, 1), "pairings_per_s": round(n / best, 1)}), flush=True)
