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

# the two Groth16 batch verifiers of ckb_zkp_b200/verifier.py through the Python API (proof objects in, decisions out):
# per-proof decisions (3 B Miller loops, B final exponentiations) vs one decision by random linear combination (B + 3, 1)
import random  # noqa: E402

from ckb_zkp_b200 import generator as zgen, groth16 as zg, verifier as zv  # noqa: E402
from ckb_zkp_b200.r1cs import ONE  # noqa: E402


class _Mini:
    """groth16/tests/mini.rs:12-44"""

    def generate_constraints(self, cs):
        vx, vy = cs.alloc(lambda: 2), cs.alloc(lambda: 3)
        vz = cs.alloc_input(lambda: 10)
        for _ in range(10):
            cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])


for curve, name in ((_lib.BN254, "bn254"), (_lib.BLS12_381, "bls12_381")):
    rng = random.Random(7)
    data = zgen.generate_random_parameters(ctx, curve, _Mini(), rng)
    params = data.upload(ctx)
    base = [zg.create_random_proof(params, _Mini(), rng) for _ in range(8)]
    params.free()
    pvk = zv.prepare_verifying_key(ctx, curve, data.vk)
    for B in (1024, 8192):
        proofs = [base[i % 8] for i in range(B)]
        inputs = [[10]] * B
        for fn, label in ((lambda: all(zv.verify_proofs(pvk, proofs, inputs)), "each"),
                          (lambda: zv.verify_proofs_batched(pvk, proofs, inputs, rng), "random_linear_combination")):
            assert fn()
            t = time.perf_counter()
            ok = fn()
            dt = time.perf_counter() - t
            print(json.dumps({"curve": name, "groth16_proofs": B, "mode": label, "accepted": bool(ok), "ms": round(dt * 1e3, 2),
                              "proofs_per_s": round(B / dt, 1)}), flush=True)
    pvk.free()
