#!/usr/bin/env python3
"""PLONK prove on one B200: BLS12-381, a chain of 2^log_n - 3 multiplication / addition gates,
keygen once, then `--steps` proofs through ckb_zkp_b200.plonk.prove (index and round state resident in HBM by default,
`--host-buffers` for numpy arrays between primitives).  Wall clock per proof with a device sync; the conversion of the
Composer's Python-int assignment to limbs happens once per circuit state (the reference's Composer holds field elements
from alloc_and_assign on) and is reported separately.  Every proof is self-checked: the verifier's equality check (ahp/verifier.rs:105-150)
under the transcript's challenges.  Prints one JSON line."""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from ckb_zkp_b200 import _lib, marlin as zm, plonk as zp  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=16)
ap.add_argument("--curve", type=int, default=_lib.BLS12_381)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--host-buffers", action="store_true")
a = ap.parse_args()

ctx = Context(0)
curve = a.curve
p = zp.FR_MODULUS[curve]
n = (1 << a.log_n) - 3
rng = random.Random(1)
t0 = time.perf_counter()
cs = zp.Composer(p)
x_v = rng.randrange(p)
x = cs.alloc_and_assign(x_v)
for i in range(n):
    c = rng.randrange(p)
    if i % 2 == 0:                               # y = x * x + c
        y_v = (x_v * x_v + c) % p
        y = cs.alloc_and_assign(y_v)
        cs.create_mul_gate(x, x, y, None, 1, c, 0)
    else:                                        # y = 3 x + c (the public input carries c on every 64th gate)
        pi = c if i % 64 == 1 else 0
        y_v = (3 * x_v + (0 if pi else c) + pi) % p
        y = cs.alloc_and_assign(y_v)
        cs.create_add_gate((x, 3), (cs.null_var, 0), y, None, 0 if pi else c, pi)
    x, x_v = y, y_v
compose_s = time.perf_counter() - t0
ks = [1, 7, 13, 17]
t0 = time.perf_counter()
srs = zm.universal_setup(ctx, curve, 1 << a.log_n, random.Random(2))
pk, vk = zp.keygen(ctx, srs, cs, ks, resident=not a.host_buffers)
ctx.sync()
keygen_s = time.perf_counter() - t0
proof, ch = zp.prove(ctx, pk, cs)               # warm-up: domains, pools
ctx.sync()
times, launches = [], []
for i in range(a.steps):
    l0 = ctx.launch_count
    t1 = time.perf_counter()
    proof, ch = zp.prove(ctx, pk, cs)
    ctx.sync()
    times.append(time.perf_counter() - t1)
    launches.append(ctx.launch_count - l0)
t1 = time.perf_counter()
cs._limbs = None
cs.witness_limbs()                                                    # int -> limb conversion of the assignment (once per circuit state)
synth_s = time.perf_counter() - t1
ok = zp.verifier_equality_check(ctx, pk.index, ch["beta"], ch["gamma"], ch["alpha"], ch["zeta"], ch["evals"], cs.public_inputs())
ms = sorted(times)[len(times) // 2] * 1e3
print(json.dumps({"metric": "plonk_prove_ms", "curve": "bls12_381" if curve == _lib.BLS12_381 else "bn254", "gates": cs.size(),
                  "domain_n": pk.index.n, "domain_4n": 4 * pk.index.n, "ms_per_proof": ms, "proofs_per_s": 1e3 / ms,
                  "assignment_to_limbs_once_ms": synth_s * 1e3, "gpu_launches_per_proof": launches[-1],
                  "round_state": "host buffers" if a.host_buffers else "resident in HBM",
                  "commitments": [len(r) for r in proof.commitments], "evaluations": len(proof.evaluations),
                  "equality_check_accepts": bool(ok), "compose_s": round(compose_s, 2), "keygen_s": round(keygen_s, 2)}))
pk.ck.free()
ctx.close()
