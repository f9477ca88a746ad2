#!/usr/bin/env python3
"""Profiling target: set up the 2^L-constraint BLS12-381 instance, run `--proofs` device-resident
proofs and print how many kernels one proof launches (so an ncu launch list can be cut per proof).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/prove_once.py --log-constraints 20 --proofs 2
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from ckb_zkp_b200 import synth  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-constraints", type=int, default=20)
ap.add_argument("--proofs", type=int, default=2)
ap.add_argument("--curve", type=int, default=1)
a = ap.parse_args()

ctx = Context(0)
n = 1 << a.log_constraints
inst = synth.MimcInstance(a.curve, n)
A, B, C, z = inst.device_form(ctx)
domain = 1 << (n + inst.n_inputs - 1).bit_length()
key = synth.SyntheticKey(inst.n_inputs + inst.n_aux, inst.n_inputs, domain, b_zero_cols=np.arange(4, 4 + n, 2))
params = key.upload(ctx, a.curve)
r = synth.ints_to_limbs([0x1234567])[0]
s = synth.ints_to_limbs([0x89ABCDE])[0]
ctx.groth16_stage(params.pk, A, B, C, z, inst.n_inputs, inst.n_aux)
ctx.sync()
before = ctx.launch_count
print("SETUP_LAUNCHES", before, flush=True)
for i in range(a.proofs):
    l0 = ctx.launch_count
    ctx.groth16_prove_staged(params.pk, r, s)
    ctx.sync()
    print("PROOF_LAUNCHES", ctx.launch_count - l0, flush=True)
ctx.groth16_fetch_proof(params.pk)
