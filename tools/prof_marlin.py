#!/usr/bin/env python3
"""Where does a Marlin proof spend its host time?  cProfile over ckb_zkp_b200.marlin.create_random_proof at |H| = 2^log_h
(the bench.py `marlin` sub-record's workload), sorted by cumulative time, plus wall time per C-ABI entry point."""
import argparse
import cProfile
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from ckb_zkp_b200 import _lib, marlin as zm, synth  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-h", type=int, default=18)
a = ap.parse_args()
ctx = Context(0)
curve = _lib.BN254
p = synth.FR_MODULUS[curve]
n = (1 << a.log_h) - 4
inst = synth.MimcInstance(curve, n)
circuit = bench._MarlinCircuit(ctx, inst)
need = 3 * (1 << (a.log_h + 1)) - 3
srs = zm.universal_setup(ctx, curve, need, random.Random(2718))
ipk, ivk = zm.index_keys(ctx, srs, circuit)
for i in range(2):
    zm.create_random_proof(ctx, ipk, circuit, bench._MarlinDraws(100 + i, p))
ctx.sync()

calls = {}
lib = ctx.lib
for name in list(_lib.SIGNATURES):
    fn = getattr(lib, name)
    def wrap(fn=fn, name=name):
        def inner(*args):
            t = time.perf_counter()
            r = fn(*args)
            c = calls.setdefault(name, [0, 0.0])
            c[0] += 1
            c[1] += time.perf_counter() - t
            return r
        return inner
    setattr(lib, name, wrap())
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
zm.create_random_proof(ctx, ipk, circuit, bench._MarlinDraws(300, p))
ctx.sync()
pr.disable()
print("proof wall ms (under cProfile): %.1f" % ((time.perf_counter() - t0) * 1e3))
print("C-ABI calls: total %.1f ms" % (sum(v[1] for v in calls.values()) * 1e3))
for k, (cnt, sec) in sorted(calls.items(), key=lambda kv: -kv[1][1])[:14]:
    print("  %-28s n=%4d %8.2f ms" % (k, cnt, sec * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
