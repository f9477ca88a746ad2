#!/usr/bin/env python3
"""Kernel timeline of one device-resident proof (CUPTI through torch.profiler; there is no nsys in
the image).  Prints, per stream, when each kernel started and ended relative to the first kernel of the
proof, plus the busy time per kernel name and the union busy time -- the concurrency picture that the
serialised ncu launch list cannot give.

    python tools/timeline.py --log-constraints 20 --out gpurun_out/timeline.txt
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from ckb_zkp_b200 import synth  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-constraints", type=int, default=20)
ap.add_argument("--curve", type=int, default=1)
ap.add_argument("--out", default="gpurun_out/timeline.txt")
ap.add_argument("--sharded", action="store_true", help="under torchrun: ONE proof by all ranks (zkb_groth16_prove_sharded), rank 0's timeline")
ap.add_argument("--partial-of", type=int, default=0, help="single GPU: rank 0's share of a proof sharded over N ranks (zkb_groth16_prove_partial, no exchange)")
a = ap.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
torch.cuda.init()
ctx = Context(local)
shard = None
if a.sharded:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx.comm_init_torch()
    shard = (world, rank)
n = 1 << a.log_constraints
inst = synth.MimcInstance(a.curve, n)
A, B, C, z = inst.device_form(ctx)
domain = 1 << (n + inst.n_inputs - 1).bit_length()
key = synth.SyntheticKey(inst.n_inputs + inst.n_aux, inst.n_inputs, domain, b_zero_cols=np.arange(4, 4 + n, 2))
if a.partial_of:
    shard = (a.partial_of, 0)
params = key.upload(ctx, a.curve, shard=shard)
r = synth.ints_to_limbs([0x1234567])[0]
s = synth.ints_to_limbs([0x89ABCDE])[0]
ctx.groth16_stage(params.pk, A, B, C, z, inst.n_inputs, inst.n_aux)
prove = ctx.groth16_prove_sharded_staged if a.sharded else ctx.groth16_prove_staged
if a.partial_of:
    prove = lambda pk, r_, s_: ctx.groth16_prove_partial(pk, A, B, C, z, inst.n_inputs, inst.n_aux, r_, s_)
for _ in range(3):
    prove(params.pk, r, s)
ctx.sync()
if a.sharded:
    dist.barrier()

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    prove(params.pk, r, s)
    ctx.sync()
    torch.cuda.synchronize()
if rank != 0:
    sys.exit(0)

tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
trace = json.load(open(tmp))
evs = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
evs.sort(key=lambda e: e["ts"])
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
with open(a.out, "w") as f:
    if not evs:
        f.write("no device activity captured\n")
        sys.exit(0)
    t0 = evs[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in evs)
    f.write("proof span: %.3f ms, %d device activities\n" % ((end - t0) / 1e3, len(evs)))
    # union busy time (any kernel running)
    busy, cur_s, cur_e = 0.0, None, None
    for e in evs:
        s_, e_ = e["ts"], e["ts"] + e["dur"]
        if cur_e is None or s_ > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    f.write("union busy: %.3f ms\n" % (busy / 1e3))
    by = {}
    for e in evs:
        k = e["name"].split("(")[0][:70]
        by.setdefault(k, [0, 0.0])
        by[k][0] += 1
        by[k][1] += e["dur"]
    f.write("\nper kernel (concurrent durations, not serialised):\n")
    for k, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        f.write("  %-70s n=%3d %9.3f ms\n" % (k, c, d / 1e3))
    f.write("\ntimeline (start ms, dur ms, stream, name):\n")
    for e in evs:
        f.write("  %9.3f %8.3f  s%-4s %s\n" % ((e["ts"] - t0) / 1e3, e["dur"] / 1e3, e.get("args", {}).get("stream", "?"),
                                             e["name"].split("(")[0][:80]))
print(open(a.out).read()[:6000])
