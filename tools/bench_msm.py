#!/usr/bin/env python3
"""BASELINE configs[2]: standalone G1 MSM, BLS12-381, 2^L bases, sharded over N B200s (one process per
GPU; `python -m torch.distributed.run --nproc-per-node N tools/bench_msm.py --log-n 24`).

Bases are k_i * G for seeded pseudo-random k_i (so the result is checkable in the exponent), scalars are
seeded full-width residues or the boolean-heavy mix of SURVEY.md 8d (`--dist bool`).  Each rank keeps its
contiguous shard of the bases (with window tables) resident, the timed region covers the local MSM with
scalars resident in HBM, the all_gather of one partial point per rank and the fold.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ckb_zkp_b200 import parallel, synth  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=24)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--curve", type=int, default=1)
ap.add_argument("--dist", default="full", choices=["full", "bool"])
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
ctx = Context(local)
curve, group = a.curve, 1
n = 1 << a.log_n
p = synth.FR_MODULUS[curve]
lo, hi = parallel.shard_range(n, world, rank)

# ---- shard of the bases and scalars (seeded per global 2^18 block so every world size sees the same data)
BLK = 1 << 18
gen = synth.generator_mont(curve, group)
xs, infs, ks, ss = [], [], [], []
for b0 in range(lo - lo % BLK, hi, BLK):
    rng = np.random.default_rng(1000 + b0 // BLK)
    k = synth.random_exponents(rng, BLK)
    s = synth.random_exponents(rng, BLK)
    if a.dist == "bool":
        u = rng.random(BLK)
        s[u < 0.5, 1:] = 0
        s[u < 0.5, 0] &= np.uint64(1)
        mid = (u >= 0.5) & (u < 0.75)
        s[mid, 1:] = 0
        s[mid, 0] &= np.uint64(0xFFFF)
    sl = slice(max(lo, b0) - b0, min(hi, b0 + BLK) - b0)
    k, s = k[sl], s[sl]
    xy, inf = ctx.fixed_base_mul(curve, group, gen, k)
    xs.append(xy); infs.append(inf); ks.append(k); ss.append(s)
xy, inf, k_local, s_local = (np.concatenate(v) for v in (xs, infs, ks, ss))
t0 = time.perf_counter()
shard = parallel.ShardedSrs(ctx, curve, group, xy, inf, n, world, rank)
upload_s = time.perf_counter() - t0
del xy
d_scalars = torch.from_numpy(s_local.view(np.int64)).to(dev)
fold = parallel.gpu_fold(ctx, curve, group)
local_msm = lambda _s: ctx.msm_dev(shard.srs, d_scalars.data_ptr(), hi - lo)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(max(3, a.warmup)):
    res = parallel.msm_sharded(local_msm, fold, None, world, rank, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
barrier()
ctx.prof_enable(True)
times = []
for i in range(a.steps):
    flush.fill_(i)
    barrier()
    t0 = time.perf_counter()
    res = parallel.msm_sharded(local_msm, fold, None, world, rank, dev)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
prof = ctx.prof_read()
t = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
total_s = float(t.item())

# ---- check in the exponent: result == (sum s_i k_i mod r) * G
e = sum(x * y for x, y in zip(synth.limbs_to_ints(s_local), synth.limbs_to_ints(k_local))) % p
et = torch.from_numpy(synth.ints_to_limbs([e]).view(np.int64)).to(dev)
if world > 1:
    parts = [torch.zeros_like(et) for _ in range(world)]
    dist.all_gather(parts, et)
    e = sum(synth.limbs_to_ints(x.cpu().numpy().view(np.uint64))[0] for x in parts) % p
want_xy, want_inf = ctx.fixed_base_mul(curve, group, gen, synth.ints_to_limbs([e]))
ok = bool(want_inf[0]) == res[1] and (res[1] or np.array_equal(want_xy[0], res[0]))

if rank == 0:
    ms = total_s / a.steps * 1e3
    c_ref = 3 if n < 32 else ((n - 1).bit_length() * 69 // 100 + 2)
    w_ref = -(-255 // c_ref)
    ref_adds = n * w_ref + 2 * ((1 << c_ref) - 1) * w_ref
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
    alg = n * 128.0
    line = {"metric": "msm_g1_bls12_381_reference_g1_adds_per_sec", "value": ref_adds / (ms * 1e-3), "unit": "G1-adds/s",
            "n_gpus": world, "steps": a.steps, "ms_per_step": ms, "scaling": "strong", "verified_in_exponent": ok,
            "config": {"workload": "G1 MSM BLS12-381 2^%d bases, %s scalars, sharded contiguously over %d GPU(s), "
                                   "all_gather of partial points + EC-add fold" % (a.log_n, a.dist, world),
                       "g1_adds_definition": "mixed + reduction additions of the reference algorithm (ark-ec 0.2, c=%d, %d "
                                             "windows): n*W + 2*(2^c-1)*W = %d" % (c_ref, w_ref, ref_adds),
                       "timing": "wall clock around msm_sharded with device sync, max over ranks, L2 flushed"},
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / (peak * world), "algorithmic_bytes": alg,
                         "k_accumulate_ms_per_launch_rank0": prof["ms"] / max(prof["launches"], 1)},
            "srs_upload_s": round(upload_s, 2)}
    print(json.dumps(line))
shard.free()
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
