#!/usr/bin/env python3
"""Tuning sweep of the batched-affine pair levels (csrc/msm_affine.cuh) on one B200: a 2^L G1 (and 2^(L-1) G2) MSM of
BLS12-381 with the scalars resident, CUDA events, for ZKB_MSM_PAIR_LEVELS x ZKB_PAIR_SCALE.  One JSON line per setting;
every result is compared with the levels = 0 result (bit-identical or the line says so)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from ckb_zkp_b200 import synth  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=20)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--levels", type=int, nargs="*", default=[0, 1, 2, 3])
ap.add_argument("--scales", type=int, nargs="*", default=[0, 1, 2, 4])
ap.add_argument("--groups", type=int, nargs="*", default=[1, 2])
ap.add_argument("--batch", type=int, nargs="*", default=[0], help="ZKB_MSM_BATCH values (batched-affine accumulation, msm_batch.cuh)")
ap.add_argument("--curve", type=int, default=1, help="0 = BN254, 1 = BLS12-381")
a = ap.parse_args()
ctx = Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(11)
for group in a.groups:
    n = 1 << (a.log_n if group == 1 else a.log_n - 1)
    gen = synth.generator_mont(a.curve, group)
    xs, infs = [], []
    for i in range(0, n, 1 << 18):
        xy, inf = ctx.fixed_base_mul(a.curve, group, gen, synth.random_exponents(rng, min(1 << 18, n - i)))
        xs.append(xy); infs.append(inf)
    srs = ctx.srs_upload(a.curve, group, np.concatenate(xs), np.concatenate(infs))
    d = torch.from_numpy(synth.random_scalars(rng, n, a.curve).view(np.int64)).cuda()
    base = None
    for lv, bt in [(lv, bt) for bt in a.batch for lv in a.levels]:
        for sc in (a.scales if lv else [0]):
            os.environ["ZKB_MSM_BATCH"] = str(bt)
            os.environ["ZKB_MSM_PAIR_LEVELS"] = str(lv)
            os.environ["ZKB_PAIR_SCALE"] = str(sc)
            for _ in range(2):
                res = ctx.msm_dev(srs, d.data_ptr(), n)
            if base is None:
                base = res
            same = res[1] == base[1] and bool(np.array_equal(res[0], base[0]))
            st = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
            en = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
            for i in range(a.steps):
                flush.fill_(i)
                torch.cuda.synchronize()
                with torch.cuda.stream(stream):
                    st[i].record()
                    ctx.msm_dev(srs, d.data_ptr(), n)
                    en[i].record()
            torch.cuda.synchronize()
            ms = sorted(x.elapsed_time(y) for x, y in zip(st, en))
            print(json.dumps({"group": group, "log_n": n.bit_length() - 1, "batch_affine": bt, "pair_levels": lv, "pair_scale": sc,
                              "ms_median": ms[len(ms) // 2], "ms_min": ms[0], "same_result_as_levels0": same}), flush=True)
    srs.free()
    del d
ctx.close()
