#!/usr/bin/env python3
"""BASELINE configs[3]: Fr NTT sweep 2^16..2^24 on BN254 and BLS12-381, one B200.  Data resident in HBM,
CUDA events on the library stream, L2 flushed between iterations; checks ifft(fft(x)) == x per size.
Prints one JSON line per (field, size, variant)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--min-log", type=int, default=16)
ap.add_argument("--max-log", type=int, default=24)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()

ctx = Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
rng = np.random.default_rng(4)
for curve, name in ((0, "bn254"), (1, "bls12_381")):
    for log_n in range(a.min_log, a.max_log + 1):
        n = 1 << log_n
        host = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        host[:, 3] &= np.uint64((1 << 60) - 1)
        d = torch.from_numpy(host.view(np.int64)).to(dev)
        orig = d.clone()
        torch.cuda.synchronize()
        for variant, kw in (("fft", {}), ("coset_ifft", {"inverse": True, "coset": True})):
            for _ in range(3):
                ctx.ntt_dev(curve, d.data_ptr(), log_n, **kw)
            ctx.sync()
            ms = 0.0
            for i in range(a.steps):
                flush.fill_(i)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    s.record()
                    ctx.ntt_dev(curve, d.data_ptr(), log_n, **kw)
                    e.record()
                torch.cuda.synchronize()
                ms += s.elapsed_time(e)
            ms /= a.steps
            alg = 2.0 * n * 32
            print(json.dumps({"metric": "ntt_ms", "field": name + "_fr", "log_n": log_n, "variant": variant, "ms": ms,
                              "butterflies_per_s": n / 2 * log_n / (ms * 1e-3),
                              "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                           "frac": alg / (ms * 1e-3) / 1e9 / peak}}), flush=True)
        # round trip on fresh data
        d.copy_(orig)
        torch.cuda.synchronize()     # torch's stream and the library's non-blocking stream do not order implicitly
        ctx.ntt_dev(curve, d.data_ptr(), log_n)
        ctx.ntt_dev(curve, d.data_ptr(), log_n, inverse=True)
        ctx.sync()
        assert torch.equal(d, orig), (name, log_n)
        del d, orig
ctx.close()
