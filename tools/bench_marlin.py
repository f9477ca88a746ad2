#!/usr/bin/env python3
"""BASELINE configs[4]: Marlin prove, BN256, 2^18 constraints (MiMC chain with 2^18 - 4 real constraints so
that |H| = 2^18, |K| = 2^19, |B| = 2^21; committer key 3|K| - 2 powers) -- SURVEY.md 8d.

    python tools/bench_marlin.py [--log-h 18] [--curve 0] [--steps 2] [--no-verify]

Follows zkp_marlin::create_random_proof (marlin/src/lib.rs:97-181): prover_init, three AHP rounds each followed
by PC::commit, the evaluations at beta / gamma and PC::batch_open.  The verifier challenges and the blinding
draws come from a seeded generator (the Fiat-Shamir byte stream needs the Rust host); index and
committer key are built once outside the timed region, like `index()` / `universal_setup()` in the reference.
Every field / group operation runs on the GPU and the round state (assignment, oracles, index tables, committer
key) stays resident in HBM; only the scalars that feed the transcript (evaluations, commitments) reach the host.
`--host-buffers` times the same flow with host arrays between primitives (one H2D / D2H per primitive).
Full-size check (on by default): the committer key's trapdoor is known here, so every commitment must equal
(p(beta) + gamma * r(beta)) * G (shifted ones: beta^shift * p(beta) ...) -- evaluated on the GPU and compared
with the MSM results; the AHP identities themselves are checked against the oracle at small sizes in
tests/test_gpu_marlin.py.  Prints one JSON line.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from ckb_zkp_b200 import _lib, kzg10 as zk, marlin as zm, synth  # noqa: E402
from ckb_zkp_b200.backend import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-h", type=int, default=18)
ap.add_argument("--curve", type=int, default=_lib.BN254)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--no-verify", action="store_true")
ap.add_argument("--host-buffers", action="store_true", help="round state in host arrays (one H2D/D2H per primitive)")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
import torch  # noqa: E402
if world > 1:                              # weak scaling: every rank proves its own instance, no data-path collective
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
ctx = Context(local)

# per-primitive wall-clock accounting (host-buffer API: includes the copies of each call)
PROF = {}
for _name in ("msm", "ntt", "fr_vec_op", "fr_batch_inverse", "spmv", "fr_powers", "fr_convert", "poly_div_linear", "poly_lincomb",
              "poly_eval", "fixed_base_mul", "srs_upload"):
    def _wrap(fn, key):
        def inner(*args, **kw):
            t = time.perf_counter()
            try:
                return fn(*args, **kw)
            finally:
                c = PROF.setdefault(key, [0, 0.0])
                c[0] += 1
                c[1] += time.perf_counter() - t
        return inner
    setattr(ctx, _name, _wrap(getattr(ctx, _name), _name))


def prof_report(tag, total):
    if rank:
        PROF.clear()
        return
    acc = sum(v[1] for v in PROF.values())
    print("%s: %.3f s total, %.3f s inside backend calls; %s" % (tag, total, acc, ", ".join(
        "%s %dx %.3fs" % (k, v[0], v[1]) for k, v in sorted(PROF.items(), key=lambda kv: -kv[1][1]))), file=sys.stderr, flush=True)
    PROF.clear()

curve = a.curve
f = zm.Field(curve)
p = f.p
n = (1 << a.log_h) - 4                      # real constraints; n + 3 variables -> 3 empty constraints appended
t0 = time.perf_counter()
inst = synth.MimcInstance(curve, n, seed=synth.MIMC_SEED + rank)
A, B, C, z_mont = inst.device_form(ctx)
ni, nv = inst.n_inputs, inst.n_inputs + inst.n_aux
index, extra = zm.index(ctx, curve, A, B, C, ni, nv)
assert extra == 0 and index.h_size == 1 << a.log_h
index_s = time.perf_counter() - t0
prof_report('index', index_s)

# ---- committer key: powers beta^i * G and gamma * beta^i * G (kzg10.rs:27-72), exponents known -> checkable
t0 = time.perf_counter()
Hs, Ks = index.h_size, index.k_size
max_degree = max(3 * Hs + 2 - 1, 3 * Ks - 3)                       # AHP::max_degree (ahp/mod.rs:66-84)
rng = random.Random(2718)
beta_srs, gamma_srs = rng.randrange(1, p), rng.randrange(1, p)
gen = synth.generator_mont(curve, 1)


def power_points(scale):
    pw = ctx.fr_convert(curve, ctx.fr_powers(curve, f.mont(beta_srs), max_degree + 1, scale_mont=f.mont(scale)), to_mont=False)
    xs, infs = [], []
    for i in range(0, len(pw), 1 << 18):
        xy, inf = ctx.fixed_base_mul(curve, 1, gen, np.ascontiguousarray(pw[i:i + (1 << 18)]))
        xs.append(xy)
        infs.append(inf)
    return np.concatenate(xs), np.concatenate(infs)


ck = zk.CommitterKey(ctx, curve, power_points(1), power_points(gamma_srs), max_degree)
setup_s = time.perf_counter() - t0
prof_report('committer key', setup_s)


class Draws:
    """host RNG of the prover (zk_rng): scalar draws, and bulk draws for the 3|H|-coefficient mask polynomial"""

    def __init__(self, seed):
        self.r, self.np = random.Random(seed), np.random.default_rng(seed)

    def randrange(self, m):
        return self.r.randrange(m)

    def field_array(self, count):
        a = self.np.integers(0, np.iinfo(np.uint64).max, size=(count, 4), dtype=np.uint64, endpoint=True)
        a[:, 3] &= np.uint64((1 << (p.bit_length() - 1 - 192)) - 1)        # < 2^(bits - 1) < p
        return a


def outside_h(r):
    while True:
        t = r.randrange(p)
        if f.vanishing_at(Hs, t) != 0:
            return t


def prove(seed):
    """lib.rs:97-181 with seeded challenges; returns (commitments, evaluations, opening proofs, polynomials)"""
    zk_rng, ch = Draws(seed), random.Random(seed + 1)
    st = zm.prover_init(ctx, index, z_mont[:ni], z_mont[ni:], resident=not a.host_buffers)
    labeled, comms, rands = [], [], []

    def commit(round_polys):
        polys = [zk.LabeledPolynomial(label, poly, db, hb) for label, poly, db, hb in round_polys]
        c, r = zk.pc_commit(ck, polys, zk_rng)                         # lib.rs:109-110,117-118,124-125
        labeled.extend(polys)
        comms.extend(c)
        rands.extend(r)

    commit(zm.prover_first_round(st, zk_rng))
    alpha, etas = outside_h(ch), [ch.randrange(p) for _ in range(3)]
    commit(zm.prover_second_round(st, alpha, *etas))
    beta = outside_h(ch)
    commit(zm.prover_third_round(st, beta))
    gamma = ch.randrange(p)
    at_beta, at_gamma = ["w", "z_a", "z_b", "mask", "t", "g_1", "h_1"], ["g_2", "h_2"]
    query = [(l, beta) for l in at_beta] + [(l, gamma) for l in at_gamma]
    by_label = {P.label: P for P in labeled}
    evals = [ctx.poly_eval(curve, by_label[l].coeffs, f.mont(pt)) for l, pt in query]     # lib.rs:147-156
    opening_challenge = ch.randrange(1 << 128)
    proofs = zk.pc_batch_open(ck, labeled, query, opening_challenge, rands)                # lib.rs:160-166
    return comms, evals, proofs, by_label, (alpha, etas, beta, gamma), rands, labeled


t0 = time.perf_counter()
prove(100)                                   # warm-up: domains, pools
prof_report('warm-up prove', time.perf_counter() - t0)
launches0 = ctx.launch_count
times = []
for i in range(a.steps):
    t0 = time.perf_counter()
    out = prove(200 + i)
    ctx.sync()
    times.append(time.perf_counter() - t0)
    prof_report('prove', times[-1])
launches = (ctx.launch_count - launches0) // a.steps
sec = sum(times) / len(times)
if world > 1:
    t = torch.tensor([sec], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t[0])

checked = None
if not a.no_verify:
    comms, rands = out[0], out[5]
    checked = True
    bm = f.mont(beta_srs)

    def ev(poly):
        return f.to_int(ctx.poly_eval(curve, poly, bm)) if len(poly) else 0

    for P, (c, sc), (r, sr) in zip(out[6], comms, rands):
        e = (ev(P.coeffs) + gamma_srs * ev(r.blinding)) % p
        want = [(c, e)]
        if sc is not None:
            sh = pow(beta_srs, ck.supported_degree - P.degree_bound, p)
            want.append((sc, (sh * ev(P.coeffs) + gamma_srs * ev(sr.blinding)) % p))
        for got, e in want:
            xy, inf = ctx.fixed_base_mul(curve, 1, gen, synth.ints_to_limbs([e]))
            checked = checked and bool(inf[0]) == bool(got[1]) and (bool(got[1]) or np.array_equal(xy[0], got[0]))

line = {"metric": "marlin_proofs_per_sec_bn254_2e%d_constraints" % a.log_h, "value": world / sec, "unit": "proofs/s", "n_gpus": world,
        "scaling": "weak",
        "steps": a.steps, "ms_per_step": sec * 1e3, "higher_is_better": True, "data": "synthetic",
        "dtype": "u32 limbs (modular integer arithmetic, 254-bit Fr / Fq)",
        "config": {"workload": "Marlin prove, %s, MiMC chain with %d constraints: |H| = 2^%d, |K| = 2^%d, |B| = 2^%d, "
                               "committer key %d G1 powers (BASELINE configs[4] on one GPU)"
                               % ("BN254" if curve == _lib.BN254 else "BLS12-381", n, a.log_h, Ks.bit_length() - 1,
                                  index.b_size.bit_length() - 1, max_degree + 1),
                   "timing": "wall clock around create_random_proof's body, assignment uploaded inside the timed region; "
                             "round state %s; challenges and blinding draws seeded"
                             % ("in host arrays between primitives" if a.host_buffers else "resident in HBM"),
                   "commitments": len(out[0]), "openings": len(out[2])},
        "gpu_launches": launches, "index_s": round(index_s, 2), "setup_s": round(setup_s, 2), "verified": checked}
if rank == 0:
    print(json.dumps(line))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ck.free()
ctx.close()
