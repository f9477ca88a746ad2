#!/usr/bin/env python3
"""One rank of the multi-GPU parity run (needs >= 2 B200s; launched by tests/test_gpu_sharded.py::test_nccl_two_ranks
or by hand:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_worker.py).

Every check goes through the collective entry points of the C ABI (NCCL all-gather inside libzkb.so) and compares
with the committed golden vectors of the oracle, or with the same GPU's unsharded result at sizes without a golden."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ckb_zkp_b200 import groth16 as zg, parallel, synth  # noqa: E402
from ckb_zkp_b200.backend import Context, CsrMatrix  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
assert ctx.comm_init_torch() == (world, rank)
assert (ctx.comm_size, ctx.comm_rank) == (world, rank)
GOLD = os.path.join(ROOT, "tests", "golden")
load = lambda name: np.load(os.path.join(GOLD, name + ".npz"))

# ---- sharded MSM against the golden results
for name, group in (("msm_bls12_381_g1_256", 1), ("msm_bn254_g1_256", 1), ("msm_bls12_381_g2_64", 2)):
    g = load(name)
    cid, n = int(g["curve"]), len(g["bases_inf"])
    lo, hi = parallel.shard_range(n, world, rank)
    sh = parallel.ShardedSrs(ctx, cid, group, g["bases_xy"][lo:hi], g["bases_inf"][lo:hi], n, world, rank)
    c0 = ctx.collective_count
    xy, inf = sh.msm(g["scalars"])
    assert ctx.collective_count == c0 + 1, "exactly one all-gather per sharded MSM"
    assert inf == bool(g["result_inf"][0]) and np.array_equal(xy, g["result_xy"][0]), (name, "sharded msm")
    # scalars sharded like the bases and already on the device
    d = torch.from_numpy(np.ascontiguousarray(g["scalars"][lo:hi]).view(np.int64)).cuda()
    xy, inf = sh.msm_local(d.data_ptr())
    assert inf == bool(g["result_inf"][0]) and np.array_equal(xy, g["result_xy"][0]), (name, "sharded msm, local scalars")
    # a window of the logical SRS (skip_leading_zeros / shifted commitments of kzg10.rs)
    whole = ctx.srs_upload(cid, group, g["bases_xy"], g["bases_inf"])
    want = ctx.msm(whole, g["scalars"][:40], base_offset=n // 2 - 7)
    got = sh.msm(g["scalars"][:40], base_offset=n // 2 - 7)
    assert got[1] == want[1] and np.array_equal(got[0], want[0]), (name, "offset window")
    whole.free()
    sh.free()


# ---- one Groth16 proof by all ranks against the golden proofs
def params_from(g, shard):
    q = lambda k: (g[k + "_xy"], g[k + "_inf"])
    s1, s2 = g["g1_singles"], g["g2_singles"]
    return zg.Parameters(ctx, int(g["curve"]), q("a_query"), q("b_g1_query"), q("b_g2_query"), q("h_query"), q("l_query"),
                         s1[0], s1[1], s1[2], s2[0], s2[1], shard=shard)


for name in ("groth16_mini_bls12_381", "groth16_mini_bn254", "groth16_mimc_bls12_381_2e6", "groth16_mimc_bn254_2e10"):
    g = load(name)
    A, B, C = [CsrMatrix(g[w + "_ptr"], g[w + "_col"], g[w + "_val"]) for w in "abc"]
    params = params_from(g, (world, rank))
    c0 = ctx.collective_count
    proof = ctx.groth16_prove_sharded(params.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]), g["r"][0], g["s"][0])
    assert ctx.collective_count == c0 + 1, "exactly one all-gather per sharded proof"
    for key, got in (("proof_a", proof[0]), ("proof_b", proof[1]), ("proof_c", proof[2])):
        assert bool(g[key + "_inf"][0]) == got[1] and (got[1] or np.array_equal(g[key + "_xy"][0], got[0])), (name, key)
    params.free()

# ---- mid size (2^14 constraints, BLS12-381): sharded proof == this GPU's own unsharded proof, staged path twice
n = 1 << 14
inst = synth.MimcInstance(1, n)
A, B, C, z = inst.device_form(ctx)
key = synth.SyntheticKey(inst.n_inputs + inst.n_aux, inst.n_inputs, 2 * n, b_zero_cols=np.arange(4, 4 + n, 2))
whole = key.upload(ctx, 1)
sharded = key.upload(ctx, 1, shard=(world, rank))
r, s = synth.ints_to_limbs([0xABCDEF123 + 0])[0], synth.ints_to_limbs([0x13579BDF])[0]
want = ctx.groth16_prove(whole.pk, A, B, C, z, inst.n_inputs, inst.n_aux, r, s)
ctx.groth16_stage(sharded.pk, A, B, C, z, inst.n_inputs, inst.n_aux)
for _ in range(2):
    ctx.groth16_prove_sharded_staged(sharded.pk, r, s)
    got = ctx.groth16_fetch_proof(sharded.pk)
    for a, b in zip(got, want):
        assert a[1] == b[1] and np.array_equal(a[0], b[0]), "2^14 sharded proof differs from the unsharded one"
whole.free()
sharded.free()
dist.barrier()
ctx.close()
dist.destroy_process_group()
print("rank", rank, "ok", flush=True)
