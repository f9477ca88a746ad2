"""Multi-GPU sharding of the prove path, checked on ONE GPU: the ranks' work is run one after the other on the
same device (sharded SRS / sharded proving keys for rank 0 .. k-1), their partial records are concatenated as the
all-gather would and folded by the library -- everything except the ncclAllGather call itself, which
tests/multi_gpu_worker.py covers on 2+ GPUs (launched by test_nccl_two_ranks when two devices are visible).

Oracle: the committed golden MSM / Groth16 fixtures (tests/golden/, minted by the Python oracle)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from ckb_zkp_b200 import groth16 as zg
from ckb_zkp_b200 import parallel
from ckb_zkp_b200.backend import CsrMatrix, ZkbError

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name,group", [("msm_bls12_381_g1_256", 1), ("msm_bn254_g1_256", 1), ("msm_bls12_381_g2_64", 2)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_msm_partials_fold_to_the_golden_result(ctx, name, group, world):
    g = load(name)
    cid = int(g["curve"])
    n = len(g["bases_inf"])
    recs = []
    for rank in range(world):
        lo, hi = parallel.shard_range(n, world, rank)
        srs = ctx.srs_upload_shard(cid, group, g["bases_xy"][lo:hi], g["bases_inf"][lo:hi], lo, n)
        recs.append(ctx.msm_partial(srs, g["scalars"]))
        srs.free()
    xy, inf = ctx.msm_fold(cid, group, np.stack(recs))
    assert inf == bool(g["result_inf"][0])
    assert np.array_equal(xy, g["result_xy"][0])


def test_msm_partial_offsets_and_truncation(ctx):
    """base_offset / zip truncation over the LOGICAL SRS (kzg10.rs:107-121 skip_leading_zeros and shifted commitments):
    every rank sees the same scalar array and reads only the slice its bases pair with"""
    g = load("msm_bls12_381_g1_256")
    cid, n = int(g["curve"]), len(g["bases_inf"])
    whole = ctx.srs_upload(cid, 1, g["bases_xy"], g["bases_inf"])
    for off, cnt in ((0, n), (5, 100), (100, n), (250, 3), (17, 0), (256, 4)):
        sc = g["scalars"][:cnt]
        want = ctx.msm(whole, sc, base_offset=off)
        for world in (2, 5):
            recs = []
            for rank in range(world):
                lo, hi = parallel.shard_range(n, world, rank)
                srs = ctx.srs_upload_shard(cid, 1, g["bases_xy"][lo:hi], g["bases_inf"][lo:hi], lo, n)
                recs.append(ctx.msm_partial(srs, sc, base_offset=off))
                srs.free()
            got = ctx.msm_fold(cid, 1, np.stack(recs))
            assert got[1] == want[1] and np.array_equal(got[0], want[0]), (off, cnt, world)
    whole.free()
    with pytest.raises(ZkbError):
        ctx.srs_upload_shard(cid, 1, g["bases_xy"][:10], g["bases_inf"][:10], 250, n)      # shard sticks out of the logical SRS


def test_single_rank_communicator_is_a_copy(ctx):
    """n_ranks == 1: the sharded entry points work without NCCL (device-to-device copy instead of the all-gather)"""
    g = load("msm_bn254_g1_256")
    cid, n = int(g["curve"]), len(g["bases_inf"])
    assert ctx.comm_size == 1 and ctx.comm_rank == 0
    srs = ctx.srs_upload_shard(cid, 1, g["bases_xy"], g["bases_inf"], 0, n)
    xy, inf = ctx.msm_sharded(srs, g["scalars"])
    assert inf == bool(g["result_inf"][0]) and np.array_equal(xy, g["result_xy"][0])
    srs.free()


def _golden_params(ctx, g, shard=None):
    cid = int(g["curve"])
    q = lambda k: (g[k + "_xy"], g[k + "_inf"])
    s1, s2 = g["g1_singles"], g["g2_singles"]
    return zg.Parameters(ctx, cid, q("a_query"), q("b_g1_query"), q("b_g2_query"), q("h_query"), q("l_query"), s1[0],
                         s1[1], s1[2], s2[0], s2[1], shard=shard)


def _assert_proof(g, proof):
    for key, got in (("proof_a", proof[0]), ("proof_b", proof[1]), ("proof_c", proof[2])):
        assert bool(g[key + "_inf"][0]) == got[1], key
        if not got[1]:
            assert np.array_equal(g[key + "_xy"][0], got[0]), key


@pytest.mark.parametrize("name", ["groth16_mini_bls12_381", "groth16_mini_bn254", "groth16_mimc_bls12_381_2e6",
                                  "groth16_mimc_bn254_2e10"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_groth16_proof_is_the_golden_proof(ctx, name, world):
    """one proof computed from `world` per-rank partials (A_k, C_k, B2_k) == the oracle's proof, bit for bit"""
    g = load(name)
    A, B, C = [CsrMatrix(g[w + "_ptr"], g[w + "_col"], g[w + "_val"]) for w in "abc"]
    ni, na = int(g["n_inputs"]), int(g["n_aux"])
    r, s = g["r"][0], g["s"][0]
    recs, keys = [], []
    for rank in range(world):
        params = _golden_params(ctx, g, shard=(world, rank))
        recs.append(ctx.groth16_prove_partial(params.pk, A, B, C, g["z"], ni, na, r, s))
        keys.append(params)
    _assert_proof(g, ctx.groth16_fold(keys[0].pk, np.stack(recs), r, s))
    if world == 1:           # the collective entry point with a one-rank communicator
        _assert_proof(g, ctx.groth16_prove_sharded(keys[0].pk, A, B, C, g["z"], ni, na, r, s))
        proof = zg.create_proof(keys[0], _Mini(), 0, 0) if name.startswith("groth16_mini") else None
        if proof is not None:
            whole = _golden_params(ctx, g)
            assert proof == zg.create_proof(whole, _Mini(), 0, 0)      # r = 0: the guard of prover.rs:170
            whole.free()
    for k in keys:
        k.free()


class _Mini:
    """groth16/tests/mini.rs:12-44 with x = 2, y = 3, z = 10, num = 10"""

    def generate_constraints(self, cs):
        from ckb_zkp_b200.r1cs import ONE
        vx = cs.alloc(lambda: 2)
        vy = cs.alloc(lambda: 3)
        vz = cs.alloc_input(lambda: 10)
        for _ in range(10):
            cs.enforce([(1, vx)], [(1, vy), (2, ONE)], [(1, vz)])


def test_sharded_key_on_the_wrong_communicator_is_an_error(ctx):
    g = load("groth16_mini_bn254")
    A, B, C = [CsrMatrix(g[w + "_ptr"], g[w + "_col"], g[w + "_val"]) for w in "abc"]
    params = _golden_params(ctx, g, shard=(2, 1))
    with pytest.raises(ZkbError):       # key sharded for rank 1 of 2, communicator is rank 0 of 1
        ctx.groth16_prove_sharded(params.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]), g["r"][0], g["s"][0])
    whole = _golden_params(ctx, g)
    with pytest.raises(ZkbError):       # an unsharded key through the sharded entry point
        ctx.groth16_prove_sharded(whole.pk, A, B, C, g["z"], int(g["n_inputs"]), int(g["n_aux"]), g["r"][0], g["s"][0])
    whole.free()
    params.free()


def test_nccl_two_ranks():
    """the real thing when the box has >= 2 GPUs: torchrun-style launch of tests/multi_gpu_worker.py (NCCL all-gather
    inside libzkb.so); on a one-GPU box only the single-GPU simulation above runs"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the ncclAllGather itself needs two (covered by tests/multi_gpu_worker.py under "
                    "gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for rank in range(world):
        assert "rank %d ok" % rank in out.stdout
