"""CPU-side checks: the C-ABI library loads and exports every symbol include/zkb.h declares, the
host mirror of the reference's R1CS / prover interfaces, the synthetic workload generator, and the
multi-GPU sharding logic under gloo with world_size 2.  No GPU, no compute calls into libzkb."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from ckb_zkp_b200 import _lib, parallel, synth
from ckb_zkp_b200.r1cs import ONE, AssignmentMissing, ProvingAssignment, ints_to_limbs, limbs_to_int
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FR
from oracle.pyref.r1cs import ConstraintSystem, mimc_circuit, mini_circuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "zkb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()          # dlopen + getattr on every symbol
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_device_is_a_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from ckb_zkp_b200.backend import Context, ZkbError
    with pytest.raises(ZkbError):
        Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ckb_zkp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "zkref" not in src, f


def test_proving_assignment_matches_reference_layout():
    """ProvingAssignment (groth16/src/prover.rs:16-95) flattening == the oracle's CSR of the same circuit"""
    p = FR[BLS12_381].p
    pa = ProvingAssignment(p)
    pa.alloc_input(1)
    x = pa.alloc(lambda: 2)
    y = pa.alloc(lambda: 3)
    z = pa.alloc_input(lambda: 10)
    for _ in range(10):
        pa.enforce([(1, x)], [(1, y), (2, ONE)], [(1, z)])
    cs = mini_circuit(ConstraintSystem(p))
    assert pa.input_assignment + pa.aux_assignment == cs.full_assignment()
    for w in "abc":
        ptr, cols, vals = pa.csr(w)
        optr, ocols, ovals = cs.csr(w)
        assert list(ptr) == optr and list(cols) == ocols and vals == ovals
    with pytest.raises(AssignmentMissing):
        pa.alloc(lambda: None)
    assert limbs_to_int(ints_to_limbs([p - 1])[0]) == p - 1


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_synthetic_mimc_instance(cid):
    p = FR[cid].p
    for n in (2, 6, 128):
        inst = synth.MimcInstance(cid, n)
        cs = mimc_circuit(ConstraintSystem(p), n)
        assert cs.is_satisfied()
        assert inst.z == cs.full_assignment()
        assert (inst.n_inputs, inst.n_aux) == (2, n + 1)
        for w, W in zip("abc", "ABC"):
            ptr, cols, vals = cs.csr(w)
            rp, cc, _, _ = getattr(inst, W)
            assert list(rp) == ptr and list(cc) == cols and inst.coeff_ints(W) == vals
    c = CURVES[(cid, 1)]
    L = 4 if cid == BN254 else 6
    g = synth.generator_mont(cid, 1)
    R = (1 << (64 * L)) % c.F.p
    assert limbs_to_int(g[:L]) == c.gen[0] * R % c.F.p


@pytest.mark.parametrize("cid", [BN254, BLS12_381])
def test_synthetic_mimc_instance_other_seed_is_satisfied(cid):
    """every rank of bench.py proves MimcInstance(seed = MIMC_SEED + rank): the matrices must carry the round
    constants the witness was built with (A z o B z == C z), not the default seed's"""
    p = FR[cid].p
    inst = synth.MimcInstance(cid, 32, seed=synth.MIMC_SEED + 3)
    assert inst.consts != synth.MimcInstance(cid, 32).consts

    def mat_vec(which):
        rp, cc, _, _ = getattr(inst, which)
        vals = inst.coeff_ints(which)
        return [sum(vals[k] * inst.z[cc[k]] for k in range(rp[i], rp[i + 1])) % p for i in range(len(rp) - 1)]

    az, bz, cz = mat_vec("A"), mat_vec("B"), mat_vec("C")
    assert all(x * y % p == w for x, y, w in zip(az, bz, cz))


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1 << 24, (1 << 24) + 5):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from ckb_zkp_b200 import parallel
from ckb_zkp_b200.backend import Context
from oracle.pyref.curves import CURVES
from oracle.pyref.msm import msm_naive
from tests import helpers as H

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
cid, group, n = 0, 1, 37
c = CURVES[(cid, group)]
pts = H.multiples(cid, group, n, start=4)
pts[5] = None
sc = [(i * 7919 + 13) ** 5 %% c.r for i in range(n)]
lo, hi = parallel.shard_range(n, world, rank)
W = H.points_array(cid, group, [pts[0]])[0].shape[1]

def record(P):             # a partial point as a byte record (what zkb_msm_partial hands to a transport)
    xy, inf = H.points_array(cid, group, [P])
    return np.concatenate([xy[0].view(np.uint8), inf.astype(np.uint8)])

def partial(scalars=None):  # stands in for ctx.msm_partial on this rank's resident shard
    return record(c.to_affine(msm_naive(c, pts[lo:hi], sc[lo:hi] if scalars is None else scalars)))

def fold(records):         # stands in for ctx.msm_fold (EC additions in rank order)
    acc = c.identity()
    for rec in records:
        P = H.array_point(cid, group, np.ascontiguousarray(rec[:8 * W]).view(np.uint64), bool(rec[8 * W]))
        acc = c.add_mixed(acc, P)
    return c.to_affine(acc)

got = parallel.msm_sharded_via(partial, fold, parallel.all_gather_bytes, world)
assert got == c.to_affine(msm_naive(c, pts, sc)), (rank, "mismatch")
# a rank whose shard sums to the identity must still take part
zero = parallel.msm_sharded_via(lambda: partial([0] * (hi - lo)) if rank == 0 else partial(), fold,
                                parallel.all_gather_bytes, world)
lo1, hi1 = parallel.shard_range(n, world, 1)
assert zero == c.to_affine(msm_naive(c, pts[lo1:hi1], sc[lo1:hi1]))

# rendezvous of the library's communicator: rank 0's id reaches every rank unchanged
class FakeCtx:
    comm_init_torch = Context.comm_init_torch
    def comm_unique_id(self):
        return (np.arange(128, dtype=np.uint8) * 3 + 1).astype(np.uint8)
    def comm_init(self, n_ranks, rank, unique_id=None):
        self.got = (n_ranks, rank, None if unique_id is None else np.array(unique_id))
f = FakeCtx()
assert f.comm_init_torch() == (world, rank)
assert f.got[:2] == (world, rank) and np.array_equal(f.got[2], (np.arange(128, dtype=np.uint8) * 3 + 1).astype(np.uint8))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_msm_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "rank %d ok" % rank in out
