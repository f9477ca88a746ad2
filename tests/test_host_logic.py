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

def local_msm(s):          # stands in for ctx.msm on this rank's resident shard
    P = c.to_affine(msm_naive(c, pts[lo:hi], s))
    xy, inf = H.points_array(cid, group, [P])
    return xy[0], bool(inf[0])

def fold(xy, inf):         # stands in for parallel.gpu_fold (EC additions in rank order)
    acc = c.identity()
    for P in H.array_points(cid, group, xy, inf):
        acc = c.add_mixed(acc, P)
    out, oinf = H.points_array(cid, group, [c.to_affine(acc)])
    return out[0], bool(oinf[0])

got = parallel.msm_sharded(local_msm, fold, sc[lo:hi], world, rank)
want = c.to_affine(msm_naive(c, pts, sc))
assert H.array_point(cid, group, got[0], got[1]) == want, (rank, "mismatch")
# a rank whose shard sums to the identity must still take part
zero = parallel.msm_sharded(lambda s: local_msm([0] * (hi - lo)) if rank == 0 else local_msm(s), fold, sc[lo:hi], world, rank)
lo1, hi1 = parallel.shard_range(n, world, 1)
want2 = c.to_affine(msm_naive(c, pts[lo1:hi1], sc[lo1:hi1]))
assert H.array_point(cid, group, zero[0], zero[1]) == want2
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
