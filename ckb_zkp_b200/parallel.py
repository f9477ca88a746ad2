"""Multi-GPU sharding of the prove path: one process per GPU, the collective lives INSIDE the C-ABI library
(csrc/comm.cu: one ncclAllGather of fixed-size partial points over NVLink, folded by a kernel in rank order).

An MSM is a plain sum over (scalar, base) pairs (curve/src/lib.rs:38-45), so any partition of the pairs is
valid; the five Groth16 MSMs share only read-only inputs (groth16/src/prover.rs:164-190) and Marlin commits
polynomial by polynomial (marlin/src/pc/mod.rs:42-69).  Each rank keeps a contiguous slice of the bases resident
in its own HBM and every rank ends with the identical canonical affine result.

What is left on the host is the partition rule (`shard_range`, mirrored by the library for the sharded Groth16 key)
and the rendezvous: rank 0's 128-byte NCCL id travels over whatever process group the host already has
(`Context.comm_init_torch` uses torch.distributed; gloo in the CPU tests).  `msm_sharded_via` runs the same
partial -> gather -> fold sequence over a caller-supplied transport (the two halves zkb_msm_partial / zkb_msm_fold);
the CPU tests drive it over gloo, a host with MPI would do the same.
"""
import numpy as np


def shard_range(n, world, rank):
    """contiguous, balanced partition of range(n): sizes differ by at most one, earlier ranks larger"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedSrs:
    """The local shard [lo, hi) of a logical SRS of `n_total` bases (zkb_srs_upload_shard)."""

    def __init__(self, ctx, curve, group, xy_local, inf_local, n_total, world, rank, precompute=True):
        self.lo, self.hi = shard_range(n_total, world, rank)
        if len(inf_local) != self.hi - self.lo:
            raise ValueError("local shard has %d bases, expected %d" % (len(inf_local), self.hi - self.lo))
        self.n_total, self.world, self.rank = n_total, world, rank
        self.curve, self.group, self.ctx = curve, group, ctx
        self.srs = ctx.srs_upload_shard(curve, group, xy_local, inf_local, self.lo, n_total, precompute=precompute)

    def msm(self, scalars, base_offset=0, mont=False):
        """VariableBaseMSM::multi_scalar_mul over the logical SRS; collective, every rank passes the same scalars"""
        return self.ctx.msm_sharded(self.srs, scalars, base_offset=base_offset, mont=mont)

    def msm_local(self, d_scalars_ptr):
        """stand-alone sharded MSM: this rank's canonical scalars (device pointer) pair with its own bases"""
        return self.ctx.msm_sharded_local(self.srs, d_scalars_ptr, self.hi - self.lo)

    def free(self):
        self.srs.free()


def all_gather_bytes(mine, world):
    """all_gather of one fixed-size byte record per rank over torch.distributed -> uint8[world, len]"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(mine, dtype=np.uint8))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.stack(parts).cpu().numpy()


def msm_sharded_via(partial, fold, gather, world):
    """partial() -> this rank's partial record; gather(record, world) -> all records in rank order;
    fold(records) -> result.  world == 1 still folds (one record): the code path is the same on every world size."""
    mine = partial()
    return fold(gather(mine, world) if world > 1 else np.ascontiguousarray(mine)[None, :])
