"""Multi-GPU sharding of one large MSM (BASELINE configs[2]: 2^24 bases over 1/2/4/8 B200).

An MSM is a plain sum over (scalar, base) pairs (curve/src/lib.rs:38-45), so any partition of the pairs
is valid.  Each rank keeps a contiguous shard of the bases resident in its own HBM, runs the full
single-GPU MSM on its shard, and the ranks exchange ONE fixed-size partial each:

    all_gather(partial affine point + identity flag)  ->  fold with group additions in rank order

NCCL has no user-defined reduction and EC addition is not ncclSum, so the "allreduce of partials" is an
all-gather of `world` points (<= 8 x 200 B) plus a local fold; every rank ends with the identical
canonical affine result.  One process per GPU; torch.distributed supplies the collective (NCCL over
NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n, world, rank):
    """contiguous, balanced partition of range(n): sizes differ by at most one, earlier ranks larger"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedSrs:
    """The local shard [lo, hi) of a logical SRS of `n_total` bases."""

    def __init__(self, ctx, curve, group, xy_local, inf_local, n_total, world, rank, precompute=True):
        self.lo, self.hi = shard_range(n_total, world, rank)
        if len(inf_local) != self.hi - self.lo:
            raise ValueError("local shard has %d bases, expected %d" % (len(inf_local), self.hi - self.lo))
        self.n_total, self.world, self.rank = n_total, world, rank
        self.curve, self.group = curve, group
        self.srs = ctx.srs_upload(curve, group, xy_local, inf_local, precompute=precompute)

    def free(self):
        self.srs.free()


def gpu_fold(ctx, curve, group):
    """fold(points_xy[world, words], inf[world]) -> (xy, is_identity): sum of the partials in rank
    order, on the device (an MSM with unit scalars over the gathered points)."""
    def fold(xy, inf):
        srs = ctx.srs_upload(curve, group, xy, inf, precompute=False)
        try:
            ones = np.zeros((len(inf), 4), dtype=np.uint64)
            ones[:, 0] = 1
            return ctx.msm(srs, ones)
        finally:
            srs.free()
    return fold


def all_gather_partials(xy, is_inf, world, rank, device=None):
    """all_gather of one (affine point, identity flag) per rank -> (xy[world, words], inf[world])"""
    import torch
    import torch.distributed as dist
    words = xy.shape[0]
    mine = torch.zeros(words + 1, dtype=torch.int64)
    mine[:words] = torch.from_numpy(xy.view(np.int64))
    mine[words] = 1 if is_inf else 0
    if device is not None:
        mine = mine.to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    stacked = torch.stack(parts).cpu().numpy()
    return np.ascontiguousarray(stacked[:, :words]).view(np.uint64), stacked[:, words].astype(np.uint8)


def msm_sharded(local_msm, fold, scalars_local, world, rank, device=None):
    """local_msm(scalars_local) -> (xy, is_identity) on this rank's shard; returns the folded result,
    identical on every rank.  world == 1 short-circuits (no collective)."""
    xy, is_inf = local_msm(scalars_local)
    if world == 1:
        return xy, is_inf
    all_xy, all_inf = all_gather_partials(xy, is_inf, world, rank, device)
    return fold(all_xy, all_inf)
