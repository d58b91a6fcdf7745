"""Interval sharding over the GPUs of one box.

Regions of interest are independent, so rank r of N takes a contiguous slice of the region list (cut only at region
boundaries, balanced by read count: the cost driver), runs the whole per-region path on its own GPU, and the records are
gathered on rank 0.  There is NO data-path collective: the only communication is the gather of the per-shard VCF text
(kilobytes).  Rank 0 concatenates the shards in rank order -- which is the reference's emission order (target, region
ordinal, contig, event) -- and only then applies the order-dependent dedup of src/indelope.nim:604-608, whose state
spans shard boundaries.
"""
import numpy as np

from . import host


def plan_shards(rois, n):
    """[(lo, hi)] * n: contiguous region ranges with ~equal read counts"""
    m = rois.n_rois
    if m == 0:
        return [(0, 0)] * n
    reads = np.ctypeslib.as_array(rois.c.roi_n_reads, shape=(m,)).astype(np.int64)
    cum = np.cumsum(reads)
    cuts = [0]
    for r in range(1, n):
        cuts.append(int(np.searchsorted(cum, cum[-1] * r / n)))
    cuts.append(m)
    cuts = [min(max(c, cuts[i - 1] if i else 0), m) for i, c in enumerate(cuts)]
    return [(cuts[i], cuts[i + 1]) for i in range(n)]


def merge_records(shard_texts):
    """shard_texts in rank order (each without dedup) -> final record text"""
    return host.dedup_records("".join(shard_texts))


def call_sharded(rois, call_shard, group=None):
    """run `call_shard(lo, hi) -> record text (no dedup)` on this rank's slice and gather on rank 0.
    Returns the merged, dedup'ed record text on rank 0 and None elsewhere.  Works with any torch.distributed backend."""
    import torch.distributed as dist
    if not dist.is_initialized():
        return merge_records([call_shard(0, rois.n_rois)])
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = plan_shards(rois, world)[rank]
    mine = call_shard(lo, hi)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    return merge_records(gathered) if rank == 0 else None
