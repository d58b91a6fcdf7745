"""Multi-GPU host logic on CPU: two gloo ranks shard the regions by interval, each runs the per-region path on its
slice (the CPU oracle stands in for the GPU engine here: this test is about sharding, gathering, ordering and the
order-dependent dedup across shard boundaries), rank 0 merges.  The merged VCF must equal the single-process one."""
import os
import socket
import sys

import numpy as np
import pytest

import idl_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from indelope_b200 import shard
    from oracle import pyoracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = util.small_dataset("pr1", chrom_len=300_000, n_events=60, max_indel=40, seed=21)
    rois = ds.sweep(min_reads=5)
    a = rois.arrays()

    def call_shard(lo, hi):
        sub = dict(a)
        for k in ("roi_chrom", "roi_start", "roi_stop", "roi_read_begin", "roi_n_reads"):
            sub[k] = a[k][lo:hi]
        # raw (pre-dedup) records of this slice: dedup state spans shard boundaries, so it is applied after the merge only
        _, vcf, _ = orc.call(sub, min_reads=5, min_ctg_len=73, min_event_len=5, dump_level=32)
        return vcf
    merged = shard.call_sharded(rois, call_shard)
    if rank == 0:
        with open(out_path, "w") as f:
            f.write(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_interval_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import pyoracle as orc
    out = str(tmp_path / "merged.vcf")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ds = util.small_dataset("pr1", chrom_len=300_000, n_events=60, max_indel=40, seed=21)
    rois = ds.sweep(min_reads=5)
    _, vcf, cnt = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, dump_level=0)
    assert cnt["variants"] > 5
    assert open(out).read() == vcf


def test_plan_shards_is_a_balanced_partition():
    from indelope_b200 import shard
    ds = util.small_dataset("pr1", chrom_len=300_000, n_events=60, seed=22)
    rois = ds.sweep(min_reads=5)
    reads = rois.arrays()["roi_n_reads"]
    for n in (1, 2, 3, 4, 8):
        plan = shard.plan_shards(rois, n)
        assert plan[0][0] == 0 and plan[-1][1] == rois.n_rois and all(plan[i][1] == plan[i + 1][0] for i in range(n - 1))
        loads = [int(reads[a:b].sum()) for a, b in plan]
        assert max(loads) <= sum(loads) / n + int(reads.max()) + 1


def test_dedup_of_merged_records_is_order_dependent():
    from indelope_b200 import host
    l = lambda pos, ref="A", alt="AT": "chr1\t%d\t.\t%s\t%s\t1.00\tPASS\tx\tGT:GQ:GL\t0/1:1:1,1,1\n" % (pos, ref, alt)
    # same(last_var) and same(last_var2) are dropped; a third-to-last duplicate is kept (src/indelope.nim:604-608)
    text = l(10) + l(10) + l(20) + l(10) + l(30) + l(40) + l(10)
    assert host.dedup_records(text) == l(10) + l(20) + l(30) + l(40) + l(10)
