"""Interval sharding with the CUDA engine on more than one GPU: every rank owns one device, runs the whole per-region path on its
slice of ONE dataset through the C ABI (records without the order-dependent dedup), rank 0 gathers the shards in rank order,
applies the dedup of src/indelope.nim:604-608 across the shard boundaries and the merged VCF must equal the oracle's single
sweep byte for byte.  No data-path collective exists; torch.distributed only carries the gather of the record text.
Skipped on a one-GPU box (run it with `gpurun --gpus 2`)."""
import os
import socket
import sys

import pytest

import idl_testutil as util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(n_chroms=3, chrom_len=400_000, n_events=70, max_indel=40, tr_fraction=0.4, tr_max_unit=4, seed=77, locus_only=0)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from indelope_b200 import api, shard
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ds = util.small_dataset("pr1", **CFG)
    rois = ds.sweep(min_reads=5)
    caller = api.Caller(rank, min_reads=5, min_ctg_len=73, min_event_len=5)
    try:
        merged = shard.call_sharded(rois, lambda lo, hi: caller.call(rois, lo=lo, hi=hi, dedup=False, max_reads=60_000)[0])
    finally:
        caller.close()
    if rank == 0:
        with open(out_path, "w") as f:
            f.write(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_one_dataset_merged_vcf_matches_the_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle import pyoracle as orc
    world = min(4, torch.cuda.device_count())
    out = str(tmp_path / "merged.vcf")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ds = util.small_dataset("pr1", **CFG)
    rois = ds.sweep(min_reads=5)
    raw, vcf, cnt = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, dump_level=32, use_ref_ksw2=orc.have_ref())
    _, dedup_vcf, _ = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, dump_level=0, use_ref_ksw2=orc.have_ref())
    assert cnt["variants"] > 20 and len(set(l.split("\t")[0] for l in dedup_vcf.splitlines())) == 3  # records on all three contigs
    assert open(out).read() == dedup_vcf
