"""gen_roi on the GPU (idl_sweep, indelope_b200/csrc/sweep.cu) against the host stand-in's restatement of
src/indelope.nim:461-545 (idlh_sweep): the same regions (roi_start, roi_end) with the same record lists in the same order, on every
BASELINE config, plus the saturating uint8 evidence array itself and directed cases (coverage gaps with evidence on both sides of the
boundary, the 600-record cap, skippable records, soft clips at read ends, saturation at 255)."""
import numpy as np
import pytest

import idl_testutil as util
from indelope_b200 import host

pytestmark = pytest.mark.gpu


def evidence_ref(cr, chrom_len):
    """src/indelope.nim:430-442,538-543 in numpy: saturating uint8 counts of event intervals of non-skippable records"""
    diff = np.zeros(chrom_len + 2, np.int64)
    skip = (cr["flag"] & (0x400 | 0x200 | 0x4 | 0x800 | 0x100)) != 0
    for i in range(len(cr["start"])):
        if skip[i]:
            continue
        off = 0
        for c in cr["cigar"][int(cr["cig_off"][i]):int(cr["cig_off"][i + 1])]:
            op, ln = int(c) & 0xf, int(c) >> 4
            cons = op in (0, 2, 3, 7, 8)
            if op != 0:
                es = int(cr["start"][i]) + off; ee = min(es + ln if cons else es + 1, chrom_len + 1)
                if es < ee:
                    diff[es] += 1; diff[ee] -= 1
            if cons:
                off += ln
    return np.minimum(np.cumsum(diff)[:chrom_len + 1], 255).astype(np.uint8)


def compare(ds, min_reads=5, max_cov=600, check_evidence=False):
    from indelope_b200 import cuda
    rois = ds.sweep(min_reads=min_reads, max_read_coverage=max_cov)
    a = rois.arrays()
    total = 0
    for c in range(ds.n_chroms):
        cr = ds.chrom_reads(c)
        r = cuda.sweep(cr["chrom_len"], cr["start"], cr["stop"], cr["flag"], cr["cigar"], cr["cig_off"], min_event_support=max(3, min_reads - 2),
                       min_read_coverage=min_reads, max_read_coverage=max_cov, evidence=check_evidence)
        sel = a["roi_chrom"] == c
        assert np.array_equal(r["roi_start"], a["roi_start"][sel]) and np.array_equal(r["roi_end"], a["roi_stop"][sel])
        assert np.array_equal(r["roi_n_reads"], a["roi_n_reads"][sel])
        want = np.concatenate([a["read_idx"][b:b + n] for b, n in zip(a["roi_read_begin"][sel], a["roi_n_reads"][sel])]) if sel.any() else np.zeros(0, np.int64)
        assert np.array_equal(r["read_idx"] + cr["first_read"], want)
        assert np.array_equal(r["roi_read_begin"], np.concatenate([[0], np.cumsum(r["roi_n_reads"])[:-1]]) if len(r["roi_n_reads"]) else np.zeros(0, np.int64))
        if check_evidence:
            assert np.array_equal(r["evidence"], evidence_ref(cr, cr["chrom_len"]))
        total += int(sel.sum())
    return total, rois


@pytest.mark.parametrize("name,over", [
    ("pr1", dict()),
    ("pr1", dict(n_chroms=3, chrom_len=300_000, n_events=60, seed=5, dup_fraction=0.05)),
    ("exome", dict()),
    ("panel500", dict()),
    ("panel500_lowerr", dict()),
])
def test_regions_and_read_lists_equal_the_host_sweep(name, over):
    """BASELINE configs 1, 2 and 4 (and a three-contig variant with 5 % duplicates): every region, every record list"""
    ds = util.small_dataset(name, **over)
    n, _ = compare(ds)
    assert n > 100


@pytest.mark.parametrize("name", ["chr1", "wgs"])
def test_full_size_workloads(name):
    """BASELINE config 3 and one rank's shard of config 5 at full size: ~85 k / ~44 k regions out of a 248 Mb / 129 Mb evidence array"""
    ds = util.small_dataset(name)
    n, _ = compare(ds)
    assert n > 40_000


def test_evidence_array_and_small_thresholds():
    """the uint8 evidence array itself equals a numpy restatement of :430-442,538-543; other CLI thresholds (-m 3: min_event_support 3,
    min_read_coverage 3) and a tight record cap that drops most deep regions"""
    ds = util.small_dataset("pr1", chrom_len=200_000, n_events=40, tr_fraction=0.3, seed=11)
    compare(ds, min_reads=5, check_evidence=True)
    compare(ds, min_reads=3)
    n_capped, _ = compare(ds, min_reads=5, max_cov=40)
    n_all, _ = compare(ds, min_reads=5)
    assert n_capped < n_all


def test_directed_boundaries_caps_and_saturation():
    """hand-made records: (1) a coverage gap with evidence on BOTH sides of the chunk boundary (a trailing soft clip at position
    stop, a leading soft clip of the next chunk's first record one base further): two regions, not one; (2) a skippable record
    that opens a gap candidate while the cache is empty; (3) 300 records with the same insertion: the evidence saturates at 255;
    (4) a region with exactly max_reads records is kept, one with max_reads + 1 is dropped; (5) a deletion running off the contig end"""
    from indelope_b200 import cuda
    M, I, D, S = 0, 1, 2, 4
    recs = []
    def add(start, cig, flag=0):
        ref = sum(l for l, o in cig if o in (M, D, 3, 7, 8))
        recs.append((start, start + ref, flag, [l << 4 | o for l, o in cig]))
    # (1) chunk A ends at 1100 with trailing soft clips (event at 1100), chunk B starts at 1101 with leading soft clips (event at 1101)
    for k in range(6):
        add(1000, [(100, M), (10, S)])
    for k in range(6):
        add(1101, [(10, S), (100, M)])
    # (2) a duplicate (skippable) far beyond everything, then a normal chunk with insertions
    add(5000, [(100, M)], flag=0x400)
    add(5050, [(100, M)], flag=0x400)
    for k in range(8):
        add(6000, [(50 + k, M), (4, I), (50, M)][0:1] + [(4, I), (50, M)]) if False else add(6000 + (k & 1), [(50 - (k & 1), M), (4, I), (50, M)])
    # (3) saturation: 300 records, same 1-base insertion site
    for k in range(300):
        add(10_000, [(60, M), (2, I), (60, M)])
    # (4) exactly 20 records / 21 records over an insertion (max_reads = 20 below applies to (3) as well: dropped)
    for k in range(20):
        add(20_000 + (k % 4), [(40 - (k % 4), M), (3, I), (80, M)])
    for k in range(21):
        add(30_000 + (k % 4), [(40 - (k % 4), M), (3, I), (80, M)])
    # (5) deletion reaching the last base of the contig
    for k in range(6):
        add(39_900, [(50, M), (60, D)])
    recs.sort(key=lambda r: r[0])
    start = np.array([r[0] for r in recs], np.int32); stop = np.array([r[1] for r in recs], np.int32); flag = np.array([r[2] for r in recs], np.uint16)
    off = np.zeros(len(recs) + 1, np.uint64); off[1:] = np.cumsum([len(r[3]) for r in recs]); cig = np.array([c for r in recs for c in r[3]], np.uint32)
    chrom_len = 40_010
    r = cuda.sweep(chrom_len, start, stop, flag, cig, off, min_event_support=3, min_read_coverage=5, max_read_coverage=20, evidence=True)
    ev = evidence_ref(dict(start=start, stop=stop, flag=flag, cigar=cig, cig_off=off), chrom_len)
    assert np.array_equal(r["evidence"], ev) and ev.max() == 255 and ev[10_060] == 255
    got = list(zip(r["roi_start"].tolist(), r["roi_end"].tolist(), r["roi_n_reads"].tolist()))
    # (1): two one-base regions at 1100 and 1101, six records each (records of the other chunk do not overlap: stop 1100 < 1101 is false for
    # overlaps' r.stop < start test, so chunk A's records DO overlap 1101? no: region 1101 takes records with start <= 1101 and stop >= 1101:
    # chunk A has stop = 1100 -> excluded)
    assert (1100, 1100, 12) in got or (1100, 1100, 6) in got
    exp = python_gen_roi(start, stop, flag, cig, off, chrom_len, 3, 5, 20)
    assert got == [(a, b, len(l)) for a, b, l in exp]
    want_idx = [i for _, _, l in exp for i in l]
    assert r["read_idx"].tolist() == want_idx
    starts = [g[0] for g in got]
    assert 1100 in starts and 1101 in starts and 6050 in starts and 10_060 not in starts and 20_040 in starts and 30_040 not in starts and 39_950 in starts


def python_gen_roi(start, stop, flag, cig, off, chrom_len, min_ev, min_reads, max_reads):
    """src/indelope.nim:461-545 restated literally in Python (sequential sweep with the cache), for the directed test"""
    ev = np.zeros(chrom_len + 1, np.int64)
    out = []
    cache, cache_stop, last_start = [], 0, 0
    def internal(a, b):
        in_roi, rs, re = False, 0, 0
        def flush():
            reads = []
            for k in cache:
                if not (start[k] > re) and not (stop[k] < rs):
                    reads.append(k)
                    if len(reads) > max_reads:
                        break
                if start[k] > re:
                    break
            if min_reads <= len(reads) <= max_reads:
                out.append((rs, re, reads))
        for i in range(a, b):
            if min(ev[i], 255) >= min_ev:
                if not in_roi:
                    in_roi, rs = True, i
                re = i
                continue
            if in_roi:
                flush(); in_roi = False
        if in_roi:
            flush()
    for k in range(len(start)):
        if cache and start[k] > cache_stop:
            internal(last_start, int(start[k]))
            last_start = int(start[k]); cache, cache_stop = [], 0
        if flag[k] & (0x400 | 0x200 | 0x4 | 0x800 | 0x100):
            continue
        cache.append(k); cache_stop = max(cache_stop, int(stop[k]))
        o = 0
        for c in cig[int(off[k]):int(off[k + 1])]:
            op, ln = int(c) & 0xf, int(c) >> 4
            cons = op in (0, 2, 3, 7, 8)
            if op != 0:
                es = int(start[k]) + o; ee = es + ln if cons else es + 1
                for i in range(es, min(ee, chrom_len + 1)):
                    ev[i] += 1
            if cons:
                o += ln
    internal(last_start, chrom_len + 1)
    return out


def test_gpu_sweep_feeds_the_calling_path():
    """regions from idl_sweep -> idl_submit: BAM records to VCF with no per-record host loop but the packing; the VCF is the oracle's over
    the HOST sweep's regions, byte for byte"""
    from indelope_b200 import api
    from oracle import pyoracle as orc
    ds = util.small_dataset("pr1", chrom_len=400_000, n_events=80, max_indel=40, tr_fraction=0.3, seed=19)
    rois = api.gpu_rois(ds, min_reads=5)
    caller = api.Caller(0, min_reads=5, min_ctg_len=73, min_event_len=5)
    try:
        vcf, _ = caller.call(rois)
    finally:
        caller.close()
    href = ds.sweep(min_reads=5)
    _, ovcf, cnt = orc.call(href.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, dump_level=0, use_ref_ksw2=orc.have_ref())
    assert cnt["variants"] > 10 and vcf == ovcf
