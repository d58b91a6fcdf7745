"""SURVEY 8(f)3 on the GPU: idl_bam_open (BGZF inflate + CRC, record boundaries, fields), idl_bam_fetch, idl_bam_sweep against (1) a BAM
parsed with struct + zlib alone, (2) the host reader of the stand-in (bamio.cpp) and (3) idl_sweep on the host reader's arrays."""
import gzip
import random
import struct
import zlib

import numpy as np
import pytest

import ctypes as C

from indelope_b200 import abi, cuda, host
from oracle import pyoracle as orc
import idl_testutil as util

pytestmark = pytest.mark.gpu

COLS = ("chrom", "start", "stop", "len", "mapq", "flag", "seq_off", "cig_off", "bases", "quals", "cigar")


def check_against_parse(raw_bam, file_bytes):
    text, refs, want, n_unplaced = util.parse_bam(raw_bam)
    b = cuda.Bam(file_bytes)
    assert b.header == text and list(zip(b.ref_names, b.ref_len)) == refs
    assert b.n_records == len(want["start"]) and b.n_unplaced == n_unplaced
    got = b.fetch()
    for k in COLS:
        assert np.array_equal(got[k], want[k]), k
    # the per-target ranges
    for c in range(len(refs)):
        lo, hi = b.ref_first[c], b.ref_first[c + 1]
        assert np.all(want["chrom"][lo:hi] == c) and (lo == 0 or want["chrom"][lo - 1] < c) and (hi == len(want["chrom"]) or want["chrom"][hi] > c)
    return b, want


def random_records(rng, refs, n, max_len=250, long_every=0):
    recs, keys = [], []
    for _ in range(n):
        c = rng.randrange(len(refs)); pos = rng.randrange(0, refs[c][1] - 1)
        keys.append((c, pos))
    keys.sort()
    for k, (c, pos) in enumerate(keys):
        l = rng.randrange(0, max_len)
        if long_every and k % long_every == long_every - 1:
            l = rng.randrange(70_000, 200_000)       # a record longer than a 64 KiB segment and than a BGZF member
        seq = "".join(rng.choice("ACGTN" if rng.random() < 0.98 else util.SEQ16) for _ in range(l))
        cig, left = [], l
        if l and rng.random() < 0.9:
            while left > 0:
                op = rng.choice("MMMMIDSNX=") if cig else rng.choice("MS")
                ln = rng.randrange(1, max(2, min(left, 120) + 1))
                if op in "MISX=":
                    ln = min(ln, left); left -= ln
                cig.append((op, ln))
        flag = rng.choice([0, 16, 99, 147, 1024, 256, 4, 2048, 512])
        recs.append(util.bam_record(c, pos, cig, seq, qual=[rng.randrange(0, 42) for _ in range(l)], mapq=rng.randrange(0, 61), flag=flag,
                                    name=("q%d" % k).encode() * rng.randrange(1, 4), tags=b"NMC\x03" if rng.random() < 0.5 else b""))
    return recs


def test_members_of_every_kind_and_records_across_boundaries():
    rng = random.Random(11)
    refs = [("chrA", 500_000), ("chrB", 80_000), ("chrEmpty", 1000), ("chrC", 3_000_000)]
    recs = [r for r in random_records(rng, refs, 6000)]
    recs += [util.bam_record(-1, -1, [], "ACGT", flag=4, name=b"unplaced%d" % k) for k in range(5)]
    raw = util.bam_bytes(refs, recs)
    kinds = [(6, zlib.Z_DEFAULT_STRATEGY), (0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_HUFFMAN_ONLY), (4, zlib.Z_RLE)]
    for block in (0xff00, 777, 30_000):
        data = util.bgzf_compress(raw, levels=kinds, block=block)
        assert gzip.decompress(data) == raw
        b, want = check_against_parse(raw, data)
        assert b.info["n_members"] == (len(raw) + block - 1) // block + 1 and b.info["inflated_bytes"] == len(raw)
        b.close()


def test_long_records_span_segments():
    rng = random.Random(12)
    refs = [("chr1", 5_000_000)]
    recs = random_records(rng, refs, 300, long_every=7)
    raw = util.bam_bytes(refs, recs)
    assert len(raw) > 3_000_000
    b, want = check_against_parse(raw, util.bgzf_compress(raw, level=1))
    b.close()


def test_bait_inside_a_record_is_not_taken_for_a_record():
    """qualities that spell out a chain of perfectly plausible records, placed so that a 64 KiB segment starts inside them: the guess takes the bait,
    the chain check (segment k's exit must be segment k+1's first record) re-walks the segment from the true offset -- the result stays exact"""
    refs = [("chr1", 1_000_000)]
    fake = b"".join(util.bam_record(0, 1000 + k, [("M", 50)], "A" * 50, name=b"fake") for k in range(40))
    head = util.bam_bytes(refs, [])
    recs, size, k = [], len(head), 0
    while size < 400_000:
        # filler up to ~2 KiB before the next segment boundary (segments are counted from the first record), then a record whose qualities hold the bait
        seg_left = 65536 - ((size - len(head)) % 65536)
        if seg_left > 6000:
            r = util.bam_record(0, 10 + k, [("M", 100)], "ACGT" * 25, name=b"fill%d" % k)
        else:
            qual = bytes([20] * (seg_left + 100)) + fake + bytes([20] * 500)
            l = len(qual)
            r = util.bam_record(0, 10 + k, [("M", l)], "C" * l, qual=qual, name=b"bait%d" % k)
        recs.append(r); size += len(r); k += 1
    raw = util.bam_bytes(refs, recs)
    b, want = check_against_parse(raw, util.bgzf_compress(raw, level=6))
    assert b.info["boundary_fixups"] >= 3
    b.close()


def test_empty_and_header_only_files():
    refs = [("chr1", 1000), ("chr2", 2000)]
    raw = util.bam_bytes(refs, [])
    b, want = check_against_parse(raw, util.bgzf_compress(raw))
    assert b.n_records == 0 and b.ref_first == [0, 0, 0]
    s = b.sweep(0)
    assert len(s["roi_start"]) == 0
    b.close()
    raw = util.bam_bytes(refs, [util.bam_record(-1, -1, [], "AC", flag=4)])      # only unplaced records
    b, want = check_against_parse(raw, util.bgzf_compress(raw))
    assert b.n_records == 0 and b.n_unplaced == 1
    b.close()
    big = [("contig%d" % k, 1000 + k) for k in range(40_000)]                      # a header of more than 1 MiB
    raw = util.bam_bytes(big, [util.bam_record(39_999, 5, [("M", 4)], "ACGT")])
    b, want = check_against_parse(raw, util.bgzf_compress(raw, level=1))
    assert b.ref_first[39_999] == 0 and b.ref_first[40_000] == 1
    b.close()


def test_bad_files_are_refused():
    rng = random.Random(13)
    refs = [("chr1", 100_000)]
    recs = random_records(rng, refs, 2000)
    raw = util.bam_bytes(refs, recs)
    data = util.bgzf_compress(raw, level=6)
    def refused(d, match):
        with pytest.raises(cuda.IdlError, match=match):
            cuda.Bam(d)
    refused(b"not a bam at all, just text " * 4, "not a BGZF")
    refused(data[:len(data) // 2], "truncated BGZF")
    bad = bytearray(data); bad[5000] ^= 0x10
    refused(bytes(bad), "failed to inflate")                                       # a flipped bit in the deflate data: decoder error or CRC mismatch
    member = util.bgzf_member(raw[:30_000])
    wrong_crc = member[:-8] + struct.pack("<I", zlib.crc32(raw[:30_000]) ^ 1) + member[-4:]
    refused(wrong_crc + util.BGZF_EOF, "CRC mismatch")
    refused(util.bgzf_compress(b"BAX\1" + raw[4:]), "not a BAM")
    refused(util.bgzf_compress(raw[:len(raw) - 10]), "truncated BAM record")
    refused(util.bgzf_compress(util.bam_bytes(refs, [recs[-1], recs[0]])), "not coordinate sorted")
    refused(util.bgzf_compress(util.bam_bytes(refs, [util.bam_record(3, 5, [("M", 4)], "ACGT")])), "unknown reference id")
    refused(util.bgzf_compress(util.bam_bytes(refs, [util.bam_record(-1, -1, [], "AC", flag=4), recs[0]])), "unplaced records are only supported as the tail")
    broken = bytearray(util.bam_record(0, 5, [("M", 4)], "ACGT")); broken[20:24] = struct.pack("<i", 4000)   # l_seq larger than the record
    refused(util.bgzf_compress(util.bam_bytes(refs, [bytes(broken)])), "malformed BAM record")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("gbam")
    ds = util.small_dataset("pr1", chrom_len=400_000, n_events=80, max_indel=40, n_chroms=3, n_base_rate=0.0005, dup_fraction=0.02)
    fa, bam = str(d / "ref.fa"), str(d / "reads.bam")
    ds.write_fasta(fa)
    ds.write_bam(bam, level=6)
    return ds, fa, bam


def test_same_records_and_regions_as_the_host_reader(files):
    """the stand-in's own BAM through both readers: identical records; idl_bam_sweep == idl_sweep on the host arrays == the host sweep"""
    ds, fa, bam = files
    data = open(bam, "rb").read()
    b, want = check_against_parse(gzip.decompress(data), data)
    full = host.Dataset.load(fa, bam, threads=2)
    rois = full.sweep(min_reads=5)
    a = rois.arrays()
    got = b.fetch(what=cuda.BAM_SEQ)
    for k in ("start", "stop", "len", "mapq", "flag"):
        assert np.array_equal(got[k], a[k]), k
    assert np.array_equal(got["bases"], a["bases"][:len(got["bases"])]) and np.array_equal(got["quals"], a["quals"][:len(got["quals"])])
    n_regions = 0
    for c in range(b.n_ref):
        cr = full.chrom_reads(c)
        assert cr["first_read"] == b.ref_first[c] and len(cr["start"]) == b.ref_first[c + 1] - b.ref_first[c]
        s_dev = b.sweep(c, min_event_support=3, min_read_coverage=5, evidence=True)
        s_host = cuda.sweep(cr["chrom_len"], cr["start"], cr["stop"], cr["flag"], cr["cigar"], cr["cig_off"], min_event_support=3, min_read_coverage=5, evidence=True)
        for k in ("roi_start", "roi_end", "roi_n_reads", "roi_read_begin", "evidence"):
            assert np.array_equal(s_dev[k], s_host[k]), k
        assert np.array_equal(s_dev["read_idx"], s_host["read_idx"] + cr["first_read"])
        sel = a["roi_chrom"] == c
        assert np.array_equal(s_dev["roi_start"], a["roi_start"][sel]) and np.array_equal(s_dev["roi_end"], a["roi_stop"][sel])
        n_regions += len(s_dev["roi_start"])
    assert n_regions == len(a["roi_start"]) > 50
    # a subset, in an order of the caller's choosing
    idx = np.array([5, 0, b.n_records - 1, 17, 17], dtype=np.int64)
    sub = b.fetch(idx)
    allr = b.fetch()
    for j, i in enumerate(idx):
        assert sub["start"][j] == allr["start"][i] and sub["flag"][j] == allr["flag"][i]
        assert np.array_equal(sub["bases"][sub["seq_off"][j]:sub["seq_off"][j + 1]], allr["bases"][allr["seq_off"][i]:allr["seq_off"][i + 1]])
        assert np.array_equal(sub["quals"][sub["seq_off"][j]:sub["seq_off"][j + 1]], allr["quals"][allr["seq_off"][i]:allr["seq_off"][i + 1]])
        assert np.array_equal(sub["cigar"][int(sub["cig_off"][j]):int(sub["cig_off"][j + 1])], allr["cigar"][int(allr["cig_off"][i]):int(allr["cig_off"][i + 1])])
    b.close()


def _batch_bytes(b):
    """the filled part of an idl_batch as bytes per array (+ the summary)"""
    c = b.contents
    out = dict(sizes=(c.n_regions, c.n_reads, c.n_seq_bases, c.n_ref_bases), summary=(c.summary_valid, c.max_trim_len, c.max_ref_len, c.max_region_reads, c.n_small_regions))
    out["region"] = C.string_at(c.region, c.n_regions * C.sizeof(abi.Region)); out["read"] = C.string_at(c.read, c.n_reads * C.sizeof(abi.Read))
    out["seq2"] = C.string_at(c.seq2, c.n_seq_bases // 4 + 16); out["seqn"] = C.string_at(c.seqn, c.n_seq_bases // 8 + 16)
    out["ref2"] = C.string_at(c.ref2, c.n_ref_bases // 4 + 16); out["refn"] = C.string_at(c.refn, c.n_ref_bases // 8 + 16)
    return out


def _compare_device_and_host_pack(fa, bam, min_reads=5, expect_flags=None):
    data = open(bam, "rb").read()
    b = cuda.Bam(data)
    full = host.Dataset.load(fa, bam, threads=2)
    rois = full.sweep(min_reads=min_reads)
    a = rois.arrays()
    n = rois.n_rois
    assert n > 0
    for c in range(b.n_ref):
        b.set_reference(c, a["chrom_seqs"][c])
    params = abi.default_params(min_reads=min_reads, min_ctg_len=73, min_event_len=5)
    ctx = cuda.Context(0, params)
    nr, sb, rb = rois.pack_size(0, n, params)
    hb = ctx.batch_alloc(n + 8, nr + 8, sb + 64, rb + 64); db = ctx.batch_alloc(n + 8, nr + 8, sb + 64, rb + 64)
    rois.pack(0, n, params, hb)
    ctx.bam_pack(b, a["roi_chrom"], a["roi_start"], a["roi_stop"], a["roi_n_reads"], a["read_idx"], db)
    H, D = _batch_bytes(hb), _batch_bytes(db)
    for k in ("sizes", "summary", "region", "read", "seq2", "seqn", "ref2", "refn"):
        assert H[k] == D[k], k
    if expect_flags is not None:
        flags = [hb.contents.region[k].flags for k in range(n)]
        assert expect_flags(flags), flags
    # ... and through the calling chain: idl_bam_submit == idl_submit of the host batch == the oracle
    w1, w2 = host.VcfWriter(), host.VcfWriter()
    t = ctx.submit(hb); v1, _ = w1.records(rois, 0, params, ctx.wait(t)); ctx.release(t)
    t = ctx.bam_submit(b, a["roi_chrom"], a["roi_start"], a["roi_stop"], a["roi_n_reads"], a["read_idx"]); v2, _ = w2.records(rois, 0, params, ctx.wait(t)); ctx.release(t)
    assert v1 == v2
    ctx.batch_free(hb); ctx.batch_free(db); ctx.close(); b.close()
    return rois, v2


def test_device_built_batch_equals_the_host_pack(files):
    """idl_bam_pack / idl_bam_submit: quality trim, windows, records and the 2-bit pools built on the device == idlh_pack on the host reader's arrays,
    byte for byte; the VCF through idl_bam_submit == the oracle's"""
    ds, fa, bam = files
    rois, vcf = _compare_device_and_host_pack(fa, bam)
    _, ovcf, cnt = orc.call(rois.arrays(), dump_level=0, min_reads=5, min_ctg_len=73, min_event_len=5, use_ref_ksw2=orc.have_ref())
    assert vcf == ovcf and cnt["variants"] >= 3


def test_device_pack_of_odd_letters_trims_and_long_reads(tmp_path):
    """hand-made files: lower-case, N and IUPAC letters in the reference window and IUPAC codes in reads (folded + IDL_RF_ALPHABET), reads whose
    qualities trim to nothing / to one base / at both ends, an empty read, a read longer than max_read_len (IDL_RF_READ_TOO_LONG), low MAPQ"""
    rng = random.Random(21)
    L = 6000
    seq = [rng.choice("ACGT") for _ in range(L)]
    for i in range(1000, 1100):
        seq[i] = seq[i].lower()
    seq[1200] = "N"; seq[1201] = "n"; seq[1250] = "R"; seq[4100] = "y"
    ref = "".join(seq)
    fa = tmp_path / "r.fa"
    fa.write_text(">c1\n" + "\n".join(ref[i:i + 60] for i in range(0, L, 60)) + "\n>c2\n" + "ACGT" * 50 + "\n")
    recs = []
    def reads_at(center, n, odd=None):
        out = []
        for k in range(n):
            start = center - 100 + 3 * k
            left = center - start                                        # every read of the cluster carries the insertion at `center`
            s = ref[start:center].upper().replace("N", "A").replace("R", "A").replace("Y", "C") + "TTGCA" + ref[center:start + 140].upper().replace("Y", "C")
            q = [30] * len(s)
            cig = [("M", left), ("I", 5), ("M", 140 - left)]
            mapq = 60
            if odd == "letters" and k % 4 == 0:
                s = s[:20] + "MRN" + s[23:]
            if odd == "quals":
                if k == 0: q = [2] * len(s)
                if k == 1: q = [2] * 50 + [30] + [2] * (len(s) - 51)
                if k == 2: q = [2] * 7 + [30] * (len(s) - 20) + [5] * 13
                if k == 3: mapq = 3
            out.append((start, util.bam_record(0, start, cig, s, qual=q, mapq=mapq, name=b"r%d_%d" % (center, k))))
        if odd == "long":
            s = "".join(rng.choice("ACGT") for _ in range(700))
            out.append((center - 50, util.bam_record(0, center - 50, [("M", 300), ("I", 100), ("M", 300)], s, name=b"long")))
        if odd == "quals":
            out.append((center - 10, util.bam_record(0, center - 10, [], "", name=b"empty")))
        return out
    allr = reads_at(1150, 12, "letters") + reads_at(2500, 12, "quals") + reads_at(4000, 12, "long") + reads_at(5200, 10)
    allr.sort(key=lambda t: t[0])
    raw = util.bam_bytes([("c1", L), ("c2", 200)], [r for _, r in allr])
    bam = tmp_path / "r.bam"
    bam.write_bytes(util.bgzf_compress(raw, level=6))
    def flags_ok(flags):
        return any(f & 1 for f in flags) and any(f & 2 for f in flags) and any(f == 0 for f in flags)
    _compare_device_and_host_pack(str(fa), str(bam), min_reads=5, expect_flags=flags_ok)
