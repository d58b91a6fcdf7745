"""CPU-only checks of the host side and of the boundary: the C-ABI libraries load and export every declared symbol,
struct layouts match the headers, packing round-trips, the sweep restates gen_roi, and the product-side VCF writer
agrees with the oracle's on oracle-produced integers (no GPU compute here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from indelope_b200 import abi, build, host
from oracle import pyoracle as orc
import idl_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(idlh?_[a-z0-9_]+)\s*\(", txt)))


def test_cuda_library_exports_every_declared_symbol():
    build.build_cuda()
    lib = C.CDLL(build.CUDA_LIB)  # loads without a GPU: the runtime is linked statically, no device call at load time
    names = declared("indelope_cuda.h")
    assert "idl_submit" in names and "idl_ksw2_batch" in names and len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    lib.idl_device_count.restype = C.c_int
    p = abi.Params()
    lib.idl_default_params(C.byref(p))
    q = abi.default_params()
    assert bytes(p) == bytes(q)


def test_no_cpu_fallback_without_device():
    if util.has_gpu():
        pytest.skip("GPU present")
    lib = C.CDLL(build.build_cuda())
    h = C.c_void_p()
    p = abi.default_params()
    rc = lib.idl_create(0, C.byref(p), C.byref(h))
    assert rc == -1 and not h.value  # IDL_E_NO_DEVICE: the product fails loudly, it never computes on the CPU


def test_host_library_exports_every_declared_symbol():
    lib = host.lib()
    for n in declared("indelope_host.h"):
        assert hasattr(lib, n), n


def test_struct_sizes_match_header():
    assert C.sizeof(abi.Region) == 48 and C.sizeof(abi.Read) == 24
    assert C.sizeof(abi.RegionResult) == 16 and C.sizeof(abi.ContigResult) == 24
    assert C.sizeof(abi.AlnResult) == 72 and C.sizeof(abi.EventResult) == 128


def test_trim_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(0, 40))
        q = rng.choice([2, 30], size=n, p=[0.4, 0.6]).astype(np.uint8)
        assert host.trim(q) == orc.trim(q)


def test_pack_roundtrip_and_read_records():
    ds = util.small_dataset("pr1", chrom_len=120_000, n_events=20, n_base_rate=0.01)
    rois = ds.sweep(min_reads=5)
    assert rois.n_rois > 10
    a = rois.arrays()
    p = abi.default_params(min_reads=5)
    nr, sb, rb = rois.pack_size(0, rois.n_rois, p)
    b = host.host_batch(rois.n_rois, nr, sb, rb)
    rois.pack(0, rois.n_rois, p, b)
    bt = b.contents
    assert bt.n_regions == rois.n_rois and bt.n_reads == nr
    k = 0
    for g in range(rois.n_rois):
        reg = bt.region[g]
        assert reg.n_reads == a["roi_n_reads"][g] and reg.read_begin == k
        idx = a["read_idx"][a["roi_read_begin"][g]:a["roi_read_begin"][g] + a["roi_n_reads"][g]]
        for j, i in enumerate(idx):
            rd = bt.read[k + j]
            seq = bytes(a["bases"][a["seq_off"][i]:a["seq_off"][i] + a["len"][i]]).decode()
            assert host.unpack(bt.seq2, bt.seqn, rd.seq_off, rd.len) == seq
            ta, tl = orc.trim(a["quals"][a["seq_off"][i]:a["seq_off"][i] + a["len"][i]])
            assert (rd.trim_a, rd.trim_len, rd.min_overlap) == (ta, tl, int(0.88 * tl))
            assert rd.seq_off % 64 == 0 and rd.mapq == a["mapq"][i] and rd.start == a["start"][i] and rd.stop == a["stop"][i]
        ref = bytes(a["chrom_seqs"][0][reg.ref_start:reg.ref_start + reg.ref_len]).decode()
        assert host.unpack(bt.ref2, bt.refn, reg.ref_off, reg.ref_len) == ref
        k += reg.n_reads
    host.host_batch_free(b)


def _pack(rois, p, threads=0):
    host.set_threads(threads)
    try:
        nr, sb, rb = rois.pack_size(0, rois.n_rois, p)
        b = host.host_batch(rois.n_rois, nr, sb, rb)
        rois.pack(0, rois.n_rois, p, b)
    finally:
        host.set_threads(0)
    return b, (nr, sb, rb)


def _pool_bytes(bt):
    return tuple(C.string_at(ptr, n) for ptr, n in ((bt.region, bt.n_regions * 48), (bt.read, bt.n_reads * 24), (bt.seq2, bt.n_seq_bases // 4),
                                                    (bt.seqn, bt.n_seq_bases // 8), (bt.ref2, bt.n_ref_bases // 4), (bt.refn, bt.n_ref_bases // 8)))


def test_pack_is_thread_count_invariant_and_fills_the_summary():
    """idlh_pack on 1 thread and on 7 threads writes the same bytes; the batch summary idl_submit relies on equals a scan"""
    ds = util.small_dataset("pr1", chrom_len=200_000, n_events=40, n_base_rate=0.005)
    rois = ds.sweep(min_reads=5)
    p = abi.default_params(min_reads=5)
    b1, s1 = _pack(rois, p, 1)
    b7, s7 = _pack(rois, p, 7)
    assert s1 == s7 and _pool_bytes(b1.contents) == _pool_bytes(b7.contents)
    bt = b7.contents
    assert bt.summary_valid == 1
    assert bt.max_trim_len == max(bt.read[i].trim_len for i in range(bt.n_reads))
    assert bt.max_ref_len == max(bt.region[i].ref_len for i in range(bt.n_regions))
    assert bt.max_region_reads == max(bt.region[i].n_reads for i in range(bt.n_regions))
    assert bt.n_small_regions == sum(bt.region[i].n_reads <= 126 for i in range(bt.n_regions))
    assert all(bt.region[i].flags == 0 for i in range(bt.n_regions))  # ACGTN only: nothing was folded
    host.host_batch_free(b1); host.host_batch_free(b7)


def test_pack_flags_folded_bytes_and_overlong_reads():
    """bytes outside {A,C,G,T,N} (soft-masked lower case, IUPAC codes, '=') are folded AND reported per region; a read longer
    than max_read_len is packed empty and flagged instead of failing the batch (src/contig.nim:93 compares raw characters)"""
    rng = np.random.default_rng(3)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 4000))
    sets = []
    for case in range(4):
        reads = [dict(start=1000 + i, seq=ref[1000 + i:1150 + i]) for i in range(0, 60, 5)]
        if case == 1:
            reads[3]["seq"] = reads[3]["seq"][:40] + "r" + reads[3]["seq"][41:]        # IUPAC, lower case -> N
        if case == 2:
            reads[5]["seq"] = reads[5]["seq"][:17].lower() + reads[5]["seq"][17:]      # soft-masked -> upper case
        if case == 3:
            reads[2] = dict(start=1010, seq=ref[1010:1010 + 700])                      # longer than max_read_len = 512
        _, arrays = util.rois_from_reads(reads, ref)
        sets.append(arrays)
    arrays = util.merge_rois(sets)
    rois = host.Rois(arrays=arrays)
    p = abi.default_params(min_reads=3)
    b, _ = _pack(rois, p, 2)
    bt = b.contents
    assert [bt.region[i].flags for i in range(4)] == [0, 1, 1, 2]
    rd = bt.read[bt.region[1].read_begin + 3]
    assert host.unpack(bt.seq2, bt.seqn, rd.seq_off, rd.len)[38:43] == reads_text(arrays, 1, 3)[38:40] + "N" + reads_text(arrays, 1, 3)[41:43]
    rd = bt.read[bt.region[2].read_begin + 5]
    assert host.unpack(bt.seq2, bt.seqn, rd.seq_off, rd.len) == reads_text(arrays, 2, 5).upper()
    rd = bt.read[bt.region[3].read_begin + 2]
    assert rd.len == 0 and rd.trim_len == 0
    host.host_batch_free(b)
    # a soft-masked reference window is folded and flagged as well
    _, arrays = util.rois_from_reads([dict(start=1000 + i, seq=ref[1000 + i:1150 + i]) for i in range(0, 60, 5)], ref[:1100] + ref[1100:1200].lower() + ref[1200:])
    rois = host.Rois(arrays=arrays)
    b, _ = _pack(rois, p, 1)
    assert b.contents.region[0].flags == 1
    host.host_batch_free(b)


def reads_text(a, region, j):
    i = a["read_idx"][a["roi_read_begin"][region] + j]
    return bytes(a["bases"][a["seq_off"][i]:a["seq_off"][i] + a["len"][i]]).decode()


def test_int_088_equals_integer_form():
    # int(0.88*len) == 88*len div 100 for every len <= 2000
    for n in range(0, 2001):
        assert int(0.88 * n) == (88 * n) // 100


def test_sweep_regions_cover_planted_events():
    ds = util.small_dataset("pr1", chrom_len=200_000, n_events=30)
    rois = ds.sweep(min_reads=5)
    a = rois.arrays()
    truth = ds.truth()
    hit = 0
    for ev in truth:
        pos = ev[1]
        hit += bool(np.any((a["roi_start"] <= pos + ev[3] + 2) & (a["roi_stop"] >= pos - 2)))
    assert hit >= 0.9 * len(truth)
    assert np.all(a["roi_n_reads"] >= 5) and np.all(a["roi_n_reads"] <= 600)
    # reads of a region are in BAM order and overlap it (src/indelope.nim:449-452,480-484)
    for g in range(rois.n_rois):
        idx = a["read_idx"][a["roi_read_begin"][g]:a["roi_read_begin"][g] + a["roi_n_reads"][g]]
        assert np.all(np.diff(idx) > 0)
        assert np.all(a["start"][idx] <= a["roi_stop"][g]) and np.all(a["stop"][idx] >= a["roi_start"][g])


def test_header_matches_oracle():
    ds = util.small_dataset("pr1", chrom_len=50_000, n_events=3)
    rois = ds.sweep(min_reads=5)
    a = rois.arrays()
    assert rois.header() == orc.vcf_header(a["chrom_names"], [len(s) for s in a["chrom_seqs"]])
    assert rois.header().startswith("##fileformat=VCFv4.2\n") and rois.header().endswith("\tFORMAT\tsample\n")


def _dedup_reference(data):
    """src/indelope.nim:604-608 over record lines, the plain sequential way: drop a line whose CHROM, POS, REF, ALT equal those of one of the last two kept"""
    out, last = [], []
    for line in data.split(b"\n"):
        if not line:
            continue
        f = line.split(b"\t")
        if len(f) >= 6:
            key = (f[0], f[1], f[3], f[4])
            if key in last[-2:]:
                continue
            last = (last + [key])[-2:]
        out.append(line + b"\n")
    return b"".join(out)


def test_parallel_dedup_equals_the_sequential_machine():
    """idlh_vcf_dedup_inplace cuts the buffer into one chunk per thread and stitches the chunks exactly: streams with few distinct keys (long
    interactions across the cuts: ...Y Z Y Z...), empty lines, lines without the fields, a last line without newline, any thread count"""
    import ctypes
    rng = np.random.default_rng(11)
    for trial in range(12):
        nkeys = int(rng.choice([2, 3, 4, 6, 50, 5000]))
        n = int(rng.choice([40_000, 60_000]))
        ks = rng.integers(0, nkeys, n)
        lines = []
        for i, k in enumerate(ks):
            r = rng.random()
            if r < 0.01:
                lines.append(b"")
            elif r < 0.02:
                lines.append(b"# a line without the fields %d" % i)
            else:
                lines.append(b"chr%d\t%d\t.\t%s\t%s\t30\tPASS\tX=%d;pad=%s" % (k % 3, 100 + k, b"A" * (1 + k % 2), b"AC" * (1 + k % 5), i, b"p" * int(rng.integers(0, 40))))
        data = b"\n".join(lines) + (b"" if trial % 3 == 0 else b"\n")
        assert len(data) > (1 << 20)
        want = _dedup_reference(data)
        for threads in (1, 2, 3, 8, 16):
            host.set_threads(threads)
            buf = ctypes.create_string_buffer(data, len(data) + 2)
            kept = host.dedup_inplace(ctypes.addressof(buf), len(data))
            assert buf.raw[:kept] == want, (trial, nkeys, threads)
    host.set_threads(0)
    assert host.dedup_records(b"") == b"" and host.dedup_records(b"\n\n") == b""
