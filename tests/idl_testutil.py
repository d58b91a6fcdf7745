"""helpers shared by the parity tests"""
import os

import numpy as np

from indelope_b200 import host

HAS_GPU = None


def has_gpu():
    global HAS_GPU
    if HAS_GPU is None:
        try:
            import torch
            HAS_GPU = bool(torch.cuda.is_available())
        except Exception:
            HAS_GPU = False
    return HAS_GPU


def small_dataset(name="pr1", **over):
    cfg = dict(host.CONFIGS[name])
    cfg.update(over)
    return host.Dataset(**cfg)


def diff_lines(a, b, limit=12):
    """first differing lines of two dumps, grouped by record type"""
    la, lb = a.splitlines(), b.splitlines()
    out = []
    if len(la) != len(lb):
        out.append("line counts differ: %d vs %d" % (len(la), len(lb)))
    n = 0
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            out.append("line %d:\n  oracle: %s\n  gpu   : %s" % (i, x[:600], y[:600]))
            n += 1
            if n >= limit:
                break
    return "\n".join(out)


def by_type(dump):
    d = {}
    for l in dump.splitlines():
        d.setdefault(l[:1], []).append(l)
    return d


def rois_from_reads(reads, chrom_seq, roi_start=None, roi_stop=None, chrom_name="chrT"):
    """build a one-region Rois from a list of dicts(start, seq, qual=None, mapq=60, flag=0, stop=None)"""
    starts, stops, mapq, flag, lens, offs = [], [], [], [], [], []
    bases, quals = [], []
    off = 0
    for r in reads:
        s = r["seq"]
        starts.append(r["start"]); stops.append(r.get("stop", r["start"] + len(s))); mapq.append(r.get("mapq", 60)); flag.append(r.get("flag", 0))
        lens.append(len(s)); offs.append(off); off += len(s)
        bases.append(np.frombuffer(s.encode(), dtype=np.uint8))
        q = r.get("qual")
        quals.append(np.full(len(s), 30, dtype=np.uint8) if q is None else np.asarray(q, dtype=np.uint8))
    n = len(reads)
    arrays = dict(
        start=np.array(starts, np.int32), stop=np.array(stops, np.int32), mapq=np.array(mapq, np.uint8), flag=np.array(flag, np.uint16),
        len=np.array(lens, np.int32), seq_off=np.array(offs, np.int64),
        bases=np.concatenate(bases) if bases else np.zeros(0, np.uint8), quals=np.concatenate(quals) if quals else np.zeros(0, np.uint8),
        roi_chrom=np.zeros(1, np.int32), roi_start=np.array([roi_start if roi_start is not None else min(starts)], np.int32),
        roi_stop=np.array([roi_stop if roi_stop is not None else max(stops)], np.int32), roi_read_begin=np.zeros(1, np.int64),
        roi_n_reads=np.array([n], np.int32), read_idx=np.arange(n, dtype=np.int64),
        chrom_names=[chrom_name], chrom_seqs=[np.frombuffer(chrom_seq.encode(), dtype=np.uint8) if isinstance(chrom_seq, str) else chrom_seq],
    )
    return host.Rois(arrays=arrays), arrays


def merge_rois(list_of_arrays):
    """concatenate several one-chromosome region sets that share the same chromosome into one"""
    out = {k: [] for k in ("start", "stop", "mapq", "flag", "len", "seq_off", "bases", "quals", "roi_chrom", "roi_start", "roi_stop", "roi_read_begin",
                           "roi_n_reads", "read_idx")}
    nread = nbase = nidx = 0
    for a in list_of_arrays:
        for k in ("start", "stop", "mapq", "flag", "len", "bases", "quals", "roi_chrom", "roi_start", "roi_stop", "roi_n_reads"):
            out[k].append(a[k])
        out["seq_off"].append(a["seq_off"] + nbase)
        out["roi_read_begin"].append(a["roi_read_begin"] + nidx)
        out["read_idx"].append(a["read_idx"] + nread)
        nread += len(a["start"]); nbase += len(a["bases"]); nidx += len(a["read_idx"])
    m = {k: np.concatenate(v) for k, v in out.items()}
    m["chrom_names"] = list_of_arrays[0]["chrom_names"]; m["chrom_seqs"] = list_of_arrays[0]["chrom_seqs"]
    return m


def out_dir():
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return d


# ---------------------------------------------------------------------------------------------------------------
# BAM files made and parsed with struct + zlib alone (SAM spec 4.1, 4.2): the independent side of the idl_bam_* tests
# ---------------------------------------------------------------------------------------------------------------
import struct
import zlib

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
SEQ16 = "=ACMGRSVTWYHKDBN"
CIG_OPS = "MIDNSHP=X"


def bgzf_member(chunk, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    c = co.compress(chunk) + co.flush()
    bsize = 18 + len(c) + 8 - 1
    assert bsize < 65536
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + c + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def bgzf_compress(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, block=0xff00, levels=None, eof=True):
    """data -> BGZF; `levels` cycles (level, strategy) pairs over the members so that one file holds stored, fixed and dynamic blocks"""
    out, k = [], 0
    for at in range(0, len(data), block):
        lv, st = (levels[k % len(levels)] if levels else (level, strategy)); k += 1
        chunk = data[at:at + block]
        if lv == 0 and len(chunk) > 0xff00 - 64:
            lv = 1  # a stored member must still fit 64 KiB with its framing
        out.append(bgzf_member(chunk, lv, st))
    if eof:
        out.append(BGZF_EOF)
    return b"".join(out)


def bam_record(ref_id, pos, cigar, seq, qual=None, mapq=60, flag=0, name=b"r", next_ref=-1, next_pos=-1, tlen=0, tags=b""):
    """cigar: list of (op char, len); seq: str over SEQ16"""
    l_seq = len(seq)
    nib = [SEQ16.index(c) for c in seq]
    if l_seq & 1:
        nib.append(0)
    packed = bytes(nib[i] << 4 | nib[i + 1] for i in range(0, len(nib), 2))
    qual = bytes([30] * l_seq) if qual is None else bytes(qual)
    assert len(qual) == l_seq
    cg = b"".join(struct.pack("<I", ln << 4 | CIG_OPS.index(op)) for op, ln in cigar)
    nm = name + b"\0"
    body = struct.pack("<iiBBHHHiiii", ref_id, pos, len(nm), mapq, 0, len(cigar), flag, l_seq, next_ref, next_pos, tlen) + nm + cg + packed + qual + tags
    return struct.pack("<i", len(body)) + body


def bam_bytes(refs, records, text=None):
    """refs: list of (name, length); records: list of bytes from bam_record -> the uncompressed BAM stream"""
    text = ("@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)).encode() if text is None else text
    out = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", len(refs))]
    for name, ln in refs:
        nm = name.encode() + b"\0"
        out.append(struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln))
    return b"".join(out) + b"".join(records)


def parse_bam(data):
    """uncompressed BAM stream -> header + per-record arrays as idl_bam_fetch lays them out (records without a target dropped)"""
    assert data[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].decode()
    at = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, at); at += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, at); at += 4
        name = data[at:at + l_name - 1].decode(); at += l_name
        ln, = struct.unpack_from("<i", data, at); at += 4
        refs.append((name, ln))
    cols = dict(chrom=[], start=[], stop=[], len=[], mapq=[], flag=[])
    seq_off, cig_off, bases, quals, cigar = [0], [0], [], [], []
    n_unplaced = 0
    while at < len(data):
        bs, ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiiBBHHHi", data, at)
        b = at + 4; at += 4 + bs
        if ref_id < 0:
            n_unplaced += 1
            continue
        cg = struct.unpack_from("<%dI" % n_cig, data, b + 32 + l_name)
        rlen = sum(c >> 4 for c in cg if (c & 15) in (0, 2, 3, 7, 8))
        if (flag & 4) or n_cig == 0 or rlen == 0:
            rlen = 1
        so = b + 32 + l_name + 4 * n_cig
        packed = np.frombuffer(data, dtype=np.uint8, count=(l_seq + 1) // 2, offset=so)
        nib = np.empty(2 * len(packed), dtype=np.uint8); nib[0::2] = packed >> 4; nib[1::2] = packed & 15
        bases.append(np.frombuffer(SEQ16.encode(), dtype=np.uint8)[nib[:l_seq]])
        quals.append(np.frombuffer(data, dtype=np.uint8, count=l_seq, offset=so + (l_seq + 1) // 2))
        cigar.extend(cg)
        for k, v in zip(("chrom", "start", "stop", "len", "mapq", "flag"), (ref_id, pos, pos + rlen, l_seq, mapq, flag)):
            cols[k].append(v)
        seq_off.append(seq_off[-1] + l_seq); cig_off.append(cig_off[-1] + n_cig)
    out = dict(chrom=np.array(cols["chrom"], np.int32), start=np.array(cols["start"], np.int32), stop=np.array(cols["stop"], np.int32), len=np.array(cols["len"], np.int32),
               mapq=np.array(cols["mapq"], np.uint8), flag=np.array(cols["flag"], np.uint16), seq_off=np.array(seq_off, np.int64), cig_off=np.array(cig_off, np.uint64),
               bases=np.concatenate(bases) if bases else np.zeros(0, np.uint8), quals=np.concatenate(quals) if quals else np.zeros(0, np.uint8),
               cigar=np.array(cigar, np.uint32))
    return text, refs, out, n_unplaced
