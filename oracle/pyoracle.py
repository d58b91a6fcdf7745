"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package indelope_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libksw2_ref.so")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)

EZ_FIELDS = "max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar".split()


class Ez(C.Structure):
    _fields_ = [(n, C.c_int32) for n in EZ_FIELDS] + [("cells", C.c_int64), ("diagonals", C.c_int32), ("status", C.c_int32)]


class RoiSet(C.Structure):
    _fields_ = [
        ("n_reads", C.c_int64), ("start", i32p), ("stop", i32p), ("mapq", u8p), ("flag", u16p), ("len", i32p),
        ("seq_off", i64p), ("bases", u8p), ("quals", u8p),
        ("n_rois", C.c_int64), ("roi_chrom", i32p), ("roi_start", i32p), ("roi_stop", i32p), ("roi_read_begin", i64p),
        ("roi_n_reads", i32p), ("read_idx", i64p),
        ("n_chroms", C.c_int32), ("chrom_name", C.POINTER(C.c_char_p)), ("chrom_seq", C.POINTER(u8p)), ("chrom_len", i64p),
    ]


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in "min_reads min_ctg_len min_event_len use_ref_ksw2 dump_level n_threads".split()]


COUNTER_FIELDS = ("regions reads slide_calls offsets char_compares exhaustive_compares contigs_pre contigs_post dp_a dp_b "
                  "cells_a cells_b events kmer_reads kmer_windows kmer_bytes al_events variants").split()


COUNTER_TAIL = "corrections vote_invariant_violations left_merges merges".split()


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in COUNTER_FIELDS] + [("seconds", C.c_double)] + [(n, C.c_int64) for n in COUNTER_TAIL]


class Match(C.Structure):
    _fields_ = [("matches", C.c_int64), ("offset", C.c_int64), ("mismatches", C.c_int64), ("aligned", C.c_int32),
                ("n_corr", C.c_int32), ("corr", C.c_int32 * 192)]


def build(force=False):
    """compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)"""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
            os.path.getmtime(os.path.join(HERE, f)) for f in ("oracle.cpp", "oracle.h", "ksw2_lane.c", "ksw2_lane.h")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(REF_LIB)):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_vcf_header.restype = C.c_void_p
        _lib.orc_assemble_strings.restype = C.c_void_p
        _lib.orc_free.argtypes = [C.c_void_p]
    return _lib


def have_ref():
    return os.path.exists(REF_LIB)


def ref():
    """the reference's own ksw2_extz2_sse.c, compiled unmodified (oracle/_ref/libksw2_ref.so)"""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_LIB)
    return _ref


def load_ref_into_oracle():
    return lib().orc_load_ref(REF_LIB.encode())


def _take(ptr):
    s = C.string_at(ptr).decode()
    lib().orc_free(ptr)
    return s


ENC = np.full(256, 4, dtype=np.uint8)
for _c, _v in zip("ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
    ENC[ord(_c)] = _v


def encode(s):
    """src/ksw2/ksw2.nim:127-132"""
    if isinstance(s, str):
        s = s.encode()
    return ENC[np.frombuffer(s, dtype=np.uint8)] if not isinstance(s, np.ndarray) else s


def cigar_str(c):
    return "".join("%d%s" % (x >> 4, "MID"[x & 15]) for x in c)


def ksw2(q, t, match=1, mismatch=-2, gapo=4, gape=1, w=-1, zdrop=-1, flag=0, impl="lane"):
    """returns (fields dict, full cigar tuple, Ez). impl: 'lane' (own restatement) or 'ref' (compiled reference)"""
    qa = np.ascontiguousarray(encode(q), dtype=np.uint8)
    ta = np.ascontiguousarray(encode(t), dtype=np.uint8)
    cap = len(qa) + len(ta) + 8
    cig = (C.c_uint32 * cap)()
    ez = Ez()
    args = [len(qa), qa.ctypes.data_as(u8p), len(ta), ta.ctypes.data_as(u8p), C.c_int8(match), C.c_int8(mismatch),
            C.c_int8(gapo), C.c_int8(gape), C.c_int(w), C.c_int(zdrop)]
    if impl == "ref":
        ref().orc_ksw2_ref(*args, C.c_int(flag), C.byref(ez), cig, cap)
    else:
        assert flag == 0
        lib().orc_ksw2_lane(*args, C.byref(ez), cig, cap)
    f = {n: getattr(ez, n) for n in EZ_FIELDS}
    return f, tuple(cig[:max(ez.n_cigar, 0)]), ez


def slide_align(q, t, min_overlap=50, max_mismatch=0, qsup=1, tsup=1, rule=0, qreads=None, treads=None):
    """src/contig.nim:152-154 (string form builds make_contig(.., support))"""
    qs = np.full(len(q), qsup, dtype=np.uint32) if np.isscalar(qsup) else np.asarray(qsup, dtype=np.uint32)
    ts = np.full(len(t), tsup, dtype=np.uint32) if np.isscalar(tsup) else np.asarray(tsup, dtype=np.uint32)
    qreads = int(qsup) if qreads is None else qreads
    treads = int(tsup) if treads is None else treads
    m = Match()
    lib().orc_slide_align(q.encode(), qs.ctypes.data_as(u32p), C.c_int64(qreads), t.encode(), ts.ctypes.data_as(u32p),
                          C.c_int64(treads), C.c_int64(min_overlap), C.c_int64(max_mismatch), rule, C.byref(m))
    return m


def insert(t, tstart, tsup, q, qstart, qsup, m):
    """src/contig.nim:156-222 on two make_contig(seq, start, support) contigs; returns (seq, support list, start, nreads)"""
    n = len(t) + len(q) + 1
    tb = C.create_string_buffer(t.encode(), n)
    qb = C.create_string_buffer(q.encode(), n)
    tsa = np.zeros(n, dtype=np.uint32); tsa[:len(t)] = tsup
    qsa = np.zeros(n, dtype=np.uint32); qsa[:len(q)] = qsup
    tlen = C.c_int64(len(t)); treads = C.c_int64(int(tsup)); tst = C.c_int64(tstart)
    lib().orc_insert(tb, tsa.ctypes.data_as(u32p), C.byref(tlen), C.byref(treads), C.byref(tst), qb, qsa.ctypes.data_as(u32p),
                     C.c_int64(len(q)), C.c_int64(int(qsup)), C.c_int64(qstart), C.byref(m))
    return tb.value.decode()[:tlen.value], [int(x) for x in tsa[:tlen.value]], tst.value, treads.value


def assemble_strings(seqs, starts, min_overlaps, combine=True):
    n = len(seqs)
    arr = (C.c_char_p * n)(*[s.encode() for s in seqs])
    st = np.asarray(starts, dtype=np.int64); mo = np.asarray(min_overlaps, dtype=np.int64)
    npre = C.c_int64()
    p = lib().orc_assemble_strings(n, arr, st.ctypes.data_as(i64p), mo.ctypes.data_as(i64p), 1 if combine else 0, C.byref(npre))
    return _take(p), npre.value


def genotype(r, a, error):
    buf = C.create_string_buffer(256)
    q = C.c_double()
    g = lib().orc_genotype(C.c_int64(r), C.c_int64(a), C.c_double(error), buf, 256, C.byref(q))
    return g, buf.value.decode(), q.value


def trim(quals):
    qa = np.ascontiguousarray(quals, dtype=np.uint8)
    n = C.c_int32()
    a = lib().orc_trim(qa.ctypes.data_as(u8p), len(qa), C.byref(n))
    return a, n.value


def mincode(s, k=27):
    code = C.c_uint64()
    rc = lib().orc_mincode(s.encode(), k, C.byref(code))
    return None if rc else code.value


def vcf_header(names, lens):
    n = len(names)
    arr = (C.c_char_p * n)(*[s.encode() for s in names])
    la = np.asarray(lens, dtype=np.int64)
    return _take(lib().orc_vcf_header(n, arr, la.ctypes.data_as(i64p)))


class RoiSetArrays:
    """owns the numpy arrays behind an orc_roiset_t (same flat layout indelope_b200.host.RoiSet exports)"""

    def __init__(self, d):
        dt = dict(start=np.int32, stop=np.int32, mapq=np.uint8, flag=np.uint16, len=np.int32, seq_off=np.int64, bases=np.uint8,
                  quals=np.uint8, roi_chrom=np.int32, roi_start=np.int32, roi_stop=np.int32, roi_read_begin=np.int64,
                  roi_n_reads=np.int32, read_idx=np.int64)
        self.d = {k: (np.ascontiguousarray(v, dtype=dt[k]) if k in dt else v) for k, v in d.items()}
        self.d["chrom_seqs"] = [np.ascontiguousarray(s, dtype=np.uint8) for s in d["chrom_seqs"]]
        r = RoiSet()
        g = self.d
        r.n_reads = len(g["start"])
        r.start = g["start"].ctypes.data_as(i32p)
        r.stop = g["stop"].ctypes.data_as(i32p)
        r.mapq = g["mapq"].ctypes.data_as(u8p)
        r.flag = g["flag"].ctypes.data_as(u16p)
        r.len = g["len"].ctypes.data_as(i32p)
        r.seq_off = g["seq_off"].ctypes.data_as(i64p)
        r.bases = g["bases"].ctypes.data_as(u8p)
        r.quals = g["quals"].ctypes.data_as(u8p)
        r.n_rois = len(g["roi_start"])
        r.roi_chrom = g["roi_chrom"].ctypes.data_as(i32p)
        r.roi_start = g["roi_start"].ctypes.data_as(i32p)
        r.roi_stop = g["roi_stop"].ctypes.data_as(i32p)
        r.roi_read_begin = g["roi_read_begin"].ctypes.data_as(i64p)
        r.roi_n_reads = g["roi_n_reads"].ctypes.data_as(i32p)
        r.read_idx = g["read_idx"].ctypes.data_as(i64p)
        names = g["chrom_names"]
        self._names = (C.c_char_p * len(names))(*[s.encode() for s in names])
        self._seqs = (u8p * len(names))(*[s.ctypes.data_as(u8p) for s in g["chrom_seqs"]])
        self._lens = np.array([len(s) for s in g["chrom_seqs"]], dtype=np.int64)
        r.n_chroms = len(names)
        r.chrom_name = self._names
        r.chrom_seq = self._seqs
        r.chrom_len = self._lens.ctypes.data_as(i64p)
        self.c = r


DUMP_ALL = 31


def call(roiset, min_reads=3, min_ctg_len=73, min_event_len=4, use_ref_ksw2=False, dump_level=DUMP_ALL, n_threads=1):
    """run the oracle over a flat region set (dict of arrays, see RoiSetArrays). Returns (dump, vcf_records, counters dict)"""
    rs = roiset if isinstance(roiset, RoiSetArrays) else RoiSetArrays(roiset)
    if use_ref_ksw2:
        assert load_ref_into_oracle() == 0
    p = Params(min_reads, min_ctg_len, min_event_len, 1 if use_ref_ksw2 else 0, dump_level, n_threads)
    dump = C.c_void_p(); vcf = C.c_void_p(); cnt = Counters()
    rc = lib().orc_call(C.byref(rs.c), C.byref(p), C.byref(dump), C.byref(vcf), C.byref(cnt))
    assert rc == 0, rc
    cd = {n: getattr(cnt, n) for n in COUNTER_FIELDS + COUNTER_TAIL}
    cd["seconds"] = cnt.seconds
    return _take(dump.value), _take(vcf.value), cd
