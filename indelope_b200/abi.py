"""ctypes mirrors of the structs in include/indelope_cuda.h (no library is loaded here)."""
import ctypes as C

ABI_VERSION = 2
u32p = C.POINTER(C.c_uint32)

STAGE_ASSEMBLE, STAGE_ALIGN, STAGE_GENOTYPE, STAGE_ALL = 1, 2, 4, 7
OUT_SUPPORT = 1
RS_CONTIG_OVERFLOW, RS_CORR_OVERFLOW, RS_DP_OVERFLOW, RS_CIGAR_OVERFLOW, RS_READ_TOO_LONG, RS_ALPHABET, RS_BAD_INPUT = 1, 2, 4, 8, 16, 32, 64

PARAM_FIELDS = ("abi_version min_reads min_ctg_len min_event_len asm_min_mapq combine_min_support combine_min_overlap max_contigs "
                "stop_min_mapq window_pad match mismatch a_gapo a_gape a_bw a_zdrop b_gapo b_gape b_bw b_zdrop max_events count_min_mapq "
                "max_contig_len max_read_len max_reads_per_region n_streams").split()


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in PARAM_FIELDS] + [("stages", C.c_uint32), ("out_flags", C.c_uint32)]


def default_params(**kw):
    """the reference's literals; mirrors idl_default_params"""
    p = Params(abi_version=ABI_VERSION, min_reads=3, min_ctg_len=73, min_event_len=4, asm_min_mapq=20, combine_min_support=3,
               combine_min_overlap=65, max_contigs=20, stop_min_mapq=5, window_pad=63, match=1, mismatch=-2, a_gapo=4, a_gape=1, a_bw=50,
               a_zdrop=400, b_gapo=5, b_gape=1, b_bw=-1, b_zdrop=-1, max_events=4, count_min_mapq=10, max_contig_len=4096, max_read_len=512,
               max_reads_per_region=601, n_streams=2, stages=STAGE_ALL, out_flags=0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class Region(C.Structure):
    _fields_ = [("chrom_id", C.c_int32), ("roi_start", C.c_int32), ("roi_end", C.c_int32), ("read_begin", C.c_uint32), ("n_reads", C.c_uint32),
                ("ref_start", C.c_int32), ("ref_off", C.c_uint32), ("ref_len", C.c_uint32), ("max_stop", C.c_int32), ("ordinal", C.c_uint32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class Read(C.Structure):
    _fields_ = [("start", C.c_int32), ("stop", C.c_int32), ("seq_off", C.c_uint32), ("len", C.c_uint16), ("trim_a", C.c_uint16),
                ("trim_len", C.c_uint16), ("min_overlap", C.c_uint16), ("mapq", C.c_uint8), ("flags", C.c_uint8), ("reserved", C.c_uint16)]


class Batch(C.Structure):
    _fields_ = [("cap_regions", C.c_size_t), ("cap_reads", C.c_size_t), ("cap_seq_bases", C.c_size_t), ("cap_ref_bases", C.c_size_t),
                ("n_regions", C.c_size_t), ("n_reads", C.c_size_t), ("n_seq_bases", C.c_size_t), ("n_ref_bases", C.c_size_t),
                ("region", C.POINTER(Region)), ("read", C.POINTER(Read)), ("seq2", u32p), ("seqn", u32p), ("ref2", u32p), ("refn", u32p),
                ("impl", C.c_void_p), ("summary_valid", C.c_uint32), ("max_trim_len", C.c_uint32), ("max_ref_len", C.c_uint32),
                ("max_region_reads", C.c_uint32), ("n_small_regions", C.c_size_t)]


class RegionResult(C.Structure):
    _fields_ = [("status", C.c_uint32), ("n_contigs_pre", C.c_int32), ("n_contigs", C.c_int32), ("contig_begin", C.c_uint32)]


class ContigResult(C.Structure):
    _fields_ = [("start", C.c_int32), ("nreads", C.c_int32), ("len", C.c_int32), ("seq_off", C.c_uint32), ("aln", C.c_int32), ("region", C.c_uint32)]


class AlnResult(C.Structure):
    _fields_ = [("region", C.c_uint32), ("contig", C.c_uint32), ("ref_len", C.c_int32)] + \
               [(n, C.c_int32) for n in "max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar n_cigar_trunc".split()] + \
               [("cigar_off", C.c_uint32), ("n_events", C.c_int32), ("event_begin", C.c_uint32), ("status", C.c_uint32)]


class EventResult(C.Structure):
    _fields_ = [("aln", C.c_uint32)] + \
               [(n, C.c_int32) for n in ("index type t_start t_stop q_start q_stop len reject tstart qstart offset min_flank k_ref k_alt k_both aligned "
                                         "ref_support alt_support both_found n_adist n_rdist").split()] + \
               [("sum_adist", C.c_int64), ("sum_rdist", C.c_int64), ("amq_median", C.c_int32), ("rmq_median", C.c_int32),
                ("ref_code", C.c_uint64), ("alt_code", C.c_uint64)]


class Results(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in "n_regions n_contigs n_alns n_events n_cigar_ops n_contig_bases".split()] + \
               [("region", C.POINTER(RegionResult)), ("contig", C.POINTER(ContigResult)), ("aln", C.POINTER(AlnResult)),
                ("event", C.POINTER(EventResult)), ("cigar", u32p), ("contig_seq", C.POINTER(C.c_char)), ("contig_support", u32p)] + \
               [(n, C.c_float) for n in "ms_h2d ms_assemble ms_align ms_genotype ms_al ms_d2h ms_total".split()] + \
               [(n, C.c_uint64) for n in "offsets_tested dp_cells_a dp_cells_b dp_a dp_b kmer_reads kmer_bytes al_events".split()] + \
               [("kernel_launches", C.c_uint32), ("pool_retries", C.c_uint32)]


class Ez(C.Structure):
    _fields_ = [(n, C.c_int32) for n in "max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar status reserved".split()] + [("cells", C.c_int64)]


assert C.sizeof(Region) == 48 and C.sizeof(Read) == 24
