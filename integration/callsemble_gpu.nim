## integration/callsemble_gpu.nim -- the reference's calling loop over libindelope_cuda.so.
##
## UNVERIFIED (no Nim toolchain in the build image).  Meant to be `include`d at the bottom of src/indelope.nim (it uses that
## file's private names: roi, Variant, info_add, trim, skippable, genotype, mean, header ...), see integration/README.md for the
## three-line change of `main`.  The same recipe runs, compiled and tested byte for byte against the oracle, in C++:
## indelope_b200/csrc/host/pack_vcf.cpp (idlh_pack = `add` below, idlh_vcf_records = `variants` below) and
## indelope_b200/csrc/host/indelope_main.cpp (the modified `main`).
##
##   for target in targets:                                 # src/indelope.nim:601-608, was: for r in gen_roi: for v in callsemble
##     for r in gen_roi(b, target, ...):
##       for v in caller.add(r, target_index, faidx): emit(v)        # yields only when a batch came back
##   for v in caller.finish(faidx): emit(v)

import indelope_cuda

type
  PendingRoi = object
    chrom: string
    chrom_len, n_reads: int
  GpuCaller* = ref object
    ctx: IdlCtx
    params: IdlParams
    batch: array[2, ptr IdlBatch]        ## two pinned batches: one is packed while the other is in flight
    ticket: array[2, uint64]
    pending: array[2, seq[PendingRoi]]   ## what the host needs again when the results come back
    inflight: array[2, bool]
    cur: int
    soff, roff: uint64                   ## next free base offset in the read / reference pools (multiples of 64)
    ordinal: uint32

const
  BatchReads = 400_000
  BatchRegions = 20_000

proc check(rc: cint, what: string) =
  if rc != 0: quit("indelope: " & what & ": " & $idl_strerror(rc))

proc newGpuCaller*(min_reads, min_ctg_len, min_event_len: int, device = 0): GpuCaller =
  new(result)
  idl_default_params(addr result.params)
  result.params.min_reads = min_reads.int32
  result.params.min_ctg_len = min_ctg_len.int32
  result.params.min_event_len = min_event_len.int32
  check(idl_create(device.cint, addr result.params, addr result.ctx), "idl_create")   # IDL_E_NO_DEVICE: there is no CPU path
  for k in 0..1:
    check(idl_batch_alloc(result.ctx, BatchRegions + 1, BatchReads + 601, (BatchReads + 601) * 192, (BatchRegions + 1) * 1152, addr result.batch[k]),
          "idl_batch_alloc")

proc code(c: char): uint32 {.inline.} =
  ## src/ksw2/ksw2.nim:127-132; 4 = not ACGT.  (Lower case and IUPAC codes are folded: set IDL_RF_ALPHABET on the region, see the header.)
  case c
  of 'A', 'a': 0'u32
  of 'C', 'c': 1'u32
  of 'G', 'g': 2'u32
  of 'T', 't': 3'u32
  else: 4'u32

proc pack(pool2, pooln: ptr UncheckedArray[uint32], off: uint64, s: string): bool =
  ## 2 bits per base + the non-ACGT plane; every record starts on a 64-base boundary and all of its words are written.
  ## Returns true if a byte outside "ACGTN" was folded.  (hts-nim hands out the 4-bit BAM codes as a string; a production shim
  ## converts the nibbles with a 256-entry table, two bases per lookup, as pack_vcf.cpp does with 32 bases per AVX2 step.)
  let words = (s.len + 63) div 64 * 4
  for w in 0..<words: pool2[int(off shr 4) + w] = 0
  for w in 0..<(words div 2): pooln[int(off shr 5) + w] = 0
  for i, c in s:
    let b = off + uint64(i)
    let k = code(c)
    if k > 3'u32: pooln[int(b shr 5)] = pooln[int(b shr 5)] or (1'u32 shl (b and 31))
    else: pool2[int(b shr 4)] = pool2[int(b shr 4)] or (k shl (2 * (b and 15)))
    if c notin {'A', 'C', 'G', 'T', 'N'}: result = true

iterator variants(g: GpuCaller, k: int, fai: Fai): Variant =
  ## the filters, genotype likelihoods and VCF fields of src/indelope.nim:375-428 over the integers one batch returned
  var res: ptr IdlResults
  check(idl_wait(g.ctx, g.ticket[k], addr res), "idl_wait")
  for i in 0..<int(res.n_regions):
    let rr = res.region[i]
    let pr = g.pending[k][i]
    if rr.status != 0: stderr.write_line("indelope: region status " & $rr.status & " on " & pr.chrom & " (include/indelope_cuda.h IDL_RS_*)")
    for ci in 0..<int(rr.n_contigs):
      let c = res.contig[int(rr.contig_begin) + ci]
      if c.aln < 0: continue
      let a = res.aln[c.aln]
      if a.n_events < 1 or a.n_events > 4 or a.event_begin == IDL_NO_EVENTS or a.status != 0: continue       # :229
      var cc = ""                                                                                          # ez.cigar_string over the truncated CIGAR
      for x in 0..<int(a.n_cigar_trunc):
        let op = res.cigar[int(a.cigar_off) + x]
        cc &= $(op shr 4) & "MID"[int(op and 15)]
      for ei in 0..<int(a.n_events):
        let e = res.event[int(a.event_begin) + ei]
        if e.reject != IDL_EV_COUNTED: continue
        let ref_support = int(e.ref_support)
        let alt_support = int(e.alt_support)
        let both_found = int(e.both_found)
        let offset = int(e.offset)
        if alt_support < int(g.params.min_reads): continue                                                 # :375
        if float64(alt_support) / float64(pr.n_reads) < 0.1: continue                                      # :377
        var gt = genotype(ref_support, alt_support, 1e-3)
        if gt.GT == GT.HOM_REF: continue
        let ref_kmer = fai.get(pr.chrom, int(c.start) + int(e.tstart), int(c.start) + int(e.tstart) + IDL_KMER - 1)
        var alt_kmer = newString(IDL_KMER)
        for x in 0..<IDL_KMER: alt_kmer[x] = res.contig_seq[int(c.seq_off) + int(e.qstart) + x]
        var v = Variant(chrom: pr.chrom, start: int(e.t_start), genotype: gt, ref_kmer: ref_kmer, qual: gt.qual, alt_kmer: alt_kmer,
                        AD: [ref_support, alt_support])
        if offset == 0 and both_found >= int(0.75 * float64(min(ref_support, alt_support))): continue      # :385
        v.info_add("DP=" & $pr.n_reads)
        if offset < 5:
          v.info_add("LO"); v.qual /= 2'f64
        if both_found > 0:
          v.info_add("BS=" & $both_found); v.qual /= 1.5
        else:
          v.qual *= 2
        v.info_add("CC=" & cc)
        if e.aligned != 0: v.info_add("AL")
        if (int(e.min_flank) - 1) < max(int(e.t_stop - e.t_start), int(e.q_stop - e.q_start)): continue     # :401
        v.info_add("MF=" & $e.min_flank)
        v.info_add("CF=" & $offset)
        v.info_add("NC=" & $rr.n_contigs_pre)
        if offset == 0: v.qual /= 4'f64
        let ake = float64(e.sum_adist) / float64(e.n_adist)      # mean() of an empty list is 0/0 = NaN in the reference as well (:146-150)
        let rke = float64(e.sum_rdist) / float64(e.n_rdist)
        v.info_add("AKE=" & formatFloat(ake, precision = 2, format = ffDecimal))
        v.info_add("RKE=" & formatFloat(rke, precision = 2, format = ffDecimal))
        if e.n_adist > 0: v.info_add("AMQ=" & $e.amq_median)
        if e.n_rdist > 0: v.info_add("RMQ=" & $e.rmq_median)
        if ake < 5: continue                                                                               # :412
        if e.typ == 1:                                                                                     # deletion, :413-415
          v.reference = fai.get(pr.chrom, int(e.t_start) - 1, int(e.t_stop) - 1)
          v.alternate = v.reference[0..<1]
        else:                                                                                              # insertion, :420-426
          v.reference = fai.get(pr.chrom, int(e.t_start) - 1, int(e.t_start) - 1)
          v.alternate = newString(int(e.q_stop - (e.q_start - 1)))
          for x in 0..<v.alternate.len: v.alternate[x] = res.contig_seq[int(c.seq_off) + int(e.q_start) - 1 + x]
          var vset = v.alternate[1..v.alternate.high].toSet
          if vset.len == 1 and alt_kmer[alt_kmer.high-10..alt_kmer.high].toSet.len == 1 and
              ref_kmer[ref_kmer.high-10..ref_kmer.high].toSet.len == 1:
            continue
        yield v
  check(idl_release(g.ctx, g.ticket[k]), "idl_release")
  g.inflight[k] = false
  g.pending[k].setLen(0)

proc submit(g: GpuCaller) =
  let k = g.cur
  let b = g.batch[k]
  if b.n_regions == 0: return
  b.n_seq_bases = csize_t(g.soff); b.n_ref_bases = csize_t(g.roff)
  b.summary_valid = 0                      # let the library scan the records (or fill max_trim_len / max_ref_len / n_small_regions here)
  check(idl_submit(g.ctx, b, addr g.ticket[k]), "idl_submit")
  g.inflight[k] = true
  g.cur = 1 - k
  g.soff = 0; g.roff = 0

iterator add*(g: GpuCaller, r: roi, chrom_id: int, chrom_len: int, fai: Fai): Variant =
  ## append one region of interest (src/indelope.nim:21) to the batch being packed; when it is full, send it and yield the variants of
  ## the batch that was in flight (emission order = region order, so the last_var / last_var2 dedup of :604-608 sees the same sequence)
  var b = g.batch[g.cur]
  if int(b.n_reads) + r.reads.len > BatchReads or int(b.n_regions) >= BatchRegions:
    g.submit()
    if g.inflight[g.cur]:
      for v in g.variants(g.cur, fai): yield v
    b = g.batch[g.cur]
    b.n_regions = 0; b.n_reads = 0
  var reg: IdlRegion
  reg.chrom_id = chrom_id.int32; reg.roi_start = r.start.int32; reg.roi_end = r.stop.int32
  reg.read_begin = uint32(b.n_reads); reg.n_reads = uint32(r.reads.len); reg.ordinal = g.ordinal; inc g.ordinal
  var ws = high(int); var far = -1; var max_stop = -1
  var s = ""
  for rd in r.reads:
    var x: IdlRead
    let t = rd.trim()                                       # src/indelope.nim:23-38 -> (a, b) inclusive, or an empty range
    let tl = max(0, t.b - t.a + 1)
    discard rd.sequence(s)
    x.start = rd.start.int32; x.stop = rd.stop.int32; x.mapq = rd.qual.uint8
    x.flags = if rd.skippable: 1 else: 0
    if s.len > int(g.params.max_read_len): reg.flags = reg.flags or 2'u32        # IDL_RF_READ_TOO_LONG: packed empty, the region is dropped and reported
    else:
      x.seq_off = uint32(g.soff); x.len = s.len.uint16; x.trim_a = t.a.uint16; x.trim_len = tl.uint16
      x.min_overlap = uint16(int(0.88 * float64(tl)))       # :169
      if pack(b.seq2, b.seqn, g.soff, s): reg.flags = reg.flags or 1'u32        # IDL_RF_ALPHABET
      g.soff += uint64((s.len + 63) div 64 * 64)
    ws = min(ws, rd.start + t.a); far = max(far, rd.start + t.a + tl)
    if rd.qual > 5'u8: max_stop = max(max_stop, rd.stop)    # :213-216
    b.read[int(b.n_reads)] = x; b.n_reads += 1
  if ws == high(int) or ws < 0: ws = 0
  let we = min(chrom_len - 1, max(far, max_stop) + 63)       # the device slices fai.get(chrom, ctg.start, max_stop + 63) of :220 out of this window
  let win = fai.get(r.reads[0].chrom, ws, we)
  reg.ref_start = ws.int32; reg.ref_off = uint32(g.roff); reg.ref_len = uint32(win.len); reg.max_stop = max_stop.int32
  if pack(b.ref2, b.refn, g.roff, win): reg.flags = reg.flags or 1'u32
  g.roff += uint64((win.len + 63) div 64 * 64)
  b.region[int(b.n_regions)] = reg; b.n_regions += 1
  g.pending[g.cur].add(PendingRoi(chrom: r.reads[0].chrom, chrom_len: chrom_len, n_reads: r.reads.len))

iterator finish*(g: GpuCaller, fai: Fai): Variant =
  ## drain: the batch in flight first (it holds the earlier regions), then the one being packed
  let first = 1 - g.cur
  if g.inflight[first]:
    for v in g.variants(first, fai): yield v
  g.submit()
  let last = 1 - g.cur
  if g.inflight[last]:
    for v in g.variants(last, fai): yield v

proc close*(g: GpuCaller) =
  for k in 0..1: idl_batch_free(g.ctx, g.batch[k])
  idl_destroy(g.ctx)
