// indelope_b200/csrc/genotype.cuh -- the part of callsemble after assembly (src/indelope.nim:208-372):
//
//   align_kernel  call-site A of kernel 2 (contig -> reference window, bw=50 z=400, :213-221) followed, in the same
//                 warp, by the glue: truncated CIGAR (src/ksw2/ksw2.nim:22-33), target/query event locations
//                 (:71-91), ref/alt 27-mer selection and rejects (src/indelope.nim:229-281), get_min_flank (:118-132)
//   kmer_kernel   kernel 3: ref/alt canonical 27-mer counting over the reads of the region (:283-311), one CTA per
//                 event, one read per thread, 128-bit loads of the 2-bit packed reads, rolling forward/reverse codes
//   al_kernel     AL fallback (:312-372): call-site B of kernel 2, two unbanded alignments per read, vote by
//                 count_flanked_cigar (:185-199)
#pragma once
#include "common.cuh"
#include "ksw2.cuh"

#define DP_WARPS 8
#define DP_THREADS (DP_WARPS * 32)
#define KMER_THREADS 128

struct AlEntry { unsigned event; unsigned base; unsigned n_reads; unsigned pad; };

struct GenoArgs {
	const idl_region *region; const idl_read *read;
	const uint32_t *seq2, *seqn;
	const uint8_t *refcodes; const uint8_t *ctg_codes;
	idl_region_result *rres; idl_contig_result *cres; idl_aln_result *ares; idl_event_result *eres; uint32_t *cigar;
	unsigned cap_events, cap_cigar, cap_al, cap_alns;
	AlEntry *al_list;
	idl_params P;
	DevCounters *cnt;
	// DP workspaces: one per resident warp
	uint8_t *pmat; size_t p_cap;
	uint32_t *cig_scratch; int cig_cap;
	int8_t *spill; int spill_tcap;   // global-memory lane storage for targets that do not fit shared memory
	int t_cap, hr, qcap;             // shared-memory lane capacity, H ring size, query buffer bytes (AL)
	int seq_cap;                     // shared-memory bytes per warp for the staged target / reversed query
	size_t spill_bytes;              // global-memory spill area per warp: lanes, then sequences
};

__host__ __device__ inline size_t dp_smem_per_warp(int t_cap, int hr, int qcap, int seq_cap)
{
	return ((ksw_lane_bytes(t_cap) + 15) & ~(size_t)15) + (size_t)hr * 4 + KSW_BTILE_BYTES + (size_t)((qcap + 15) & ~15) + (size_t)((seq_cap + 15) & ~15);
}

struct DpWarp { int8_t *lanes; int *H; uint8_t *btile; uint8_t *qbuf; uint8_t *seq; uint8_t *pmat; uint32_t *cig; int8_t *spill; uint8_t *spill_seq; };

__device__ __forceinline__ DpWarp dp_carve(const GenoArgs &g, unsigned char *smem)
{
	DpWarp d;
	const size_t per = dp_smem_per_warp(g.t_cap, g.hr, g.qcap, g.seq_cap);
	unsigned char *base = smem + per * warp_id();
	d.lanes = (int8_t*)base;
	d.H = (int*)(base + ((ksw_lane_bytes(g.t_cap) + 15) & ~(size_t)15));
	d.btile = (uint8_t*)(d.H + g.hr);
	d.qbuf = d.btile + KSW_BTILE_BYTES;
	d.seq = d.qbuf + ((g.qcap + 15) & ~15);
	const size_t gw = (size_t)blockIdx.x * DP_WARPS + warp_id();
	d.pmat = g.pmat + gw * g.p_cap;
	d.cig = g.cig_scratch + gw * (size_t)g.cig_cap;
	d.spill = g.spill + gw * g.spill_bytes;
	d.spill_seq = (uint8_t*)d.spill + ((ksw_lane_bytes(g.spill_tcap) + 15) & ~(size_t)15);
	return d;
}

// run one alignment, spilling the lane arrays to global memory when the target does not fit shared memory
__device__ __forceinline__ void dp_run(const GenoArgs &g, const DpWarp &d, int qlen, const uint8_t *q, int tlen, const uint8_t *t, KswParams kp, KswOut &o)
{
	const int T16 = (tlen + 15) & ~15;
	const bool fits = T16 <= g.t_cap;
	uint8_t *seq = ksw_seq_bytes(qlen, tlen) <= (size_t)g.seq_cap ? d.seq : d.spill_seq;
	ksw2_warp(qlen, q, tlen, t, kp, fits ? d.lanes : d.spill, fits ? g.t_cap : g.spill_tcap, seq, d.H, g.hr, d.btile, d.pmat, g.p_cap, d.cig, g.cig_cap, o);
}

__device__ __forceinline__ unsigned dp_status_bits(int st)
{
	if (st == KSW_ST_CIGCAP) return IDL_RS_CIGAR_OVERFLOW;
	if (st < 0) return IDL_RS_DP_OVERFLOW;
	return 0;
}

// canonical 2-bit code of K codes (declared semantics of kmer.mincode); ~0 if a base is not ACGT
__device__ uint64_t canon_code(const uint8_t *s, int K)
{
	uint64_t f = 0, rc = 0;
	for (int i = 0; i < K; ++i) {
		const unsigned b = s[i];
		if (b > 3) return ~0ULL;
		f = (f << 2) | b;
		rc |= (uint64_t)(3 - b) << (2 * i);
	}
	return f < rc ? f : rc;
}

__device__ __forceinline__ bool same_k(const uint8_t *a, const uint8_t *b, int K)
{
	for (int i = 0; i < K; ++i) if (a[i] != b[i]) return false;
	return true;
}
__device__ __forceinline__ int distinct_k(const uint8_t *a, int K)
{
	unsigned m = 0;
	for (int i = 0; i < K; ++i) m |= 1u << a[i];
	return __popc(m);
}

__global__ void __launch_bounds__(DP_THREADS, 3) align_kernel(GenoArgs g)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DpWarp d = dp_carve(g, smem_raw);
	const int lane = lane_id();
	const idl_params &P = g.P;
	const int K = IDL_KMER, width = (K + 1) / 2 - 1; // :218
	KswParams kp; kp.match = (int8_t)P.match; kp.mismatch = (int8_t)P.mismatch; kp.q = (int8_t)P.a_gapo; kp.e = (int8_t)P.a_gape; kp.w = P.a_bw; kp.zdrop = P.a_zdrop;
	const unsigned n_alns = g.cnt->n_alns < g.cap_alns ? g.cnt->n_alns : g.cap_alns;
	for (;;) {
		unsigned ai = 0;
		if (lane == 0) ai = atomicAdd(&g.cnt->aln_next, 1u);
		ai = __shfl_sync(FULL_MASK, ai, 0);
		if (ai >= n_alns) break;
		idl_aln_result ar = g.ares[ai];
		const idl_contig_result cr = g.cres[ar.contig];
		const idl_region R = g.region[ar.region];
		// reference window of :213-220: fai.get(chrom, ctg.start, max_stop + width + 50), clipped to the shipped window
		const int win_end = R.ref_start + (int)R.ref_len - 1;
		int max_stop = cr.start > R.max_stop ? cr.start : R.max_stop;
		int end = max_stop + P.window_pad; if (end > win_end) end = win_end;
		int tlen = end - cr.start + 1;
		if (tlen < 0 || cr.start < R.ref_start) tlen = 0;
		const uint8_t *tq = g.refcodes + R.ref_off + (cr.start - R.ref_start);
		const uint8_t *qq = g.ctg_codes + cr.seq_off;
		KswOut o;
		dp_run(g, d, cr.len, qq, tlen, tq, kp, o);
		const int n = o.n_cigar;
		int ntr = 0, nev = 0;
		if (lane == 0) {
			ntr = ksw_trunc_count(d.cig, n, o.max_q);
			for (int k = 0; k < ntr; ++k) nev += (d.cig[n - 1 - k] & 0xf) != 0;
		}
		ntr = __shfl_sync(FULL_MASK, ntr, 0); nev = __shfl_sync(FULL_MASK, nev, 0);
		unsigned coff = 0, eoff = 0;
		const bool want_events = (P.stages & IDL_STAGE_GENOTYPE) && nev >= 1 && nev <= P.max_events; // :229
		if (lane == 0) {
			coff = atomicAdd(&g.cnt->n_cigar_ops, (unsigned)n);
			if (want_events) eoff = atomicAdd(&g.cnt->n_events, (unsigned)nev);
		}
		coff = __shfl_sync(FULL_MASK, coff, 0); eoff = __shfl_sync(FULL_MASK, eoff, 0);
		unsigned st = dp_status_bits(o.status);
		if (coff + (unsigned)n > g.cap_cigar) { st |= IDL_RS_CIGAR_OVERFLOW; if (lane == 0) atomicOr(&g.cnt->overflow, 8u); }
		else for (int k = lane; k < n; k += 32) g.cigar[coff + k] = d.cig[n - 1 - k]; // forward order
		bool ev_ok = want_events;
		if (ev_ok && eoff + (unsigned)nev > g.cap_events) { ev_ok = false; if (lane == 0) atomicOr(&g.cnt->overflow, 16u); }
		if (lane == 0) {
			ar.ref_len = tlen; ar.max = o.max; ar.zdropped = o.zdropped; ar.max_q = o.max_q; ar.max_t = o.max_t; ar.mqe = o.mqe; ar.mqe_t = o.mqe_t;
			ar.mte = o.mte; ar.mte_q = o.mte_q; ar.score = o.score; ar.n_cigar = n; ar.n_cigar_trunc = ntr; ar.cigar_off = coff; ar.n_events = nev;
			ar.event_begin = ev_ok ? eoff : IDL_NO_EVENTS; ar.status = st;
			g.ares[ai] = ar;
			if (st) atomicOr(&g.rres[ar.region].status, st);
			atomicAdd(&g.cnt->dp_a, 1ULL); atomicAdd(&g.cnt->dp_cells_a, (unsigned long long)o.cells);
		}
		// ---- glue: lane e handles event e (at most 4)
		if (ev_ok && lane < nev && st) { // slots were handed out before the DP status was known: mark them unusable
			idl_event_result ev; memset(&ev, 0, sizeof ev);
			ev.aln = ai; ev.index = lane; ev.reject = IDL_EV_DP_ERROR; ev.min_flank = -1; ev.amq_median = ev.rmq_median = -1;
			g.eres[eoff + lane] = ev;
		}
		if (ev_ok && lane < nev && !st) {
			idl_event_result ev; memset(&ev, 0, sizeof ev);
			ev.aln = ai; ev.index = lane; ev.min_flank = -1; ev.amq_median = ev.rmq_median = -1; ev.ref_code = ev.alt_code = ~0ULL;
			int toff = 0, qoff = 0, seen = 0; // target_locations / query_locations, src/ksw2/ksw2.nim:71-91
			for (int k = 0; k < ntr; ++k) {
				const uint32_t c = d.cig[n - 1 - k]; const int op = c & 0xf, len = (int)(c >> 4);
				if (op != 0) {
					if (seen == lane) {
						ev.len = len;
						if (op == 1) { ev.type = 0; ev.t_start = cr.start + toff; ev.t_stop = ev.t_start + 1; ev.q_start = qoff; ev.q_stop = qoff + len; }
						else { ev.type = 1; ev.t_start = cr.start + toff; ev.t_stop = ev.t_start + len; ev.q_start = qoff; ev.q_stop = qoff + 1; }
					}
					++seen;
				}
				if (op != 1) toff += len;
				if (op != 2) qoff += len;
			}
			const int clen = cr.len;
			do {
				if (ev.len < P.min_event_len) { ev.reject = IDL_EV_SHORT; break; } // :234
				int tstart = ev.t_start - cr.start - width; if (tstart < 0) tstart = 0; // :236-238
				if (tstart + K > tlen) tstart = tlen - K;
				ev.tstart = tstart;
				if (tstart < 0) { ev.reject = IDL_EV_WINDOW; break; }
				int off = clen - ev.q_stop - 1; if (ev.q_start < off) off = ev.q_start; // :243
				ev.offset = off;
				int qstart = ev.q_start - width; if (qstart < 0) qstart = 0;              // :244-246
				if (qstart + K > clen) qstart = clen - K;
				ev.qstart = qstart;
				if (qstart < 0) { ev.reject = IDL_EV_WINDOW; break; }
				bool same = same_k(tq + tstart, qq + qstart, K);
				if (same) { // :255-262
					qstart = ev.q_start - 3; if (qstart < 0) qstart = 0;
					if (qstart + K > clen) { int qend = ev.q_stop + 4; if (qend > clen) qend = clen; qstart = qend - K; }
					ev.qstart = qstart;
					if (qstart < 0) { ev.reject = IDL_EV_WINDOW; break; }
					same = same_k(tq + tstart, qq + qstart, K);
				}
				if (same && (ev.q_start == 0 || distinct_k(qq + qstart, K) == 1)) { ev.reject = IDL_EV_SAME_KMER; break; } // :264
				if (distinct_k(tq + tstart, K) < 3) { ev.reject = IDL_EV_LOW_CPLX; break; }                                 // :266
				if (same) { ev.reject = IDL_EV_BUG_SAME; break; }                                                            // :268-275
				ev.ref_code = canon_code(tq + tstart, K); ev.alt_code = canon_code(qq + qstart, K);
				// get_min_flank(qloc, ez), :118-132, over the truncated CIGAR
				{
					long long result = 0x7fffffffffffffffLL; bool found = false; int mf = 0;
					for (int k = 0; k < ntr; ++k) {
						const uint32_t c = d.cig[n - 1 - k]; const int op = c & 0xf; const long long len = c >> 4;
						if (op == 0) {
							result = found ? (len < result ? len : result) : len;
							if (found) { mf = (int)result; break; }
						} else if (op - 1 == ev.type && len == ev.len) {
							if (result == 0x7fffffffffffffffLL) result = 0;
							found = true;
						}
					}
					ev.min_flank = mf;
				}
			} while (0);
			g.eres[eoff + lane] = ev;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// kernel 3: k-mer counting, one CTA per event
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KMER_THREADS) kmer_kernel(GenoArgs g)
{
	__shared__ unsigned hist_a[256], hist_r[256];
	__shared__ unsigned long long s_sum_a, s_sum_r, s_bytes;
	__shared__ unsigned s_ka, s_kr, s_kb, s_reads;
	const int tid = threadIdx.x;
	const idl_params &P = g.P;
	const int K = IDL_KMER;
	const uint64_t kmask = (1ULL << (2 * K)) - 1ULL;
	const unsigned n_events = g.cnt->n_events < g.cap_events ? g.cnt->n_events : g.cap_events;
	for (unsigned e = blockIdx.x; e < n_events; e += gridDim.x) {
		idl_event_result ev = g.eres[e];
		if (ev.reject != IDL_EV_COUNTED) continue; // uniform
		const idl_aln_result ar = g.ares[ev.aln];
		const idl_region R = g.region[ar.region];
		__syncthreads();
		for (int i = tid; i < 256; i += KMER_THREADS) { hist_a[i] = 0; hist_r[i] = 0; }
		if (tid == 0) { s_sum_a = s_sum_r = s_bytes = 0; s_ka = s_kr = s_kb = s_reads = 0; }
		__syncthreads();
		const uint64_t refe = ev.ref_code, alte = ev.alt_code;
		for (unsigned j = tid; j < R.n_reads; j += KMER_THREADS) { // :293-311
			const idl_read rd = g.read[R.read_begin + j];
			if ((int)rd.mapq < P.count_min_mapq) continue; // :294
			const int L = rd.len;
			bool rf = false, af = false; int rdist = 0, adist = 0;
			const uint4 *p2 = (const uint4*)(g.seq2 + (rd.seq_off >> 4)); // 64 bases per 128-bit load, records are 16-byte aligned
			const uint2 *pn = (const uint2*)(g.seqn + (rd.seq_off >> 5));
			uint64_t f = 0, rc = 0; int valid = 0;
			for (int blk = 0; blk * 64 < L && !(rf && af); ++blk) {
				const uint4 w4 = __ldg(p2 + blk);
				const uint2 n2 = __ldg(pn + blk);
				const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
				const uint64_t nmask = (uint64_t)n2.x | ((uint64_t)n2.y << 32);
				const int lim = L - blk * 64 < 64 ? L - blk * 64 : 64;
				for (int i = 0; i < lim; ++i) {
					const unsigned b = (ws[i >> 4] >> (2 * (i & 15))) & 3u;
					if ((nmask >> i) & 1ULL) { valid = 0; f = rc = 0; continue; } // a window holding a non-ACGT base never matches
					f = ((f << 2) | b) & kmask;
					rc = (rc >> 2) | ((uint64_t)(3u - b) << (2 * (K - 1)));
					if (++valid < K) continue;
					const uint64_t c = f < rc ? f : rc;
					if (c == refe || c == alte) {
						const int pos = blk * 64 + i - K + 1;
						const int dd = pos < (L - K) - pos ? pos : (L - K) - pos; // declared kmer.dists distance
						if (!rf && c == refe) { rf = true; rdist = dd; }
						if (!af && c == alte) { af = true; adist = dd; }
					}
				}
			}
			atomicAdd(&s_reads, 1u);
			atomicAdd(&s_bytes, (unsigned long long)((L + 3) / 4 + (L + 7) / 8 + 16));
			if (rf) { atomicAdd(&s_kr, 1u); atomicAdd(&s_sum_r, (unsigned long long)rdist); atomicAdd(&hist_r[rd.mapq], 1u); }
			if (af) { atomicAdd(&s_ka, 1u); atomicAdd(&s_sum_a, (unsigned long long)adist); atomicAdd(&hist_a[rd.mapq], 1u); }
			if (rf && af) atomicAdd(&s_kb, 1u);
		}
		__syncthreads();
		if (tid == 0) {
			ev.k_ref = (int)s_kr; ev.k_alt = (int)s_ka; ev.k_both = (int)s_kb;
			ev.n_adist = (int)s_ka; ev.n_rdist = (int)s_kr; ev.sum_adist = (long long)s_sum_a; ev.sum_rdist = (long long)s_sum_r;
			// median(): sorted[int(len/2)], src/indelope.nim:152-155
			for (int which = 0; which < 2; ++which) {
				const unsigned *h = which ? hist_r : hist_a; const unsigned nn = which ? s_kr : s_ka;
				int med = -1;
				if (nn) { unsigned acc = 0; const unsigned target = nn / 2; for (int q = 0; q < 256; ++q) { acc += h[q]; if (acc > target) { med = q; break; } } }
				if (which) ev.rmq_median = med; else ev.amq_median = med;
			}
			if (s_kb > 0) { // :313: genotype by alignment instead
				ev.aligned = 1; ev.ref_support = ev.alt_support = ev.both_found = 0;
				const unsigned long long old = atomicAdd(&g.cnt->al_pack, (1ULL << 40) | (unsigned long long)R.n_reads);
				const unsigned slot = (unsigned)(old >> 40);
				if (slot < g.cap_al) { AlEntry a; a.event = e; a.base = (unsigned)(old & ((1ULL << 40) - 1)); a.n_reads = R.n_reads; a.pad = 0; g.al_list[slot] = a; }
				else atomicOr(&g.cnt->overflow, 32u);
				atomicAdd(&g.cnt->al_events, 1ULL);
			} else { ev.aligned = 0; ev.ref_support = ev.k_ref; ev.alt_support = ev.k_alt; ev.both_found = 0; }
			g.eres[e] = ev;
			atomicAdd(&g.cnt->kmer_reads, (unsigned long long)s_reads);
			atomicAdd(&g.cnt->kmer_bytes, s_bytes + 64ULL);
		}
	}
}

// count_flanked_cigar (src/indelope.nim:185-199) over the truncated view of the reversed scratch
__device__ __forceinline__ int count_flanked(const uint32_t *cig_rev, int n, int max_q)
{
	const int ntr = ksw_trunc_count(cig_rev, n, max_q);
	bool matched = false; int cnt = 0, last_op = 0;
	for (int k = 0; k < ntr; ++k) {
		const int op = cig_rev[n - 1 - k] & 0xf;
		if (!matched) { if (op == 0) { cnt += 1; matched = true; } }
		else cnt += 1;
		last_op = op;
	}
	if (last_op != 0) cnt -= 1;
	return cnt;
}

// ---------------------------------------------------------------------------------------------------------------
// AL fallback: one warp per (event, read) work item
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DP_THREADS, 3) al_kernel(GenoArgs g)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DpWarp d = dp_carve(g, smem_raw);
	const int lane = lane_id();
	const idl_params &P = g.P;
	KswParams kp; kp.match = (int8_t)P.match; kp.mismatch = (int8_t)P.mismatch; kp.q = (int8_t)P.b_gapo; kp.e = (int8_t)P.b_gape; kp.w = P.b_bw; kp.zdrop = P.b_zdrop;
	const unsigned long long pack = g.cnt->al_pack;
	unsigned n_al = (unsigned)(pack >> 40); if (n_al > g.cap_al) n_al = g.cap_al;
	const unsigned total = (unsigned)(pack & ((1ULL << 40) - 1));
	for (;;) {
		unsigned it = 0;
		if (lane == 0) it = atomicAdd(&g.cnt->al_next, 1u);
		it = __shfl_sync(FULL_MASK, it, 0);
		if (it >= total) break;
		if (n_al == 0) break;
		// entries are sorted by base (one packed atomic hands out slot and base together)
		int lo = 0, hi = (int)n_al - 1;
		while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (g.al_list[mid].base <= it) lo = mid; else hi = mid - 1; }
		const AlEntry ae = g.al_list[lo];
		const unsigned j = it - ae.base;
		if (j >= ae.n_reads) continue; // item belongs to a dropped (overflowed) entry
		const idl_event_result ev = g.eres[ae.event];
		const idl_aln_result ar = g.ares[ev.aln];
		const idl_contig_result cr = g.cres[ar.contig];
		const idl_region R = g.region[ar.region];
		const idl_read rd = g.read[R.read_begin + j];
		if ((int)rd.mapq < P.count_min_mapq) continue;              // :328
		const int rs = rd.start + rd.trim_a;                       // :331
		if (rs > ev.t_stop) continue;                              // :332
		const int Lx = ev.type == 0 ? ev.len : 0;                  // :333-335
		if (rs + (int)rd.trim_len + Lx < ev.t_start) continue;     // :336
		const int start = (rs > cr.start ? rs : cr.start) - cr.start; // :339
		const int qlen = rd.trim_len;
		if (qlen > g.qcap) { if (lane == 0) atomicOr(&g.rres[ar.region].status, IDL_RS_READ_TOO_LONG); continue; }
		// unpack the trimmed read to 0..4 codes (src/ksw2/ksw2.nim:127-132)
		for (int i = lane; i < qlen; i += 32) {
			const unsigned b = rd.seq_off + rd.trim_a + i;
			const unsigned isn = (g.seqn[b >> 5] >> (b & 31)) & 1u;
			d.qbuf[i] = isn ? 4 : (uint8_t)((g.seq2[b >> 4] >> (2 * (b & 15))) & 3u);
		}
		__syncwarp();
		const uint8_t *refw = g.refcodes + R.ref_off + (cr.start - R.ref_start);
		int rlen = ar.ref_len - start; if (rlen < 0) rlen = 0;
		int clen = cr.len - start; if (clen < 0) clen = 0;
		KswOut o;
		dp_run(g, d, qlen, d.qbuf, rlen, refw + start, kp, o);          // read_seq.align_to(ref_sub, ez_ref) :343
		unsigned st = dp_status_bits(o.status);
		int rn = 0, an = 0;
		if (lane == 0) rn = count_flanked(d.cig, o.n_cigar, o.max_q);
		unsigned long long cells = (unsigned long long)o.cells;
		__syncwarp();
		dp_run(g, d, qlen, d.qbuf, clen, g.ctg_codes + cr.seq_off + start, kp, o); // read_seq.align_to(ctg_sub, ez_alt) :344
		st |= dp_status_bits(o.status);
		cells += (unsigned long long)o.cells;
		if (lane == 0) {
			an = count_flanked(d.cig, o.n_cigar, o.max_q);
			if (st) atomicOr(&g.rres[ar.region].status, st);
			else if (rn == 1 && an > 1) atomicAdd(&g.eres[ae.event].ref_support, 1);   // :353-356
			else if (an == 1 && rn > 1) atomicAdd(&g.eres[ae.event].alt_support, 1);
			atomicAdd(&g.cnt->dp_b, 2ULL); atomicAdd(&g.cnt->dp_cells_b, cells);
		}
		__syncwarp();
	}
}
