// indelope_b200/csrc/assemble.cuh -- kernel 1: slide-and-vote assembler, one CTA per region of interest.
//
// Replaces assemble (src/indelope.nim:157-183) over src/contig.nim: slide_align (:70-141), best_match (:224-240),
// insert (:156-222), Contig.trim (:49-68) and the two-pass combine (:254-281).
//
// Data layout.  A contig lives in a "slot": three bit-planes (low code bit, high code bit, non-ACGT flag), 32 bases
// per 32-bit word, plus one uint16 support counter per base.  Slots sit in a per-CTA arena in global memory (the
// working set of a region is a few KB and stays in L1); the query of the current best_match is staged in shared
// memory.  The overlap scan gives every (contig, offset) pair to one lane: the overlap is compared 32 bases at a time
// as  (q0 ^ t0) | (q1 ^ t1) | (qn ^ tn)  with a funnel shift aligning the shifted side, aborting at the first
// word with a mismatch the voting rule does not allow; matches = overlap - popc(allowed mismatches).  The arg-max uses
// the reference's exact order: most matches, then lowest contig index (stable sort, :239), then first offset in scan
// order (:107,135).  Reads are consumed strictly in input order -- the greedy merge is order dependent.
//
// A left-overhang merge (offset < 0, :180-205) is built in the QUERY's slot (after corrections both sides agree on the
// overlap, so only the target's tail has to be appended), which avoids shifting the target; the list entry is then
// re-pointed to that slot.
#pragma once
#include "common.cuh"

#ifndef ASM_BLOCKS
#define ASM_BLOCKS 1          // warp variant: screen (contig, 32-offset block) pairs position-parallel
#endif
#ifndef ASM_EVAL_NOINLINE
#define ASM_EVAL_NOINLINE 0
#endif
#ifndef ASM_SCREEN_UNROLL
#define ASM_SCREEN_UNROLL 4
#endif
#if ASM_EVAL_NOINLINE
#define ASM_EVAL_ATTR __noinline__
#else
#define ASM_EVAL_ATTR
#endif
#define ASM_THREADS 256        // threads per CTA of both variants
#define ASM_SMALL_READS 126    // regions up to this many reads are assembled by ONE WARP (no block barriers); larger ones by a CTA
#define ASM_SMALL_NS (ASM_SMALL_READS + 2)

// NT = threads cooperating on one region: 32 (a warp: 8 regions per CTA, __syncwarp only) or 256 (the whole CTA).
template <int NT> __device__ __forceinline__ void asm_bar() { if (NT == 32) __syncwarp(); else __syncthreads(); }
template <int NT> __device__ __forceinline__ int asm_tid() { return NT == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }

struct AsmArgs {
	// batch (device copies)
	const idl_region *region; const idl_read *read;
	const uint32_t *seq2, *seqn, *ref2, *refn;
	unsigned n_regions, n_reads, n_seq_bases;
	idl_params P;
	// arena: n_ctas * ns slots
	uint32_t *planes; uint16_t *sup; int ns, nw, cap;
	// outputs
	idl_region_result *rres; idl_contig_result *cres; idl_aln_result *ares;
	char *ctg_ascii; uint8_t *ctg_codes; uint32_t *ctg_sup; uint8_t *refcodes;
	unsigned cap_contigs, cap_bases, cap_alns;
	SortBufs sortA;
	DevCounters *cnt;
	int small;                        // 1: this launch takes the regions with <= ASM_SMALL_READS reads, 0: the others
	const unsigned *order;            // regions by read count, deepest first: the work of a region grows faster than its reads,
	                                  // so the persistent grid ends on the cheapest regions instead of idling behind a deep one
};

struct AsmS { // carved out of dynamic shared memory
	uint32_t *q0, *q1, *qn;           // staged query planes, nw words each
	int *len, *nreads, *start;        // per slot
	uint16_t *listA, *listB, *freestk;
	uint16_t *blkcum;                 // warp variant: exclusive prefix of the screen blocks per list entry
	int *corr;                        // 3 ints per correction site: qoff, toff, qbest
	unsigned long long *best;         // per warp
	int *sc;                          // scalars: see SC_*
};
enum { SC_NFREE = 0, SC_STATUS, SC_TMP0, SC_TMP1, SC_NCORR, SC_REGION, SC_HASN, SC_SLOT, SC_N };

// shared memory of one cooperating group (a warp or a CTA) for regions of up to ns-2 reads
__host__ __device__ inline size_t asm_smem_bytes(int ns, int nw, int nt)
{
	size_t b = (size_t)3 * nw * 4 + (size_t)3 * ns * 4 + (size_t)4 * ns * 2 + (size_t)3 * IDL_MAX_CORRECTIONS * 4 + (size_t)(nt / 32) * 8 + SC_N * 4 + 16;
	return (b + 15) & ~(size_t)15;
}

struct Asm {
	AsmS s; const AsmArgs *a;
	unsigned pbase, sbase; int nw, cap, ns;  // this group's arena as 32-bit element offsets into a->planes / a->sup
	unsigned long long offsets;
	// 32-bit index arithmetic on top of the kernel-parameter base pointers: the arenas stay below 2^32 elements, and these
	// accessors sit in every inner loop of an instruction-cache-bound kernel
	__device__ uint32_t *p0(int slot) const { return a->planes + (pbase + (unsigned)slot * (unsigned)(3 * nw)); }
	__device__ uint32_t *p1(int slot) const { return a->planes + (pbase + (unsigned)slot * (unsigned)(3 * nw) + (unsigned)nw); }
	__device__ uint32_t *pn(int slot) const { return a->planes + (pbase + (unsigned)slot * (unsigned)(3 * nw) + 2u * (unsigned)nw); }
	__device__ uint16_t *sup(int slot) const { return a->sup + (sbase + (unsigned)slot * (unsigned)cap); }
};

// allowable_mismatch, src/contig.nim:44-47 (uint32 products on the supports, int on the read counts)
__device__ __forceinline__ bool asm_allowed(unsigned qsup, unsigned tsup, int qreads, int treads)
{
	return (qsup < 3u && tsup > 3u * qsup && qreads > 3 * (int)qsup) || (tsup < 3u && qsup > 3u * tsup && treads > 3 * (int)tsup);
}

__device__ __forceinline__ uint32_t compress_even(uint64_t x) // bits 0,2,4,... -> 32 bits
{
	x &= 0x5555555555555555ULL;
	x = (x | (x >> 1)) & 0x3333333333333333ULL;
	x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0fULL;
	x = (x | (x >> 4)) & 0x00ff00ff00ff00ffULL;
	x = (x | (x >> 8)) & 0x0000ffff0000ffffULL;
	x = (x | (x >> 16)) & 0x00000000ffffffffULL;
	return (uint32_t)x;
}

template <int NT> __device__ int asm_alloc(Asm &A) // uniform: every thread gets the same slot
{
	asm_bar<NT>();
	if (asm_tid<NT>() == 0) {
		int n = A.s.sc[SC_NFREE];
		if (n > 0) { A.s.sc[SC_SLOT] = A.s.freestk[n - 1]; A.s.sc[SC_NFREE] = n - 1; }
		else { A.s.sc[SC_SLOT] = -1; A.s.sc[SC_STATUS] |= IDL_RS_CONTIG_OVERFLOW; }
	}
	asm_bar<NT>();
	return A.s.sc[SC_SLOT];
}
template <int NT> __device__ void asm_free(Asm &A, int slot) // call from uniform code; takes effect at the next barrier
{
	if (asm_tid<NT>() == 0) { A.s.freestk[A.s.sc[SC_NFREE]] = (uint16_t)slot; A.s.sc[SC_NFREE] += 1; }
}

// one (query, contig, offset) candidate: number of matching bases, or -1 if a mismatch is not allowed.
// dir2 == false: loop 1 of slide_align (:86-111), q[i] against t[o+i]; dir2 == true: loop 2 (:114-139), q[o+i] against t[i].
__device__ ASM_EVAL_ATTR int asm_eval(const uint32_t *sq0, const uint32_t *sq1, const uint32_t *sqn, const uint32_t *t0, const uint32_t *t1, const uint32_t *tn, const uint16_t *tsup, const uint16_t *qsup,
                        int qlen, int tlen, int qreads, int treads, bool dir2, int o, bool vote, bool has_n)
{
	const int n = dir2 ? min(qlen - o, tlen) : min(qlen, tlen - o);
	if (n <= 0) return 0; // nothing compared: ma = 0, mm = 0
	int ncorr = 0;
	const int nwords = (n + 31) >> 5;
	#pragma unroll 1
	for (int w = 0; w < nwords; ++w) {
		uint32_t m;
		const int pos = o + 32 * w;
		if (!dir2) {
			m = (sq0[w] ^ get32p(t0, pos)) | (sq1[w] ^ get32p(t1, pos));
			if (has_n) m |= sqn[w] ^ get32p(tn, pos);
		} else {
			m = (t0[w] ^ get32p(sq0, pos)) | (t1[w] ^ get32p(sq1, pos));
			if (has_n) m |= tn[w] ^ get32p(sqn, pos);
		}
		const int rem = n - 32 * w;
		if (rem < 32) m &= (1u << rem) - 1u;
		if (m) {
			if (!vote) return -1;
			while (m) {
				const int b = __ffs(m) - 1; m &= m - 1;
				const int i = 32 * w + b;
				const int qo = dir2 ? o + i : i, to = dir2 ? i : o + i;
				if (!asm_allowed(qsup[qo], tsup[to], qreads, treads)) return -1;
				++ncorr;
			}
		}
	}
	return n - ncorr;
}

struct AsmMatch { int k, offset, ma; bool aligned; };

// Exact-overlap screen of 32 consecutive offsets at once.  text0/text1(/textn) hold 64 bits of the sliding side's planes
// starting at the block's first offset, pat0/pat1(/patn) the first word of the fixed side: bit b of the result is set when
// the 16 bases the fixed side starts with equal the sliding side's bases b .. b+15 (position-parallel compare: one funnel
// shift per plane and base instead of one compare per offset).
// Kept as a short loop (the kernel lives or dies by the instruction cache), the three-plane variant out of line.
template <bool HAS_N>
__device__ __forceinline__ uint32_t asm_screen32_t(uint32_t pat0, uint32_t pat1, uint32_t patn, uint32_t ta0, uint32_t tb0, uint32_t ta1, uint32_t tb1,
                                                   uint32_t tan, uint32_t tbn)
{
	uint32_t m = 0xffffffffu;
	const uint32_t np0 = ~pat0, np1 = ~pat1, npn = ~patn;
	constexpr int UNR = ASM_SCREEN_UNROLL;
#pragma unroll UNR
	for (int j = 0; j < 16; ++j) {
		const uint32_t a = (uint32_t)((int32_t)(np0 << (31 - j)) >> 31), b = (uint32_t)((int32_t)(np1 << (31 - j)) >> 31); // ~0 where the pattern bit is 0
		m &= (__funnelshift_r(ta0, tb0, j) ^ a) & (__funnelshift_r(ta1, tb1, j) ^ b);
		if (HAS_N) m &= __funnelshift_r(tan, tbn, j) ^ (uint32_t)((int32_t)(npn << (31 - j)) >> 31);
	}
	return m;
}
__device__ __noinline__ uint32_t asm_screen32_n(uint32_t pat0, uint32_t pat1, uint32_t patn, uint32_t ta0, uint32_t tb0, uint32_t ta1, uint32_t tb1, uint32_t tan, uint32_t tbn)
{
	return asm_screen32_t<true>(pat0, pat1, patn, ta0, tb0, ta1, tb1, tan, tbn);
}

// best_match (:224-240) of slot q against list[0..nlist). Uniform result.
// One lane per (contig, offset) candidate, a warp per contig.  A pair that cannot vote (fewer than 4 reads on either
// side: every read of the greedy pass) needs an exact overlap, so a candidate is first screened on its first 32 bases
// with two funnel shifts; only survivors run the full compare.  Every lane keeps its best key over all contigs
// (matches, then lowest contig index, then first offset in scan order), one reduction per call finds the winner.
template <int NT> __device__ AsmMatch asm_best_match(Asm &A, const uint16_t *list, int nlist, int q, int mo, bool has_n)
{
	const int tid = asm_tid<NT>(), lane = tid & 31, warp = tid >> 5;
	const int qlen = A.s.len[q], qreads = A.s.nreads[q];
	const int n2 = qlen - mo >= 0 ? qlen - mo : mo - qlen; // abs(omin), :78,114
	asm_bar<NT>();
	{ // stage the query planes (zero padded two words past the end: the shifted reads of loop 2 run that far)
		const int qw = (qlen + 31) >> 5;
		const uint32_t *g0 = A.p0(q), *g1 = A.p1(q), *gn = A.pn(q);
#pragma unroll 1
		for (int w = tid; w < qw + 2 && w < A.nw; w += NT) {
			const bool in = w < qw;
			A.s.q0[w] = in ? g0[w] : 0u; A.s.q1[w] = in ? g1[w] : 0u; A.s.qn[w] = in ? gn[w] : 0u;
		}
	}
	asm_bar<NT>();
	const uint16_t *qsup = A.sup(q);
	const uint32_t qf0 = A.s.q0[0], qf1 = A.s.q1[0], qfn = A.s.qn[0];
	unsigned long long best = 0, tested = 0;
	// Warp variant: every (contig, 32-offset block) pair of the contigs that cannot vote becomes one lane task, so one
	// pass of the warp screens 32 blocks = up to 1024 offsets of many contigs at once; the few offsets that survive the
	// 16-base screen run the full compare.  Needs mo - 1 >= 16: an overlap shorter than 16 bases then cannot qualify.
	const bool blocks = ASM_BLOCKS && NT == 32 && mo >= 17;
	if (blocks) {
		const int nb2 = (n2 >> 5) + 1; // loop 2 offsets are 1 .. n2: bits 1 .. n2 of the blocks 0 .. n2 / 32
		unsigned carry = 0;
#pragma unroll 1
		for (int k0 = 0; k0 < nlist; k0 += 32) {
			const int k = k0 + lane;
			unsigned nb = 0;
			if (k < nlist) {
				const int t = list[k];
				const bool vote = qreads >= 4 && A.s.nreads[t] >= 4;
				if (!vote) { const int omax = A.s.len[t] - mo; nb = (unsigned)((omax >= 0 ? (omax >> 5) + 1 : 0) + nb2); }
			}
			unsigned inc = nb;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(FULL_MASK, inc, d); if (lane >= d) inc += o; }
			if (k < nlist) A.s.blkcum[k] = (uint16_t)(carry + inc - nb);
			carry += __shfl_sync(FULL_MASK, inc, 31);
		}
		__syncwarp();
		const int total = (int)carry;
#pragma unroll 1
		for (int task = lane; task < total; task += 32) {
			int lo = 0, hi = nlist - 1; // the last entry whose first block is <= task (entries without blocks share their successor's start)
			while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int)A.s.blkcum[mid] <= task) lo = mid; else hi = mid - 1; }
			const int k = lo, t = list[k], tlen = A.s.len[t];
			const int omax = tlen - mo, n1 = omax >= 0 ? omax + 1 : 0, nb1 = omax >= 0 ? (omax >> 5) + 1 : 0;
			int B = task - (int)A.s.blkcum[k];
			const bool dir2 = B >= nb1;
			if (dir2) B -= nb1;
			const uint32_t *t0 = A.p0(t), *t1 = t0 + A.nw, *tn = t1 + A.nw;
			// loop 1: the contig slides under the query's first bases; loop 2: the query under the contig's
			const uint32_t *x0 = dir2 ? A.s.q0 : t0, *x1 = dir2 ? A.s.q1 : t1;
			const uint32_t pat0 = dir2 ? t0[0] : qf0, pat1 = dir2 ? t1[0] : qf1;
			uint32_t m;
			if (!has_n) m = asm_screen32_t<false>(pat0, pat1, 0u, x0[B], x0[B + 1], x1[B], x1[B + 1], 0u, 0u);
			else { const uint32_t *xn = dir2 ? A.s.qn : tn; m = asm_screen32_n(pat0, pat1, dir2 ? tn[0] : qfn, x0[B], x0[B + 1], x1[B], x1[B + 1], xn[B], xn[B + 1]); }
			// offsets of this block that exist (:86,114) and overlap by at least the 16 screened bases
			const int o0 = 32 * B;
			int last = dir2 ? min(n2, min(qlen - 16, qlen - 1)) : min(n1 - 1, tlen - 16); // largest offset worth a look
			if (dir2 && tlen < 16) last = -1;
			if (!dir2 && qlen < 16) last = -1;
			const int hi_b = last - o0; // bits 0 .. hi_b
			if (hi_b < 0) m = 0; else if (hi_b < 31) m &= (2u << hi_b) - 1u;
			if (dir2 && B == 0) m &= ~1u; // offset 0 belongs to loop 1
			const unsigned long long kkey = (unsigned long long)(0xffff - k) << 16;
			while (m) {
				const int b = __ffs(m) - 1; m &= m - 1;
				const int o = o0 + b;
				const int ma = asm_eval(A.s.q0, A.s.q1, A.s.qn, t0, t1, tn, A.sup(t), qsup, qlen, tlen, qreads, 0, dir2, o, false, has_n);
				if (ma >= 0 && ma >= mo - 1) {
					const int sidx = dir2 ? n1 + o - 1 : o;
					const unsigned long long k64 = ((unsigned long long)(unsigned)(ma + 1) << 32) | kkey | (unsigned long long)(0xffff - sidx);
					best = k64 > best ? k64 : best;
				}
			}
		}
	}
#pragma unroll 1
	for (int k = warp; k < nlist; k += (NT / 32)) {
		const int t = list[k];
		const int tlen = A.s.len[t];
		const int omax = tlen - mo;
		const int n1 = omax >= 0 ? omax + 1 : 0;
		bool vote = false; // a vote needs support >= 4 on one side and reads >= 4 on the other (supports are 1..nreads)
		if (qreads >= 4) vote = A.s.nreads[t] >= 4;
		tested += (unsigned long long)(n1 + n2);
		if (blocks && !vote) continue; // screened above
		const uint32_t *t0 = A.p0(t), *t1 = t0 + A.nw, *tn = t1 + A.nw;
		const uint32_t tf0 = t0[0], tf1 = t1[0], tfn = has_n ? tn[0] : 0u;
		const unsigned long long kkey = (unsigned long long)(0xffff - k) << 16;
#pragma unroll 1
		for (int sidx = lane; sidx < n1 + n2; sidx += 32) {
			const bool dir2 = sidx >= n1;
			const int o = dir2 ? sidx - n1 + 1 : sidx;
			const int n = dir2 ? min(qlen - o, tlen) : min(qlen, tlen - o);
			int ma = -1;
			bool full = vote || n <= 0 || (ASM_BLOCKS && NT == 32); // the warp variant screens by blocks above; what is left for this loop compares in full
			if (!full) { // exact overlap needed: screen the first word
				uint32_t m;
				const int w = o >> 5, sh = o & 31;
				if (!dir2) {
					m = (qf0 ^ __funnelshift_r(t0[w], t0[w + 1], sh)) | (qf1 ^ __funnelshift_r(t1[w], t1[w + 1], sh));
					if (has_n) m |= qfn ^ __funnelshift_r(tn[w], tn[w + 1], sh);
				} else {
					m = (tf0 ^ __funnelshift_r(A.s.q0[w], A.s.q0[w + 1], sh)) | (tf1 ^ __funnelshift_r(A.s.q1[w], A.s.q1[w + 1], sh));
					if (has_n) m |= tfn ^ __funnelshift_r(A.s.qn[w], A.s.qn[w + 1], sh);
				}
				if (n < 32) m &= (1u << n) - 1u;
				if (!m) { if (n <= 32) ma = n; else full = true; }
			}
			if (full) ma = asm_eval(A.s.q0, A.s.q1, A.s.qn, t0, t1, tn, A.sup(t), qsup, qlen, tlen, qreads, vote ? A.s.nreads[t] : 0, dir2, o, vote, has_n);
			// first candidate needs ma >= mo-1 (best_ma starts at mo-1, best_mm at 1: :81-82,107); later ones strictly more
			if (ma >= 0 && ma >= mo - 1) {
				const unsigned long long k64 = ((unsigned long long)(unsigned)(ma + 1) << 32) | kkey | (unsigned long long)(0xffff - sidx);
				best = k64 > best ? k64 : best;
			}
		}
	}
#pragma unroll
	for (int d = 16; d >= 1; d >>= 1) { const unsigned long long o = __shfl_xor_sync(FULL_MASK, best, d); best = o > best ? o : best; }
	if (NT > 32) {
		if (lane == 0) A.s.best[warp] = best;
		asm_bar<NT>();
		best = 0;
		for (int w2 = 0; w2 < (NT / 32); ++w2) { const unsigned long long b = A.s.best[w2]; best = b > best ? b : best; }
	}
	if (lane == 0) A.offsets += tested;
	AsmMatch m;
	m.aligned = best != 0;
	m.ma = (int)(best >> 32) - 1;
	m.k = 0xffff - (int)((best >> 16) & 0xffff);
	const int sidx = 0xffff - (int)(best & 0xffff);
	m.offset = 0;
	if (m.aligned) {
		const int tlen = A.s.len[list[m.k]];
		const int omax = tlen - mo;
		const int n1 = omax >= 0 ? omax + 1 : 0;
		m.offset = sidx < n1 ? sidx : -(sidx - n1 + 1);
	}
	return m;
}

// Contig.insert (:156-222): merge slot q into list[k] at m.offset. Returns the slot that now holds the merged contig.
template <int NT> __device__ void asm_merge(Asm &A, uint16_t *list, int k, int q, int offset, bool has_n)
{
	const int tid = asm_tid<NT>();
	const int t = list[k];
	const int tlen = A.s.len[t], qlen = A.s.len[q], treads = A.s.nreads[t], qreads = A.s.nreads[q];
	uint16_t *tsup = A.sup(t), *qsup = A.sup(q);
	uint32_t *tp[3] = {A.p0(t), A.p1(t), A.pn(t)}, *qp[3] = {A.p0(q), A.p1(q), A.pn(q)};
	const bool vote = qreads >= 4 && treads >= 4;
	asm_bar<NT>();
	// 1. corrections of the winning offset, in position order (:93-99), then applied (:161-173)
	if (tid == 0) {
		int nc = 0;
		if (vote) {
			const bool dir2 = offset < 0; const int o = dir2 ? -offset : offset;
			const int n = dir2 ? min(qlen - o, tlen) : min(qlen, tlen - o);
			#pragma unroll 1
			for (int w = 0; w * 32 < n; ++w) {
				const int pos = o + 32 * w;
				uint32_t m;
				if (!dir2) m = (qp[0][w] ^ get32p(tp[0], pos)) | (qp[1][w] ^ get32p(tp[1], pos)) | (qp[2][w] ^ get32p(tp[2], pos));
				else m = (tp[0][w] ^ get32p(qp[0], pos)) | (tp[1][w] ^ get32p(qp[1], pos)) | (tp[2][w] ^ get32p(qp[2], pos));
				const int rem = n - 32 * w;
				if (rem < 32) m &= (1u << rem) - 1u;
				while (m) {
					const int b = __ffs(m) - 1; m &= m - 1;
					const int i = 32 * w + b;
					const int qo = dir2 ? o + i : i, to = dir2 ? i : o + i;
					if (nc < IDL_MAX_CORRECTIONS) { A.s.corr[3 * nc] = qo; A.s.corr[3 * nc + 1] = to; A.s.corr[3 * nc + 2] = qsup[qo] > tsup[to]; }
					++nc;
				}
			}
			if (nc > IDL_MAX_CORRECTIONS) { A.s.sc[SC_STATUS] |= IDL_RS_CORR_OVERFLOW; nc = IDL_MAX_CORRECTIONS; }
			#pragma unroll 1
			for (int c = 0; c < nc; ++c) {
				const int qo = A.s.corr[3 * c], to = A.s.corr[3 * c + 1];
				uint32_t **dst = A.s.corr[3 * c + 2] ? tp : qp, **src = A.s.corr[3 * c + 2] ? qp : tp;
				const int di = A.s.corr[3 * c + 2] ? to : qo, si = A.s.corr[3 * c + 2] ? qo : to;
				for (int pl = 0; pl < 3; ++pl) {
					const uint32_t bit = (src[pl][si >> 5] >> (si & 31)) & 1u;
					dst[pl][di >> 5] = (dst[pl][di >> 5] & ~(1u << (di & 31))) | (bit << (di & 31));
				}
				if (A.s.corr[3 * c + 2]) tsup[to] = qsup[qo]; else qsup[qo] = tsup[to];
			}
		}
		A.s.sc[SC_NCORR] = nc;
	}
	asm_bar<NT>();
	const int nc = A.s.sc[SC_NCORR];
	if (offset < 0) { // :180-205, built in q's slot (q frame)
		const int ao = -offset;
		const int newlen = max(qlen, ao + tlen);
		if (newlen > A.cap) { if (tid == 0) A.s.sc[SC_STATUS] |= IDL_RS_CONTIG_OVERFLOW; asm_bar<NT>(); return; }
		#pragma unroll 1
		for (int i = ao + tid; i < ao + tlen; i += NT) {
			unsigned val = tsup[i - ao];
			if (i < qlen) {
				bool dont = false;
				#pragma unroll 1
				for (int c = 0; c < nc; ++c) dont |= A.s.corr[3 * c] == i; // dont_overwrite holds qoff here (:170-171)
				if (!dont) val += qsup[i];
			}
			qsup[i] = (uint16_t)val;
		}
		if (ao + tlen > qlen) { // append the target's tail: bases [qlen, ao+tlen) come from t[i-ao]
			#pragma unroll 1
			for (int w = (qlen >> 5) + tid; w * 32 < newlen; w += NT) {
				const uint32_t low = qlen - 32 * w >= 32 ? 0xffffffffu : (qlen > 32 * w ? (1u << (qlen - 32 * w)) - 1u : 0u);
				for (int pl = 0; pl < 3; ++pl) {
					if (pl == 2 && !has_n) { qp[2][w] = 0; continue; }
					const uint32_t tb = get32(tp[pl], 32 * w - ao);
					qp[pl][w] = (qp[pl][w] & low) | (tb & ~low);
				}
			}
		}
		asm_bar<NT>();
		if (tid == 0) {
			A.s.len[q] = newlen; A.s.nreads[q] = treads + qreads; // start stays q.start (:204)
			list[k] = (uint16_t)q;
		}
		asm_free<NT>(A, t);
	} else { // :210-222, in t's slot
		const int o = offset;
		const int newlen = max(tlen, o + qlen);
		if (newlen > A.cap) { if (tid == 0) A.s.sc[SC_STATUS] |= IDL_RS_CONTIG_OVERFLOW; asm_bar<NT>(); return; }
		#pragma unroll 1
		for (int i = o + tid; i < o + qlen; i += NT) {
			if (i < tlen) {
				bool dont = false;
				#pragma unroll 1
				for (int c = 0; c < nc; ++c) dont |= A.s.corr[3 * c + 1] == i; // toff (:172-173)
				if (!dont) tsup[i] = (uint16_t)(tsup[i] + qsup[i - o]);
			} else tsup[i] = qsup[i - o];
		}
		if (o + qlen > tlen) {
			#pragma unroll 1
			for (int w = (tlen >> 5) + tid; w * 32 < newlen; w += NT) {
				const uint32_t low = tlen - 32 * w >= 32 ? 0xffffffffu : (tlen > 32 * w ? (1u << (tlen - 32 * w)) - 1u : 0u);
				for (int pl = 0; pl < 3; ++pl) {
					if (pl == 2 && !has_n) { tp[2][w] = 0; continue; }
					const uint32_t qb = get32(qp[pl], 32 * w - o);
					tp[pl][w] = (tp[pl][w] & low) | (qb & ~low);
				}
			}
		}
		asm_bar<NT>();
		if (tid == 0) { A.s.len[t] = newlen; A.s.nreads[t] = treads + qreads; }
		asm_free<NT>(A, q);
	}
	asm_bar<NT>();
}

// Contig.trim (:49-68). Returns the slot holding the trimmed contig (a fresh one if bases were dropped on the left).
template <int NT> __device__ int asm_trim(Asm &A, int c, int min_support, bool has_n)
{
	const int tid = asm_tid<NT>();
	const int L = A.s.len[c];
	const unsigned ms = (unsigned)min_support;
	const uint16_t *sp = A.sup(c);
	asm_bar<NT>();
	if (tid == 0) { A.s.sc[SC_TMP0] = 0x7fffffff; A.s.sc[SC_TMP1] = -1; }
	asm_bar<NT>();
	#pragma unroll 1
	for (int i = tid; i < L - 1; i += NT) if (sp[i] >= ms) { atomicMin(&A.s.sc[SC_TMP0], i); break; }
	asm_bar<NT>();
	int a = A.s.sc[SC_TMP0];
	if (a > L - 1) a = L - 1 > 0 ? L - 1 : 0;
	if (a >= L - 1) { // :56-60
		asm_bar<NT>();
		if (tid == 0) { A.s.start[c] += a; A.s.len[c] = 0; A.s.nreads[c] = 0; }
		asm_bar<NT>();
		return c;
	}
	#pragma unroll 1
	for (int i = L - 1 - tid; i > a; i -= NT) if (sp[i] >= ms) { atomicMax(&A.s.sc[SC_TMP1], i); break; }
	asm_bar<NT>();
	int b = A.s.sc[SC_TMP1];
	if (b < a) b = a;
	const int newlen = b - a + 1;
	if (a == 0) {
		asm_bar<NT>();
		if (tid == 0) A.s.len[c] = newlen;
		asm_bar<NT>();
		return c;
	}
	const int d = asm_alloc<NT>(A);
	if (d < 0) return c;
	const uint32_t *s0 = A.p0(c), *s1 = A.p1(c), *sn = A.pn(c);
	uint32_t *d0 = A.p0(d), *d1 = A.p1(d), *dn = A.pn(d);
	uint16_t *dsup = A.sup(d);
	#pragma unroll 1
	for (int w = tid; w * 32 < newlen; w += NT) {
		d0[w] = get32p(s0, a + 32 * w); d1[w] = get32p(s1, a + 32 * w); dn[w] = has_n ? get32p(sn, a + 32 * w) : 0u;
	}
	#pragma unroll 1
	for (int i = tid; i < newlen; i += NT) dsup[i] = sp[a + i];
	asm_bar<NT>();
	if (tid == 0) { A.s.len[d] = newlen; A.s.nreads[d] = A.s.nreads[c]; A.s.start[d] = A.s.start[c] + a; }
	asm_free<NT>(A, c);
	asm_bar<NT>();
	return d;
}

#ifndef ASM_CTAS
#define ASM_CTAS 3 /* resident CTAs per SM: 80 registers and no spills (4 CTAs at 64 registers spilled 68 bytes in the warp variant: 7.18 -> 6.59 ms on chr1; 2 CTAs 7.85) */
#endif
template <int NT>
__global__ void __launch_bounds__(ASM_THREADS, ASM_CTAS) assemble_kernel(AsmArgs args)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Asm A;
	A.a = &args; A.nw = args.nw; A.cap = args.cap; A.ns = args.ns; A.offsets = 0;
	const int grp_in_cta = NT == 32 ? (int)(threadIdx.x >> 5) : 0;
	const size_t grp = (size_t)blockIdx.x * (ASM_THREADS / NT) + grp_in_cta; // this group's arena
	A.pbase = (unsigned)(grp * args.ns * 3 * args.nw);
	A.sbase = (unsigned)(grp * args.ns * args.cap);
	{ // carve shared memory
		// plain pointer arithmetic from the shared array only (an integer round trip, e.g. to align, would turn every later
		// access into a generic 64-bit load): the 8-byte entries come first, the group's region is a multiple of 16 bytes
		unsigned char *p = smem_raw + asm_smem_bytes(args.ns, args.nw, NT) * grp_in_cta;
		A.s.best = (unsigned long long*)p; p += (size_t)(NT / 32) * 8;
		A.s.q0 = (uint32_t*)p; p += (size_t)A.nw * 4; A.s.q1 = (uint32_t*)p; p += (size_t)A.nw * 4; A.s.qn = (uint32_t*)p; p += (size_t)A.nw * 4;
		A.s.len = (int*)p; p += (size_t)A.ns * 4; A.s.nreads = (int*)p; p += (size_t)A.ns * 4; A.s.start = (int*)p; p += (size_t)A.ns * 4;
		A.s.corr = (int*)p; p += (size_t)3 * IDL_MAX_CORRECTIONS * 4;
		A.s.sc = (int*)p; p += SC_N * 4;
		A.s.listA = (uint16_t*)p; p += (size_t)A.ns * 2; A.s.listB = (uint16_t*)p; p += (size_t)A.ns * 2; A.s.freestk = (uint16_t*)p; p += (size_t)A.ns * 2;
		A.s.blkcum = (uint16_t*)p;
	}
	const int tid = asm_tid<NT>();
	const idl_params &P = args.P;
	#pragma unroll 1
	for (;;) {
		asm_bar<NT>();
		if (tid == 0) A.s.sc[SC_REGION] = (int)atomicAdd(args.small ? &args.cnt->region_next : &args.cnt->region_next2, 1u);
		asm_bar<NT>();
		const unsigned qi = (unsigned)A.s.sc[SC_REGION];
		if (qi >= args.n_regions) break;
		const unsigned rg = args.order[qi];
		const idl_region R = args.region[rg];
		if ((R.n_reads <= ASM_SMALL_READS) != (args.small != 0)) continue; // the other launch assembles this region
		// reset the slot allocator; detect non-ACGT bases in this region's reads and window
		if (tid == 0) {
			A.s.sc[SC_NFREE] = A.ns; A.s.sc[SC_HASN] = 0;
			A.s.sc[SC_STATUS] = (R.n_reads + 2 > (unsigned)A.ns ? IDL_RS_CONTIG_OVERFLOW : 0) | ((R.flags & IDL_RF_READ_TOO_LONG) ? IDL_RS_READ_TOO_LONG : 0);
		}
		#pragma unroll 1
		for (int i = tid; i < A.ns; i += NT) A.s.freestk[i] = (uint16_t)(A.ns - 1 - i);
		asm_bar<NT>();
		{
			// the read records are validated here, by the first kernel that touches them (the host checks only the region records):
			// a record that points outside the pools or whose trim range leaves the read drops the region (IDL_RS_BAD_INPUT)
			int any = 0, bad = 0;
			#pragma unroll 1
			for (unsigned j = 0; j < R.n_reads; ++j) {
				const idl_read rd = args.read[R.read_begin + j];
				if ((rd.seq_off & 63u) || (unsigned long long)rd.seq_off + (((unsigned)rd.len + 63u) & ~63u) > (unsigned long long)args.n_seq_bases ||
				    (unsigned)rd.trim_a + (unsigned)rd.trim_len > (unsigned)rd.len || (int)rd.len > P.max_read_len) { bad = 1; continue; }
				const uint32_t *pn = args.seqn + (rd.seq_off >> 5);
				#pragma unroll 1
				for (int w = tid; w * 32 < rd.len; w += NT) any |= pn[w] != 0;
			}
			if (any) A.s.sc[SC_HASN] = 1;
			if (bad && tid == 0) A.s.sc[SC_STATUS] |= IDL_RS_BAD_INPUT; // every thread saw the same records
		}
		// unpack the reference window to 0..4 codes for kernel 2 / glue (src/ksw2/ksw2.nim:127-132)
		#pragma unroll 1
		for (unsigned i = tid; i < R.ref_len; i += NT) {
			const unsigned b = R.ref_off + i;
			const unsigned isn = (args.refn[b >> 5] >> (b & 31)) & 1u;
			args.refcodes[b] = isn ? 4 : (uint8_t)((args.ref2[b >> 4] >> (2 * (b & 15))) & 3u);
		}
		asm_bar<NT>();
		const bool has_n = A.s.sc[SC_HASN] != 0;
		// ---- assemble (src/indelope.nim:163-169) and the two passes of combine (:176 -> src/contig.nim:254-281) as ONE loop:
		// every step takes an item q (phase 0: the next read, in input order, made into a contig; phases 1, 2: the next
		// contig of the previous list), finds its best match in the list under construction and merges or appends it.
		// best_match / merge / trim are instantiated once, which keeps the kernel inside the instruction cache.
		uint16_t *list = A.s.listA, *in = A.s.listB;
		int nlist = 0, n_pre = 0, n_in = 0, usedi = -1, i_in = 0, phase = 0;
		unsigned j = 0;
		#pragma unroll 1
		for (;;) {
			int q = -1, mo = 0;
			if (phase == 0) {
				bool got = false;
				while (j < R.n_reads && !A.s.sc[SC_STATUS]) {
					const idl_read rd = args.read[R.read_begin + j];
					++j;
					if ((int)rd.mapq < P.asm_min_mapq) continue;  // :164
					if (rd.flags & 1) continue;                  // :165
					q = asm_alloc<NT>(A);
					if (q < 0) break;
					const int tl = rd.trim_len;
					if (tl > A.cap) { if (tid == 0) A.s.sc[SC_STATUS] |= IDL_RS_CONTIG_OVERFLOW; asm_bar<NT>(); break; }
					{ // make_contig (:143-150) from the packed, trimmed read
						uint32_t *g0 = A.p0(q), *g1 = A.p1(q), *gn = A.pn(q);
						uint16_t *gs = A.sup(q);
						const unsigned base = rd.seq_off + rd.trim_a;
						#pragma unroll 1
						for (int w = tid; w * 32 < tl; w += NT) {
							const unsigned b = base + 32u * w;         // first base of this plane word
							const unsigned wi = b >> 4, sh = 2 * (b & 15);
							const uint32_t w0 = args.seq2[wi], w1 = args.seq2[wi + 1], w2 = args.seq2[wi + 2];
							const uint64_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
							const uint64_t bits = lo | (hi << 32);     // 32 bases, 2 bits each
							uint32_t m = 0xffffffffu;
							if (tl - 32 * w < 32) m = (1u << (tl - 32 * w)) - 1u;
							g0[w] = compress_even(bits) & m;
							g1[w] = compress_even(bits >> 1) & m;
							gn[w] = has_n ? (get32p(args.seqn, (int)b) & m) : 0u;
						}
						#pragma unroll 1
						for (int i = tid; i < tl; i += NT) gs[i] = 1;
						if (tid == 0) { A.s.len[q] = tl; A.s.nreads[q] = 1; A.s.start[q] = rd.start + rd.trim_a; }
					}
					asm_bar<NT>();
					mo = rd.min_overlap;
					got = true;
					break;
				}
				if (!got) { n_pre = nlist; phase = 1; i_in = n_in = 0; usedi = -1; } // :171; the first combine pass starts below
			}
			if (phase > 0 && !(i_in < n_in)) { // a pass of combine is over (or the greedy pass was): start the next one
				if (phase == 3 || A.s.sc[SC_STATUS]) break;
				const int min_support = phase == 1 ? 0 : P.combine_min_support; // pass A merges without trimming (:260), pass B trims (:265-267)
				{ uint16_t *t = list; list = in; in = t; } n_in = nlist; nlist = 0;
				usedi = -1;
				#pragma unroll 1
				for (int i = 0; i < n_in; ++i) {
					int c = in[i];
					if (min_support > 0) {
						const int nr = A.s.nreads[c];
						const int c2 = asm_trim<NT>(A, c, nr < min_support ? nr : min_support, has_n);
						if (c2 != c) { asm_bar<NT>(); if (tid == 0) in[i] = (uint16_t)c2; asm_bar<NT>(); c = c2; }
					}
					if (usedi < 0 && A.s.nreads[c] > 0) usedi = i;
				}
				++phase; // 2: pass A running, 3: pass B running
				i_in = 0;
				if (usedi < 0) { n_in = 0; continue; } // nothing left: the remaining pass is empty as well
				asm_bar<NT>();
				if (tid == 0) list[0] = in[usedi];
				asm_bar<NT>();
				nlist = 1;
				continue;
			}
			if (phase > 0) {
				if (i_in == usedi) { ++i_in; continue; }
				if (A.s.sc[SC_STATUS]) break;
				q = in[i_in++]; mo = P.combine_min_overlap;
			}
			const AsmMatch m = asm_best_match<NT>(A, list, nlist, q, mo, has_n);
			if (m.aligned) asm_merge<NT>(A, list, m.k, q, m.offset, has_n);
			else if (phase == 0 || A.s.nreads[q] > 0) { asm_bar<NT>(); if (tid == 0) list[nlist] = (uint16_t)q; asm_bar<NT>(); ++nlist; }
			else { asm_free<NT>(A, q); asm_bar<NT>(); }
		}
		asm_bar<NT>();
		// ---- results
		const unsigned status = (unsigned)A.s.sc[SC_STATUS];
		if (status) nlist = 0;
		if (tid == 0) {
			unsigned total = 0;
			#pragma unroll 1
			for (int i = 0; i < nlist; ++i) total += (unsigned)((A.s.len[list[i]] + 3) & ~3);
			A.s.sc[SC_TMP0] = (int)atomicAdd(&args.cnt->n_contigs, (unsigned)nlist);
			A.s.sc[SC_TMP1] = (int)atomicAdd(&args.cnt->n_contig_bases, total);
		}
		asm_bar<NT>();
		const unsigned cbegin = (unsigned)A.s.sc[SC_TMP0];
		unsigned boff = (unsigned)A.s.sc[SC_TMP1];
		if (tid == 0) {
			idl_region_result rr; rr.status = status | ((R.flags & IDL_RF_ALPHABET) ? IDL_RS_ALPHABET : 0u); rr.n_contigs_pre = n_pre; rr.n_contigs = nlist; rr.contig_begin = cbegin;
			args.rres[rg] = rr;
		}
		if (cbegin + (unsigned)nlist > args.cap_contigs) { if (tid == 0) atomicOr(&args.cnt->overflow, 1u); continue; }
		#pragma unroll 1
		for (int i = 0; i < nlist; ++i) {
			const int c = list[i];
			const int L = A.s.len[c];
			if (boff + (unsigned)L > args.cap_bases) { if (tid == 0) atomicOr(&args.cnt->overflow, 2u); break; }
			const uint32_t *g0 = A.p0(c), *g1 = A.p1(c), *gn = A.pn(c);
			const uint16_t *gs = A.sup(c);
			#pragma unroll 1
			for (int x = tid; x < L; x += NT) {
				const unsigned code = ((g0[x >> 5] >> (x & 31)) & 1u) | (((g1[x >> 5] >> (x & 31)) & 1u) << 1);
				const bool isn = has_n && ((gn[x >> 5] >> (x & 31)) & 1u);
				args.ctg_codes[boff + x] = isn ? 4 : (uint8_t)code;
				args.ctg_ascii[boff + x] = isn ? 'N' : "ACGT"[code];
				if (args.ctg_sup) args.ctg_sup[boff + x] = gs[x];
			}
			if (tid == 0) {
				idl_contig_result cr; cr.start = A.s.start[c]; cr.nreads = A.s.nreads[c]; cr.len = L; cr.seq_off = boff; cr.aln = -1; cr.region = rg;
				// gates of src/indelope.nim:209-211
				if ((P.stages & IDL_STAGE_ALIGN) && n_pre <= P.max_contigs && cr.nreads >= P.min_reads && L >= P.min_ctg_len) {
					const unsigned ai = atomicAdd(&args.cnt->n_alns, 1u);
					if (ai < args.cap_alns) {
						cr.aln = (int)ai;
						idl_aln_result ar; memset(&ar, 0, sizeof ar);
						ar.region = rg; ar.contig = cbegin + i;
						// reference window of :213-220: fai.get(chrom, ctg.start, max_stop + width + 50), clipped to the shipped window
						const int win_end = R.ref_start + (int)R.ref_len - 1;
						const int max_stop = cr.start > R.max_stop ? cr.start : R.max_stop;
						int end = max_stop + P.window_pad; if (end > win_end) end = win_end;
						int tlen = end - cr.start + 1;
						if (tlen < 0 || cr.start < R.ref_start) tlen = 0;
						ar.ref_len = tlen;
						args.ares[ai] = ar;
						const uint16_t key = sort_key_a(est_diagonals(L, tlen, P.a_bw));
						args.sortA.keys[ai] = key;
						atomicAdd(&args.sortA.hist[key], 1u);
					} else atomicOr(&args.cnt->overflow, 4u);
				}
				args.cres[cbegin + i] = cr;
			}
			boff += (unsigned)((L + 3) & ~3);
		}
	}
	if ((threadIdx.x & 31) == 0 && A.offsets) atomicAdd(&args.cnt->offsets_tested, A.offsets);
}
