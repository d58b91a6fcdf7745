// indelope_b200/csrc/ksw2.cuh -- kernel 2: ksw2 extension alignment, a sub-warp GROUP of G threads per alignment.
//
// Replaces ksw_extz2_sse (src/ksw2/csrc/ksw2_extz2_sse.c:113-388, flag == 0) bit for bit: the anti-diagonal
// difference recurrence on wrapping int8 lanes, the 16-lane rounding of the band and the stale lanes it leaves
// behind (they feed real cells), the exact 32-bit max with the SSE 4-accumulator tie
// order, z-drop, and ksw_backtrack (:47-79) to a BAM-style CIGAR.
//
// Mapping.  32/G alignments run side by side in one warp, in lockstep (the per-diagonal control code is issued once
// for all of them).  Inside a group every thread owns W consecutive packed words = 4*W consecutive columns of the
// anti-diagonal: u, v, x, y, s are int8x4 words (the SSE code holds sixteen lanes per register), wrapping add/sub and
// signed/unsigned compares are the byte-SIMD intrinsics, the x[t-1]/v[t-1] neighbour is the thread's previous word or
// comes from the previous thread through __shfl_up_sync + a one-byte funnel shift.  One fused pass per diagonal does
// the 16-lane score blocks (:215-228), the core update (:262-284) and the exact-H update (:312-349) with the columns in
// registers.  The lane arrays live in shared memory as a RING over the live band (columns left of the band are dead;
// the 16-lane block that enters the band is cleared just before, exactly reproducing calloc'ed memory and the stale
// values the SSE code leaves behind), H is a ring of the exact band, the backtrack matrix goes to global memory and is
// walked through a 32x32 shared-memory tile.  All arithmetic is integer; no tensor cores.
#pragma once
#include "common.cuh"

struct KswOut {
	int max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score;
	int n_cigar;      // ops left in the scratch buffer, REVERSED (last op first)
	int status;       // 0 ok, 1 early return (:147/:171), negative: capacity
	long long cells;  // exact in-band cells over executed diagonals
};

#define KSW_BTILE_BYTES 1024
#define KSW_PMAT_PAD 64   /* bytes in front of every backtrack matrix (the tile prefetch may start before row 0) */
#define KSW_ST_EARLY 1
#define KSW_ST_RCAP (-2)
#define KSW_ST_PCAP (-3)
#define KSW_ST_CIGCAP (-4)
#define KSW_ST_HCAP (-5)
#define KSW_ST_SEQCAP (-6)

__device__ __forceinline__ void ksw_reset(KswOut &o) // ksw_reset_extz :81-86
{
	o.max_q = o.max_t = o.mqe_t = o.mte_q = -1;
	o.max = 0; o.score = o.mqe = o.mte = KSW_NEG_INF;
	o.n_cigar = 0; o.zdropped = 0; o.status = 0; o.cells = 0;
}

// widest rounded band (bytes of one backtrack row) for a query/target/bandwidth, :164-165
__host__ __device__ inline int ksw_ncol(int qlen, int tlen, int w)
{
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	int n = qlen < tlen ? qlen : tlen;
	n = n < w + 1 ? n : w + 1;
	return ((n + 15) / 16 + 1) * 16;
}
// ring of lane columns: the rounded band, the column left of it, 16 lanes of score overrun and the 16 being cleared
__host__ __device__ inline int ksw_ring_cols(int ncol) { int r = 64; while (r < ncol + 48) r <<= 1; return r; }
__host__ __device__ inline int ksw_h_ring(int ncol) { int r = 16; while (r < ncol + 8) r <<= 1; return r; } // ncol-16 >= exact band
// bytes of sequence staging: zero-padded target (sf) and reversed, zero-padded query (qr); doubles as the backtrack tile
__host__ __device__ inline size_t ksw_seq_bytes(int qlen, int tlen)
{
	size_t b = (size_t)(((tlen + 15) & ~15) + 16) + (size_t)((qlen + 35) & ~3);
	return b < KSW_BTILE_BYTES ? KSW_BTILE_BYTES : b;
}

// off[r] / off_end[r] of the reference are pure functions of r (:196-199,205)
__device__ __forceinline__ void ksw_band(int r, int qlen, int tlen, int w, int &st0, int &en0)
{
	int st = 0, en = tlen - 1;
	if (st < r - qlen + 1) st = r - qlen + 1;
	if (en > r) en = r;
	if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
	if (en > ((r + w) >> 1)) en = (r + w) >> 1;
	st0 = st; en0 = en;
}

// ---- four int8 lanes per 32-bit register ----
__device__ __forceinline__ uint32_t sel4(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); } // m ? a : b, per byte (one LOP3)
__device__ __forceinline__ uint32_t rep4(int v) { return (uint32_t)(v & 0xff) * 0x01010101u; }
// 0xff in every byte whose top bit is set (PRMT with the sign-replicate selector bit; __byte_perm() masks that bit off)
__device__ __forceinline__ uint32_t msb_to_mask4(uint32_t v)
{
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(0u), "r"(0xba98u));
	return r;
}

// tie order of the SSE arg-max (:316-348): 4 strided accumulators (lower accumulator, then lower t), then the scalar tail
__device__ __forceinline__ unsigned ksw_tie_rank(int t, int st0, int en1) { return 1u + ((t < en1 ? (unsigned)((t - st0) & 3) : 4u) << 20) + (unsigned)(t - st0); }

// where the query of an alignment comes from: 0..4 codes, or a 2-bit packed read (+ non-ACGT plane) of the batch
struct KswQuery { const uint8_t *codes; const uint32_t *seq2, *seqn; unsigned base; };
__device__ __forceinline__ uint8_t ksw_query_code(const KswQuery &q, int i)
{
	if (q.codes) return q.codes[i];
	const unsigned b = q.base + (unsigned)i;
	return ((q.seqn[b >> 5] >> (b & 31)) & 1u) ? (uint8_t)4 : (uint8_t)((q.seq2[b >> 4] >> (2 * (b & 15))) & 3u);
}

// Memory of one group: lanes = 5 rings of ring_cols bytes, H = hr ints, seq = seq_cap bytes (reused as the backtrack
// tile) in shared memory; pmat = backtrack matrix workspace, cig = CIGAR scratch in global memory
struct KswMem { int8_t *lanes; int ring_cols; int *H; int hr; uint8_t *seq; int seq_cap; uint8_t *pmat; size_t p_cap; uint32_t *cig; int cig_cap; };
__host__ __device__ inline size_t ksw_group_smem(int ring_cols, int hr, int seq_cap)
{
	return ((size_t)5 * ring_cols + 128 + (size_t)hr * 4 + (size_t)((seq_cap + 15) & ~15) + 127) & ~(size_t)127;
}
// layout of a group's region (a multiple of 128 bytes): [lanes: 5 rings + 128 bytes of stagger slack][H][seq]
__host__ __device__ inline size_t ksw_group_h_off(int ring_cols) { return (size_t)5 * ring_cols + 128; }
// Bank staggering: the G = 8 threads of a group touch 8 consecutive words, the 32/G groups of a warp touch the same relative
// word.  This byte offset (added to the lane rings only; H and seq keep their 16-byte alignment) makes the 32 words of one
// access fall into 32 different banks.
__host__ __device__ inline int ksw_group_stagger(int grp_in_warp, int W) { (void)W; return 32 * grp_in_warp; } // the 8 threads of a group touch 8 consecutive words

// The G threads of a group call this with identical arguments; `out` comes back identical in each of them.
template <int G, int W>
__device__ void ksw2_group(int qlen, const KswQuery query, int tlen, const uint8_t *target, KswParams P, const KswMem M, KswOut &out)
{
	const int lane = lane_id();
	const int gl = lane & (G - 1);
	const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
	ksw_reset(out);
	if (qlen <= 0 || tlen <= 0) { out.status = KSW_ST_EARLY; return; } // :147
	const int qe = P.q + P.e;
	{
		int min_sc = P.mismatch < 0 ? P.mismatch : 0;
		if (P.match < min_sc) min_sc = P.match;
		if (-min_sc > 2 * qe) { out.status = KSW_ST_EARLY; return; } // :171
	}
	int w = P.w;
	if (w < 0) w = tlen > qlen ? tlen : qlen; // :161
	const int T16 = (tlen + 15) & ~15;
	const int n_col = ksw_ncol(qlen, tlen, w); // :164-165 (bytes)
	if (ksw_ring_cols(n_col) > M.ring_cols) { out.status = KSW_ST_RCAP; return; }
	if (n_col + 8 > M.hr) { out.status = KSW_ST_HCAP; return; }
	if (ksw_seq_bytes(qlen, tlen) > (size_t)M.seq_cap) { out.status = KSW_ST_SEQCAP; return; }
	if ((size_t)(qlen + tlen - 1) * (size_t)n_col + 2 * KSW_PMAT_PAD > M.p_cap) { out.status = KSW_ST_PCAP; return; }
	uint8_t *pmat = M.pmat + KSW_PMAT_PAD;
	const int hmask = M.hr - 1, rm = M.ring_cols - 1, rmw = (M.ring_cols >> 2) - 1;
	int *H = M.H; int4 *H4 = (int4*)M.H;
	int8_t *u = M.lanes, *v = u + M.ring_cols, *x = v + M.ring_cols, *y = x + M.ring_cols, *s = y + M.ring_cols;
	uint32_t *U = (uint32_t*)u, *V = (uint32_t*)v, *X = (uint32_t*)x, *Y = (uint32_t*)y, *S = (uint32_t*)s;
	uint8_t *sf = M.seq, *qr = M.seq + T16 + 16;
	const uint32_t *SF = (const uint32_t*)sf, *QR = (const uint32_t*)qr;
	const uint32_t QE2 = rep4(qe * 2), MAXSC = rep4(P.match + qe * 2), Q4 = rep4(P.q), MATQ = rep4(P.match + qe * 2), MISQ = rep4(P.mismatch + qe * 2);

	// calloc :173: columns [0,16) of u,v,x,y and [0,32) of s start as zero (later blocks are cleared as they enter);
	// stage sf (target, zero padded) and qr (reversed query, zero padded) exactly as :187-188 lay them out
	for (int i = gl; i < 16 / 4; i += G) { U[i] = 0; V[i] = 0; X[i] = 0; Y[i] = 0; }
	for (int i = gl; i < 32 / 4; i += G) S[i] = QE2; // S holds s + 2(q+e) (:117): one add less per word and diagonal
	bool wild = false; // a code 4 anywhere: only then the score needs the wildcard mask (:219,226)
	for (int i = gl; i < T16 + 16; i += G) { const uint8_t c = i < tlen ? target[i] : (uint8_t)0; wild |= c == 4; sf[i] = c; }
	{
		const int nq = (qlen + 35) & ~3;
		for (int i = gl; i < nq; i += G) { const uint8_t c = i < qlen ? ksw_query_code(query, qlen - 1 - i) : (uint8_t)0; wild |= c == 4; qr[i] = c; }
	}
	wild = __ballot_sync(gmask, wild) & gmask;
	__syncwarp(gmask);

	int last_st = -1, last_en = -1, en_clr = 15;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st0, en0;
		ksw_band(r, qlen, tlen, w, st0, en0);
		if (st0 > en0) { out.zdropped = 1; break; } // :200-203
		const int st = st0 & ~15, en = en0 | 15;    // :205
		out.cells += en0 - st0 + 1;
		if (en > en_clr) { // a 16-lane block enters the band: it must read as never written (calloc); the score overrun zone moves on
			const int b = en_clr + 1;
			for (int i = gl; i < 8; i += G) {
				if (i < 4) { const int wi = ((b >> 2) + i) & rmw; U[wi] = 0; V[wi] = 0; X[wi] = 0; Y[wi] = 0; }
				else S[(((b + 16) >> 2) + i - 4) & rmw] = QE2;
			}
			en_clr += 16;
			__syncwarp(gmask);
		}
		// boundary conditions :207-211 (values of the previous diagonal)
		int x1, v1;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) { x1 = x[(st - 1) & rm]; v1 = v[(st - 1) & rm]; }
			else x1 = v1 = 0;
		} else { x1 = 0; v1 = r ? P.q : 0; }
		const int bend = st0 + ((en0 - st0) / 16 + 1) * 16; // one past the last lane the 16-wide score blocks write
		const int w1 = (bend - 1) >> 2, wend = en >> 2, wlast = wend > w1 ? wend : w1, ws0 = st0 >> 2;
		const int en1 = st0 + (((en0 - st0) >> 2) << 2);    // end of the 4-wide vector part of the arg-max (:316)
		uint32_t xc = (uint32_t)(x1 & 0xff) << 24, vc = (uint32_t)(v1 & 0xff) << 24;
		// H[en0] is built from the OLD H[en0-1] (:318): fetch it before this diagonal's updates
		const int hprev_old = r == 0 ? 0 : (en0 > 0 ? H[(en0 - 1) & hmask] : H[en0 & hmask]);
		const unsigned bandw = (unsigned)(en0 - st0);       // columns st0 .. en0-1 are updated in the loop
		uint32_t *pr = (uint32_t*)(pmat + (size_t)r * n_col) - (st >> 2);
		int bh = (int)0x80000000, bt = st0;                  // this thread's best exact score and its column (ties in SSE order)
		const int rword = en >= r ? (r >> 2) : -1;           // :212, y[r] = 0 and u[r] = q, patched in registers
		// Words are dealt round-robin: word j of the band goes to thread j % G, so a narrow band (the first and last
		// diagonals, short targets) costs ceil(words / G) rounds instead of always W.
		for (int w0 = st >> 2; w0 <= wlast; w0 += G * W) {
			const int wb = w0 + gl;
			uint32_t xo[W], vo[W], xp[W], vp[W];
#pragma unroll
			for (int k = 0; k < W; ++k) {
				const bool core = wb + k * G <= wend;
				xo[k] = core ? X[(wb + k * G) & rmw] : 0u;
				vo[k] = core ? V[(wb + k * G) & rmw] : 0u;
			}
			// neighbour word (columns t-4..t-1 of the previous diagonal): the previous thread in the same round, or thread G-1 of
			// the round before for thread 0 (the carry of the previous trip / the boundary value for the very first word)
#pragma unroll
			for (int k = 0; k < W; ++k) {
				xp[k] = __shfl_up_sync(gmask, xo[k], 1, G); vp[k] = __shfl_up_sync(gmask, vo[k], 1, G);
				const uint32_t xw = __shfl_sync(gmask, xo[k ? k - 1 : 0], G - 1, G), vw = __shfl_sync(gmask, vo[k ? k - 1 : 0], G - 1, G);
				if (gl == 0) { xp[k] = k ? xw : xc; vp[k] = k ? vw : vc; }
			}
			xc = __shfl_sync(gmask, xo[W - 1], G - 1, G); vc = __shfl_sync(gmask, vo[W - 1], G - 1, G);
#pragma unroll
			for (int k = 0; k < W; ++k) {
				if (w0 + k * G > wlast) break; // the rest of this trip lies beyond the band (same for the whole group)
				const int wi = wb + k * G, t = wi << 2;
				const bool core = wi <= wend, sca = wi >= ws0 && wi <= w1;
				uint32_t so = (core || sca) ? S[wi & rmw] : 0u;
				if (sca) { // scores: lanes of this word inside the 16-wide blocks get fresh values, the others keep stale s
					const uint32_t sq = SF[wi];
					const int p = qlen - 1 - r + t; // qr index of lane t; negative only for lanes left of st0 (masked below)
					uint32_t sq2;
					if (p >= 0) sq2 = __funnelshift_r(QR[p >> 2], QR[(p >> 2) + 1], 8 * (p & 3));
					else sq2 = p > -4 ? QR[0] << (8 * -p) : 0u;
					const uint32_t neq = msb_to_mask4((sq ^ sq2) + 0x7f7f7f7fu);                                  // 0xff where the codes differ
					uint32_t sc = sel4(neq, MISQ, MATQ);
					if (wild) sc = sel4(msb_to_mask4(((sq ^ 0x04040404u) + 0x7f7f7f7fu) & ((sq2 ^ 0x04040404u) + 0x7f7f7f7fu)), sc, QE2); // score 0 where a code is 4
					const int lo = st0 - t, hi = bend - t; // lanes [lo, hi) of this word belong to the blocks
					uint32_t m = 0xffffffffu;
					if (lo > 0) m &= 0xffffffffu << (8 * lo);
					if (hi < 4) m &= 0xffffffffu >> (8 * (4 - hi));
					so = sel4(m, sc, so);
					S[wi & rmw] = so;
				}
				if (core) {
					uint32_t ut = U[wi & rmw], yt = Y[wi & rmw];
					if (wi == rword) {
						const int b = 8 * (r & 3);
						yt &= ~(0xffu << b);
						ut = (ut & ~(0xffu << b)) | ((uint32_t)((r ? P.q : 0) & 0xff) << b);
					}
					const uint32_t xt1 = __funnelshift_l(xp[k], xo[k], 8), vt1 = __funnelshift_l(vp[k], vo[k], 8); // lanes t-1..t+2 of the previous diagonal
					uint32_t z = so;                               // s + 2(q+e)
					uint32_t a = __vadd4(xt1, vt1);
					uint32_t b = __vadd4(yt, ut);
					uint32_t m = __vcmpgts4(a, z);                 // a > z (signed)
					uint32_t d = m & 0x01010101u;
					z = sel4(m, a, z);                             // signed max
					m = __vcmpgts4(b, z);                          // b > z (signed)
					d = sel4(m, 0x02020202u, d);
					z = sel4(__vcmpgtu4(b, z), b, z);              // unsigned max
					z = sel4(__vcmpgtu4(z, MAXSC), MAXSC, z);      // unsigned min
					const uint32_t vn = __vsub4(z, ut);
					U[wi & rmw] = __vsub4(z, vt1); V[wi & rmw] = vn;
					z = __vsub4(z, Q4);
					a = __vsub4(a, z);
					b = __vsub4(b, z);
					m = __vcmpgts4(a, 0u);
					X[wi & rmw] = a & m; d |= m & 0x08080808u;
					m = __vcmpgts4(b, 0u);
					Y[wi & rmw] = b & m; d |= m & 0x10101010u;
					pr[wi] = d;
					if (t + 3 >= st0 && t < en0) { // exact H of the in-band columns st0 .. en0-1 of this word (:323-348): H[t] += v8[t] - qe
						const int4 hq = H4[(t & hmask) >> 2];
						int hv[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
						for (int c = 0; c < 4; ++c) {
							const int tc = t + c;
							const bool in = (unsigned)(tc - st0) < bandw;
							const int h = hv[c] + (int)((vn >> (8 * c)) & 0xff) - qe;
							const bool gt = in && h > bh;
							if (in && h == bh && ksw_tie_rank(tc, st0, en1) < ksw_tie_rank(bt, st0, en1)) bt = tc; // rare
							hv[c] = in ? h : hv[c];
							bt = gt ? tc : bt;
							bh = gt ? h : bh;
						}
						H4[(t & hmask) >> 2] = make_int4(hv[0], hv[1], hv[2], hv[3]);
					}
				}
			}
		}
		__syncwarp(gmask); // this diagonal's lanes and H[st0..en0) are visible to the whole group
		// the en0 cell, :318 / :349
		int hen;
		if (r == 0) hen = (int)(uint8_t)v[0] - qe - qe;
		else hen = hprev_old + (int)(uint8_t)(en0 > 0 ? u[en0 & rm] : v[en0 & rm]) - qe;
		int max_H, max_t;
		{
			const unsigned mk = __reduce_max_sync(gmask, (unsigned)bh ^ 0x80000000u);
			const int mh = (int)(mk ^ 0x80000000u);
			if (hen >= mh) { max_H = hen; max_t = en0; } // the initial candidate (H[en0], en0) wins every tie (:318-321)
			else { // several threads may hold the max: the SSE tie order decides
				max_H = mh;
				const unsigned rk = __reduce_min_sync(gmask, bh == mh ? ksw_tie_rank(bt, st0, en1) : 0xffffffffu);
				max_t = st0 + (int)((rk - 1) & 0xfffffu);
			}
		}
		if (en0 == tlen - 1) {
			if (hen > out.mte) { out.mte = hen; out.mte_q = r - en; }
			if (r == qlen + tlen - 2) out.score = hen; // H[tlen-1]
		}
		if (r - st0 == qlen - 1) {
			const int Hst0 = st0 == en0 ? hen : H[st0 & hmask];
			if (Hst0 > out.mqe) { out.mqe = Hst0; out.mqe_t = st0; }
		}
		if (gl == 0) H[en0 & hmask] = hen;
		__syncwarp(gmask);
		{ // ksw_apply_zdrop :88-104
			bool stop = false;
			if (max_H > out.max) { out.max = max_H; out.max_t = max_t; out.max_q = r - max_t; }
			else if (max_t >= out.max_t && r - max_t >= out.max_q) {
				const int tl = max_t - out.max_t, ql = (r - max_t) - out.max_q;
				const int l = tl > ql ? tl - ql : ql - tl;
				if (P.zdrop >= 0 && out.max - max_H > P.zdrop + l * P.e) { out.zdropped = 1; stop = true; }
			}
			if (stop) break;
		}
		last_st = st; last_en = en;
	}
	__syncwarp(gmask);
	// backtrack :380-385 -> ksw_backtrack :47-79 (is_rot = 1, left-aligned gaps)
	int i, j;
	if (!out.zdropped) { i = tlen - 1; j = qlen - 1; }
	else if (out.max_t >= 0 && out.max_q >= 0) { i = out.max_t; j = out.max_q; }
	else return;
	int n = 0, ovf = 0;
	{
		// The walk itself is serial (one state machine), but its loads are not: before every stretch of <= 32 diagonals the
		// group prefetches the 32 x 32 tile of p[][] the path can touch (row r0-k can only be entered at columns
		// i0-k .. i0) into shared memory (the sequence staging area is free by now), so the walker runs on shared-memory
		// latency instead of one L2/HBM round trip per step.
		uint32_t *tile = (uint32_t*)M.seq; // 32 rows of 8 words
		uint32_t *cig = M.cig; const int cig_cap = M.cig_cap;
		int state = 0;
		unsigned cur_op = 0xffu, cur_len = 0;
		while (i >= 0 && j >= 0) {
			const int i0 = i, r0 = i + j;
			for (int row = gl; row < 32; row += G) {
				const int rr = r0 - row;
				if (rr >= 0) {
					int s0, e0;
					ksw_band(rr, qlen, tlen, w, s0, e0);
					long long x0 = (long long)rr * n_col + (i0 - 31 - (s0 & ~15)); // byte offset of column i0-31 of row rr
					// a row that holds a readable entry has x0 within 31 bytes of the matrix; anything further out is never read
					// (the walk is forced there), so clamping keeps the prefetch inside this alignment's workspace
					const long long x_hi = (long long)(qlen + tlen - 1) * n_col + KSW_PMAT_PAD - 40;
					x0 = x0 < -(long long)(KSW_PMAT_PAD - 4) ? -(long long)(KSW_PMAT_PAD - 4) : (x0 > x_hi ? x_hi : x0);
					const uint32_t *src = (const uint32_t*)(pmat + (x0 & ~3LL));
					const int sh = 8 * (int)(x0 & 3);
					uint32_t wprev = src[0];
#pragma unroll
					for (int k = 0; k < 8; ++k) {
						const uint32_t wnext = src[k + 1];
						tile[row * 8 + k] = __funnelshift_r(wprev, wnext, sh);
						wprev = wnext;
					}
				}
			}
			__syncwarp(gmask);
			if (gl == 0) {
				const uint8_t *tb = (const uint8_t*)tile;
				while (i >= 0 && j >= 0 && i + j > r0 - 32) {
					const int r = i + j;
					int s0, e0;
					ksw_band(r, qlen, tlen, w, s0, e0);
					const int off = s0 & ~15, off_end = e0 | 15;
					int force_state = -1;
					if (i < off) force_state = 2;
					if (i > off_end) force_state = 1;
					const unsigned tmp = force_state < 0 ? tb[(r0 - r) * 32 + (i - (i0 - 31))] : 0u;
					if (state == 0) state = tmp & 7;
					else if (!((tmp >> (state + 2)) & 1)) state = 0;
					if (state == 0) state = tmp & 7;
					if (force_state >= 0) state = force_state;
					unsigned op;
					if (state == 0) { op = 0; --i; --j; }
					else if (state == 1 || state == 3) { op = 2; --i; }
					else { op = 1; --j; }
					if (op == cur_op) ++cur_len;
					else {
						if (cur_len) { if (n < cig_cap) cig[n++] = cur_len << 4 | cur_op; else ovf = 1; }
						cur_op = op; cur_len = 1;
					}
				}
			}
			i = __shfl_sync(gmask, i, 0, G); j = __shfl_sync(gmask, j, 0, G);
			__syncwarp(gmask);
		}
		if (gl == 0) {
			// the two trailing pushes (:73-74) merge with an equal pending op exactly as ksw_push_cigar does
			if (i >= 0) {
				if (cur_op == 2) cur_len += i + 1;
				else { if (cur_len) { if (n < cig_cap) cig[n++] = cur_len << 4 | cur_op; else ovf = 1; } cur_op = 2; cur_len = i + 1; }
			}
			if (j >= 0) {
				if (cur_op == 1) cur_len += j + 1;
				else { if (cur_len) { if (n < cig_cap) cig[n++] = cur_len << 4 | cur_op; else ovf = 1; } cur_op = 1; cur_len = j + 1; }
			}
			if (cur_len) { if (n < cig_cap) cig[n++] = cur_len << 4 | cur_op; else ovf = 1; }
		}
	}
	n = __shfl_sync(gmask, n, 0, G);
	ovf = __shfl_sync(gmask, ovf, 0, G);
	out.n_cigar = n;
	if (ovf) out.status = KSW_ST_CIGCAP;
	__syncwarp(gmask);
}

// query-offset-limited view of the CIGAR (the `cigar` iterator of src/ksw2/ksw2.nim:22-33) over the REVERSED scratch:
// returns how many leading ops (in forward order) the iterator yields.
__device__ __forceinline__ int ksw_trunc_count(const uint32_t *cig_rev, int n, int max_q)
{
	const unsigned max_off = (unsigned)max_q;
	unsigned off = 0;
	int k = 0;
	for (; k < n; ++k) {
		if (off >= max_off) break;
		const uint32_t c = cig_rev[n - 1 - k];
		if ((c & 0xf) != 2) off += c >> 4;
	}
	return k;
}
