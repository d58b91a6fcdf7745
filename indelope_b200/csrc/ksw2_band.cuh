// indelope_b200/csrc/ksw2_band.cuh -- kernel 2, banded call-site (contig -> reference window, src/indelope.nim:213-221: w = 50,
// zdrop = 400) with the BAND RING IN REGISTERS.
//
// Same results as ksw_extz2_sse (src/ksw2/csrc/ksw2_extz2_sse.c:113-388, flag 0), bit for bit, for 0 <= w with a rounded band of at
// most 96 lanes (min(qlen, tlen, w + 1) <= 80).  The column-owned variant in ksw2.cuh keeps x, v, u, y, the stale scores, the
// target codes and the exact scores of the live band in shared-memory rings and spends most of its instructions on ring
// addressing, loads and stores (116 instructions per packed word, 33 of them the recurrence).  Here the ring is 32 packed words
// = 128 columns wide and lives in registers: column word wi belongs to thread wi % G of the group, slot (wi / G) % NS with
// NS = 32 / G; a slot whose word has left the band on the left is recycled for the word 128 columns further right (zeroed: the
// reference's calloc'ed, never written lanes, :173) when the band's right edge gets near.  The rounded band [st, en] (:205) plus
// the word left of it (x[st-1], v[st-1], :207-210) plus the 16 lanes the score blocks can overrun (:215-228) plus the 16
// lanes being entered never span more than 116 columns, so a live word is never recycled.  A cell (t, r - t) reads x, v of
// column t - 1 on the previous diagonal: the word to the left, i.e. the previous thread of the group -- one SHFL per slot carries
// the neighbour's top x and v lanes and its top exact score (the H[en0 - 1] of :318) in one register, taken for all slots before
// any word is updated, so no ordering between the words of a diagonal is needed.  u, y and the exact scores stay with their
// column.  The 16-lane rounding of the band, the stale scores and stale lanes outside the exact band (they feed real cells
// later) are reproduced exactly: every word of [st, en] runs the core on whatever its lanes hold, in
// the carry-free form when every input byte is in [0, 63] and in the wrapping int8 form otherwise (ksw_core_word).
//
// Exact scores (:312-349) are uint16 offsets g[t] = H[t] + (q+e)(r+1) + bias as in ksw2.cuh, two registers per word; the band
// maximum is a packed max per thread and one butterfly per diagonal; its position (SSE tie order) is looked up only when it can
// raise the overall maximum or trigger z-drop.  The backtrack matrix has the reference's layout p[r][t - st] (pitch rounded to
// 32 bytes) and is walked by ksw_backtrack's state machine (:47-79) through the shared 32 x 32 tile prefetch.
#pragma once
#include "ksw2.cuh"

// widest rounded band this variant serves: 128 ring columns - 4 (left neighbour) - 16 (score overrun) - 16 (entering block)
#define KSW_BAND_MAX_NCOL 96
__host__ __device__ inline bool ksw_band_serves(int qlen, int tlen, int w) { return w >= 0 && ksw_ncol(qlen > 0 ? qlen : 1, tlen > 0 ? tlen : 1, w) <= KSW_BAND_MAX_NCOL; }
// shared memory of a group: the reversed, padded query; later the 1 KB backtrack tile in the same bytes
__host__ __device__ inline size_t ksw_band_smem(int seq_cap) { size_t b = (size_t)((seq_cap + 15) & ~15); if (b < KSW_BTILE_BYTES) b = KSW_BTILE_BYTES; return (b + 127) & ~(size_t)127; }

// ksw_backtrack (:47-79, is_rot = 1) over p[r][t - off[r]] with the 32 x 32 tile prefetch; ALL 32 threads of the warp call it, G
// threads per alignment (i = j = -1 for a group without a path).  The walker is thread 0 of the group; n / ovf come back in every
// thread of the group.  `tile` = 1 KB of the group's shared memory.
template <int G>
__device__ __forceinline__ void ksw_walk_cols(int i, int j, int qlen, int tlen, int w, int pitch, const uint8_t *pmat, uint32_t *tile, uint32_t *cig, int cig_cap, int &n_out, int &ovf_out)
{
	const int lane = lane_id(), gl = lane & (G - 1);
	int n = 0, ovf = 0, state = 0;
	unsigned cur_op = 0xffu, cur_len = 0;
	const long long x_hi = (long long)(qlen + tlen - 1) * pitch + KSW_PMAT_PAD - 40;
	while (__any_sync(FULL_MASK, i >= 0 && j >= 0)) {
		const bool more = i >= 0 && j >= 0;
		const int i0 = i, r0 = i + j;
#pragma unroll 1
		for (int pass = 0; pass < 32 / (4 * G) + (32 % (4 * G) ? 1 : 0); ++pass) {
			uint32_t wv[4][9]; int shv[4];
#pragma unroll
			for (int q4 = 0; q4 < 4; ++q4) {
				const int row = gl + G * (q4 + 4 * pass), rr = r0 - row;
				int s0 = 0, e0 = 0;
				if (more && rr >= 0) ksw_band(rr, qlen, tlen, w, s0, e0);
				long long x0 = (long long)rr * pitch + (i0 - 31 - (s0 & ~15)); // byte offset of column i0-31 of row rr
				x0 = x0 < -(long long)(KSW_PMAT_PAD - 4) ? -(long long)(KSW_PMAT_PAD - 4) : (x0 > x_hi ? x_hi : x0);
				const uint32_t *src = (const uint32_t*)(pmat + (x0 & ~3LL));
				shv[q4] = 8 * (int)(x0 & 3);
				const int k0 = (31 - row) >> 2; // row r0 - row can only be entered at columns i0 - row .. i0
				const bool ld = more && rr >= 0 && row < 32;
#pragma unroll
				for (int k = 0; k < 9; ++k) wv[q4][k] = (ld && k >= k0) ? src[k] : 0u;
			}
#pragma unroll
			for (int q4 = 0; q4 < 4; ++q4) {
				const int row = gl + G * (q4 + 4 * pass), k0 = (31 - row) >> 2;
				if (row < 32) {
#pragma unroll
					for (int k = 0; k < 8; ++k) if (k >= k0) tile[row * 8 + k] = __funnelshift_r(wv[q4][k], wv[q4][k + 1], shv[q4]);
				}
			}
		}
		__syncwarp();
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
		i = __shfl_sync(FULL_MASK, i, lane & ~(G - 1)); j = __shfl_sync(FULL_MASK, j, lane & ~(G - 1));
		__syncwarp();
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
	n_out = __shfl_sync(FULL_MASK, n, lane & ~(G - 1));
	ovf_out = __shfl_sync(FULL_MASK, ovf, lane & ~(G - 1));
}

template <int G> __device__ __forceinline__ unsigned ksw_gmax(unsigned v)
{
#pragma unroll
	for (int d = 1; d < G; d <<= 1) { const unsigned o = __shfl_xor_sync(FULL_MASK, v, d); v = v > o ? v : o; }
	return v;
}
template <int G> __device__ __forceinline__ unsigned ksw_gmin(unsigned v)
{
#pragma unroll
	for (int d = 1; d < G; d <<= 1) { const unsigned o = __shfl_xor_sync(FULL_MASK, v, d); v = v < o ? v : o; }
	return v;
}
template <int G> __device__ __forceinline__ bool ksw_gany(bool p, int lane) { return ((__ballot_sync(FULL_MASK, p) >> (lane & ~(G - 1))) & ((1u << G) - 1u)) != 0u; }

// per-diagonal values shared by the slots of a thread
struct KswBandDiag {
	int st, st0, en0, en, bend;     // rounded band, exact band, one past the last lane of the score blocks
	int cq;                          // lane t meets the reversed-query byte cq + t
	unsigned uon, ubd, onm;          // slots with a word of [st, max(en, bend-1)] in this thread (onm) / in any thread of the warp (uon); ubd: some thread's word there is not interior
	bool first_const;                // the word at st takes the constant boundary (x1 = 0, v1 = v1c) instead of its left neighbour
	uint32_t v1c;                    // v1 << 8 (byte 1 of the neighbour word)
	uint8_t *prow;                   // backtrack row of this diagonal, column st
	const uint32_t *qrp;             // reversed query words (shared memory), KSW_QR_PAD bytes of zeros in front
	bool wild, fast_ok;
	int r, last_st0, last_en0, qlen, tlen, w, qe;
};

// One slot of one diagonal.  INTERIOR: every thread of the warp whose word in this slot is live has all four lanes inside the exact
// band [st0, en0) and right of st: fresh scores, no masks.  Otherwise the general form: lanes [st0, bend) get fresh scores and the
// others keep their stale ones (:215-228), the word at st may take the constant boundary, words right of en only update scores,
// the exact scores of lanes [st0, en0) are updated (:323-348) and the word that holds en0 builds H[en0] (:318).
template <int G, int NS, int S, bool INTERIOR>
__device__ __forceinline__ void ksw_band_slot(const KswParams &P, const KswBandDiag &dg, uint32_t recv, const int (&tcol)[NS], uint32_t (&x)[NS], uint32_t (&v)[NS], uint32_t (&u)[NS],
                                              uint32_t (&y)[NS], uint32_t (&Sc)[NS], const uint32_t (&T)[NS], uint32_t (&gx)[NS], uint32_t (&gy)[NS], uint32_t &bh2, unsigned &ghen)
{
	const int t = tcol[S];
	const int cqt = dg.cq + t; // >= -KSW_QR_PAD for every lane the score blocks touch
	const uint32_t *qw = dg.qrp + (cqt >> 2);
	const uint32_t sq2 = __funnelshift_r(qw[0], qw[1], 8 * (cqt & 3));
	const uint32_t sq = T[S];
	uint32_t sc = sel4(msb_to_mask4((sq ^ sq2) + 0x7f7f7f7fu), P.misq_4, P.maxsc_4);
	if (dg.wild) sc = ksw_wild_score(sq, sq2, sc, P.qe2_4);
	if (INTERIOR) {
		Sc[S] = sc;
		const uint32_t xt1 = __byte_perm(recv, x[S], 0x6540), vt1 = __byte_perm(recv, v[S], 0x6541);
		uint32_t d, un, vn, xn, yn;
		ksw_core_word(P, dg.fast_ok, sc, xt1, vt1, u[S], y[S], xn, vn, un, yn, d);
		x[S] = xn; v[S] = vn; u[S] = un; y[S] = yn;
		KSW_PSTORE((uint32_t*)(dg.prow + (t - dg.st)), d);
		gx[S] += __byte_perm(vn, 0u, 0x4140); gy[S] += __byte_perm(vn, 0u, 0x4342);
		bh2 = __vimax3_u16x2(bh2, gx[S], gy[S]);
	} else {
		{ // scores: lanes [lo, hi) of this word belong to the 16-wide blocks
			const int lo = dg.st0 - t, hi = dg.bend - t;
			if (hi > 0 && lo < 4) {
				uint32_t m = 0xffffffffu;
				if (lo > 0) m <<= 8 * lo;
				if (hi < 4) m &= 0xffffffffu >> (8 * (4 - hi));
				Sc[S] = sel4(m, sc, Sc[S]);
			}
		}
		if (t <= dg.en) { // core over the rounded band
			const uint32_t nb = (t == dg.st && dg.first_const) ? dg.v1c : recv; // x, v of the lane left of this word; recv keeps the neighbour's exact score
			const uint32_t xt1 = __byte_perm(nb, x[S], 0x6540), vt1 = __byte_perm(nb, v[S], 0x6541);
			uint32_t d, un, vn, xn, yn;
			ksw_core_word(P, dg.fast_ok, Sc[S], xt1, vt1, u[S], y[S], xn, vn, un, yn, d);
			x[S] = xn; v[S] = vn; u[S] = un; y[S] = yn;
			KSW_PSTORE((uint32_t*)(dg.prow + (t - dg.st)), d);
			const int lo = dg.st0 - t, hi = dg.en0 - t; // exact scores of the in-band columns st0 .. en0-1 of this word: g[t] += v8[t]
			const uint32_t gox = gx[S], goy = gy[S];     // the old scores: H[en0] is built from the OLD H[en0-1] (:318)
			if (hi > 0 && lo < 4) {
				if (lo <= 0 && hi >= 4) {
					gx[S] += __byte_perm(vn, 0u, 0x4140); gy[S] += __byte_perm(vn, 0u, 0x4342);
					bh2 = __vimax3_u16x2(bh2, gx[S], gy[S]);
				} else {
					uint32_t m = 0xffffffffu;
					if (lo > 0) m <<= 8 * lo;
					if (hi < 4) m &= 0xffffffffu >> (8 * (4 - hi));
					const uint32_t vm = vn & m;
					gx[S] += __byte_perm(vm, 0u, 0x4140); gy[S] += __byte_perm(vm, 0u, 0x4342);
					bh2 = __vimax3_u16x2(bh2, gx[S] & __byte_perm(m, 0u, 0x1100), gy[S] & __byte_perm(m, 0u, 0x3322));
				}
			}
			if (hi >= 0 && hi < 4) { // this word holds en0: H[en0] (:318 / :349), kept as g
				unsigned gh;
				if (dg.r == 0) gh = (unsigned)((int)(vn & 0xffu) - dg.qe + 2 * dg.qe); // H[0] = v[0] - 2(q+e), g = H + (q+e) + bias
				else {
					unsigned gprev; // g of column c = max(en0 - 1, 0) before this diagonal's update
					if (dg.en0 == 0) gprev = gox & 0xffffu;
					else if (hi == 0) gprev = recv >> 16;
					else gprev = ((hi - 1) & 2 ? goy : gox) >> (16 * ((hi - 1) & 1)) & 0xffffu;
					const int c = dg.en0 > 0 ? dg.en0 - 1 : 0;
					if (c < dg.last_st0 || c > dg.last_en0) { // the column was not part of the previous band: its offset is older (the band left the matrix on the right)
						int rr = dg.r - 1;
						for (; rr > 0; --rr) { int s_, e_; ksw_band(rr, dg.qlen, dg.tlen, dg.w, s_, e_); if (s_ <= c && c <= e_) break; }
						gprev += (unsigned)(dg.qe * (dg.r - 1 - rr));
					}
					const uint32_t add = dg.en0 > 0 ? un : vn; // + u8[en0] or + v8[en0]
					gh = gprev + ((add >> (8 * hi)) & 0xffu);
				}
				ghen = gh;
				const uint32_t ins = (gh & 0xffffu) << (16 * (hi & 1)), keep = 0xffffu << (16 * ((hi & 1) ^ 1));
				if (hi & 2) gy[S] = (gy[S] & keep) | ins; else gx[S] = (gx[S] & keep) | ins;
			}
		}
	}
}

// the slots of one diagonal, unrolled by template recursion (constant register indices from the start)
template <int G, int NS, int S>
struct KswBandSlots {
	static __device__ __forceinline__ void run(const KswParams &P, const KswBandDiag &dg, const uint32_t (&recv)[NS], const int (&tcol)[NS], uint32_t (&x)[NS], uint32_t (&v)[NS],
	                                           uint32_t (&u)[NS], uint32_t (&y)[NS], uint32_t (&Sc)[NS], const uint32_t (&T)[NS], uint32_t (&gx)[NS], uint32_t (&gy)[NS], uint32_t &bh2, unsigned &ghen)
	{
		if (dg.uon & (1u << S)) {
			const bool on = (dg.onm & (1u << S)) != 0u;
			if (dg.ubd & (1u << S)) { if (on) ksw_band_slot<G, NS, S, false>(P, dg, recv[S], tcol, x, v, u, y, Sc, T, gx, gy, bh2, ghen); }
			else if (on) ksw_band_slot<G, NS, S, true>(P, dg, recv[S], tcol, x, v, u, y, Sc, T, gx, gy, bh2, ghen);
		}
		KswBandSlots<G, NS, S + 1>::run(P, dg, recv, tcol, x, v, u, y, Sc, T, gx, gy, bh2, ghen);
	}
};
template <int G, int NS>
struct KswBandSlots<G, NS, NS> {
	static __device__ __forceinline__ void run(const KswParams &, const KswBandDiag &, const uint32_t (&)[NS], const int (&)[NS], uint32_t (&)[NS], uint32_t (&)[NS], uint32_t (&)[NS],
	                                           uint32_t (&)[NS], uint32_t (&)[NS], const uint32_t (&)[NS], uint32_t (&)[NS], uint32_t (&)[NS], uint32_t &, unsigned &) {}
};

// ALL 32 threads of a warp call this together: 32 / G groups of G threads, one alignment per group (valid = 0 for a group without one),
// the anti-diagonal loop in lockstep over the groups.  The caller has checked ksw_band_serves(qlen, tlen, P.w) for every valid group.
// M.seq (seq_cap bytes) stages the reversed query; M.xvuy points at 1 KB of shared memory for the backtrack tile (it may alias M.seq).
template <int G, bool EZ_FULL>
__device__ void ksw2_band(bool valid, int qlen, const KswQuery query, int tlen, const uint8_t *target, const KswParams P, const KswMem M, KswOut &out)
{
	constexpr int NS = 32 / G;
	static_assert(G == 4 || G == 8, "4 or 8 threads per alignment");
	const int lane = lane_id(), gl = lane & (G - 1);
	ksw_reset(out);
	const int qe = P.q + P.e, gbias = 2 * qe;
	int min_sc = P.mismatch < 0 ? P.mismatch : 0;
	if (P.match < min_sc) min_sc = P.match;
	const int w = P.w;
	const int n_col = valid ? ksw_ncol(qlen > 0 ? qlen : 1, tlen > 0 ? tlen : 1, w) : 16;
	const int pitch = ksw_pitch(n_col);
	bool live = valid;
	if (live) {
		if (qlen <= 0 || tlen <= 0) { out.status = KSW_ST_EARLY; live = false; }     // :147
		else if (-min_sc > 2 * qe) { out.status = KSW_ST_EARLY; live = false; }      // :171
		else if (w < 0 || n_col > KSW_BAND_MAX_NCOL) { out.status = KSW_ST_RCAP; live = false; } // the caller picked the wrong variant
		else if (ksw_seq_bytes(qlen, tlen) > (size_t)M.seq_cap) { out.status = KSW_ST_SEQCAP; live = false; }
		else if ((size_t)(qlen + tlen - 1) * (size_t)pitch + 2 * KSW_PMAT_PAD > M.p_cap) { out.status = KSW_ST_PCAP; live = false; }
		else if ((long long)(qlen + tlen + 2) * qe + (long long)(qlen < tlen ? qlen : tlen) * (P.match > 0 ? P.match : 0) + gbias >= 0xF000) { out.status = KSW_ST_HCAP; live = false; }
	}
	const bool run = live;
	uint8_t *pmat = M.pmat + KSW_PMAT_PAD;
	uint8_t *qrp = M.seq;
	const int nr = live ? qlen + tlen - 1 : 0;

	// the ring: slot s of this thread starts as column word s * G + gl; every lane reads as calloc'ed memory (:173), the stale scores as
	// s = 0, the target codes zero padded (:187)
	uint32_t x[NS], v[NS], u[NS], y[NS], Sc[NS], T[NS], gx[NS], gy[NS]; int tcol[NS];
	bool wild = false;
#pragma unroll
	for (int s = 0; s < NS; ++s) {
		tcol[s] = 4 * (s * G + gl);
		x[s] = v[s] = u[s] = y[s] = 0u; gx[s] = gy[s] = 0u; Sc[s] = P.qe2_4;
		T[s] = live ? ksw_target_word(target, tlen, tcol[s]) : 0u;
		wild |= ksw_has4(T[s]);
	}
	if (live) { // qr: KSW_QR_PAD zero bytes, the reversed query, zero padded (:188)
		const int nq = KSW_QR_PAD + ((qlen + 35) & ~3);
		for (int i = gl; i < nq; i += G) {
			const int k = i - KSW_QR_PAD;
			const uint8_t c = (k >= 0 && k < qlen) ? ksw_query_code(query, qlen - 1 - k) : (uint8_t)0; wild |= c == 4; qrp[i] = c;
		}
	}
	wild = ksw_gany<G>(wild, lane);
	__syncwarp();

	KswBandDiag dg;
	dg.qrp = (const uint32_t*)qrp + (KSW_QR_PAD >> 2); dg.qlen = qlen; dg.tlen = tlen; dg.w = w; dg.qe = qe;
	dg.fast_ok = P.match + 2 * qe <= 63 && P.q >= 0 && P.q + 2 * P.e + min_sc >= 0;
	const int srcl = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1)); // the thread that owns the word left of mine
	int last_st = -1, last_en = -1, last_st0 = 0, last_en0 = -1, en_clr = 15, base = 0; // ring = columns [base, base + 128)
	unsigned g_high = 0;
	int st0 = 0, en0 = 0;
	for (int r = 0; ; ++r) {
		bool act = live && r < nr;
		if (act && st0 > en0) { out.zdropped = 1; live = false; act = false; } // :200-203
		if (!__any_sync(FULL_MASK, act)) break;
		const int st = st0 & ~15, en = en0 | 15;    // :205
		if (act) out.cells += en0 - st0 + 1;
		const bool keep_prev = st > 0 && st - 1 >= last_st && st - 1 <= last_en; // :207-211
		const int bend = st0 + (int)((((unsigned)(en0 - st0) >> 4) + 1u) << 4);
		const int wl_last = (en > bend - 1 ? en : bend - 1) | 3;
		const int en1 = st0 + (((en0 - st0) >> 2) << 2);
		const unsigned bandw = (unsigned)(en0 - st0);
		const int goff = qe * (r + 1) + gbias;
		// the neighbour exchange, for every slot, before any word of this diagonal is updated: top x lane, top v lane, top exact score
		uint32_t recv[NS];
		{
			uint32_t C[NS];
#pragma unroll
			for (int s = 0; s < NS; ++s) C[s] = __byte_perm(__byte_perm(x[s], v[s], 0x0073), gy[s], 0x7610);
#pragma unroll
			for (int s = 0; s < NS; ++s) {
				uint32_t send = C[s];
				if (gl == G - 1) send = C[(s + NS - 1) % NS]; // the word left of thread 0's word in slot s sits in the previous slot of thread G-1
				recv[s] = __shfl_sync(FULL_MASK, send, srcl);
			}
		}
		// which of my slots hold a word of this diagonal, and which of those are interior
		{
			const int ilo = st0 > st ? st0 : st + 4, ihi = en0 - 4;
			unsigned onm = 0, inm = 0;
#pragma unroll
			for (int s = 0; s < NS; ++s) {
				const int t = tcol[s];
				const bool on = act && t >= st && t <= wl_last;
				onm |= (unsigned)on << s; inm |= (unsigned)(on && t >= ilo && t <= ihi) << s;
			}
			dg.onm = onm; dg.uon = __reduce_or_sync(FULL_MASK, onm); dg.ubd = __reduce_or_sync(FULL_MASK, onm & ~inm);
		}
		dg.st = st; dg.st0 = st0; dg.en0 = en0; dg.en = en; dg.bend = bend; dg.cq = qlen - 1 - r; dg.r = r; dg.last_st0 = last_st0; dg.last_en0 = last_en0;
		dg.first_const = !keep_prev; dg.v1c = (st == 0 && r) ? ((uint32_t)(P.q & 0xff) << 8) : 0u;
		dg.prow = pmat + (size_t)r * pitch; dg.wild = wild;
		uint32_t bh2 = 0; unsigned ghen = 0;
		KswBandSlots<G, NS, 0>::run(P, dg, recv, tcol, x, v, u, y, Sc, T, gx, gy, bh2, ghen);
		// band max of the updated columns, then the en0 cell (:318 / :349): its owner computed it, everyone needs it
		unsigned mg = (bh2 & 0xffffu) > (bh2 >> 16) ? (bh2 & 0xffffu) : (bh2 >> 16);
		mg = ksw_gmax<G>(mg);
		ghen = __shfl_sync(FULL_MASK, ghen, (lane & ~(G - 1)) | ((en0 >> 2) & (G - 1)));
		if (!act) ghen = 0;
		const int hen = (int)ghen - goff;
		g_high |= (unsigned)(mg >= 0xF000u) | (unsigned)(ghen >= 0xF000u);
		const int mh = (int)mg - goff;
		int max_H = hen, max_t = en0; // the initial candidate (H[en0], en0) wins every tie (:318-321)
		const bool better = act && mh > hen;
		const bool need_t = better && (mh > out.max || (P.zdrop >= 0 && out.max - mh > P.zdrop));
		if (__any_sync(FULL_MASK, need_t)) { // position of the band max in the SSE tie order (:316-348), only when it matters
			unsigned best = 0xffffffffu;
			if (need_t) {
#pragma unroll
				for (int s = 0; s < NS; ++s) {
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const unsigned val = ((c < 2 ? gx[s] : gy[s]) >> (16 * (c & 1))) & 0xffffu;
						const int t = tcol[s] + c;
						if ((unsigned)(t - st0) < bandw && val == mg) { const unsigned rk = ksw_tie_rank(t, st0, en1); best = rk < best ? rk : best; }
					}
				}
			}
			const unsigned rk = ksw_gmin<G>(best);
			if (need_t) { max_H = mh; max_t = st0 + (int)((rk - 1) & 0xfffffu); }
		}
		if (EZ_FULL) {
			if (act && en0 == tlen - 1) {
				if (hen > out.mte) { out.mte = hen; out.mte_q = r - en; }
				if (r == qlen + tlen - 2) out.score = hen; // H[tlen-1]
			}
			const bool at_qend = act && r - st0 == qlen - 1;
			if (__any_sync(FULL_MASK, at_qend)) { // H[st0] of the last query row (:353-354): its owner has it
				unsigned gv = 0;
#pragma unroll
				for (int s = 0; s < NS; ++s) if (tcol[s] == (st0 & ~3)) gv = (((st0 & 2) ? gy[s] : gx[s]) >> (16 * (st0 & 1))) & 0xffffu;
				gv = __shfl_sync(FULL_MASK, gv, (lane & ~(G - 1)) | ((st0 >> 2) & (G - 1)));
				if (at_qend) {
					const int Hst0 = st0 == en0 ? hen : (int)gv - goff;
					if (Hst0 > out.mqe) { out.mqe = Hst0; out.mqe_t = st0; }
				}
			}
		}
		// band of the next diagonal (:196-199): the block it enters, if any, and its :212 patch are applied now
		int st0n, en0n;
		ksw_band(r + 1, qlen, tlen, w, st0n, en0n);
		const bool nxt = act && r + 1 < nr && st0n <= en0n;
		const bool enter = nxt && (en0n | 15) > en_clr;
		if (__any_sync(FULL_MASK, enter)) {
			bool w4 = false;
			if (enter) {
				en_clr += 16;
				if (base + 128 < en_clr + 17) { // the ring must reach 16 lanes past the block that enters (the score overrun): recycle the 16 leftmost columns
					base += 16;
#pragma unroll
					for (int s = 0; s < NS; ++s)
						if (tcol[s] < base) {
							tcol[s] += 128;
							x[s] = v[s] = u[s] = y[s] = 0u; Sc[s] = P.qe2_4;
							T[s] = ksw_target_word(target, tlen, tcol[s]); w4 |= ksw_has4(T[s]);
						}
				}
			}
			wild |= ksw_gany<G>(w4, lane);
		}
		const bool patch = nxt && (en0n | 15) >= r + 1;
		if (__any_sync(FULL_MASK, patch)) { // :212 of the next diagonal: y[r+1] = 0, u[r+1] = q
			if (patch) {
				const int tc = (r + 1) & ~3, sh = 8 * ((r + 1) & 3);
#pragma unroll
				for (int s = 0; s < NS; ++s)
					if (tcol[s] == tc) { u[s] = (u[s] & ~(0xffu << sh)) | ((uint32_t)(P.q & 0xff) << sh); y[s] &= ~(0xffu << sh); }
			}
		}
		if (act && (better ? need_t : true)) { // ksw_apply_zdrop :88-104 (skipped when the band max can neither raise the max nor trigger z-drop)
			if (max_H > out.max) { out.max = max_H; out.max_t = max_t; out.max_q = r - max_t; }
			else if (max_t >= out.max_t && r - max_t >= out.max_q) {
				const int tl = max_t - out.max_t, ql = (r - max_t) - out.max_q;
				const int l = tl > ql ? tl - ql : ql - tl;
				if (P.zdrop >= 0 && out.max - max_H > P.zdrop + l * P.e) { out.zdropped = 1; live = false; }
			}
		}
		last_st = st; last_en = en; last_st0 = st0; last_en0 = en0;
		st0 = st0n; en0 = en0n;
	}
	if (g_high) out.status = KSW_ST_HCAP;
	__syncwarp();
	// backtrack :380-385
	int i = -1, j = -1;
	if (run) {
		if (!out.zdropped) { i = tlen - 1; j = qlen - 1; }
		else if (out.max_t >= 0 && out.max_q >= 0) { i = out.max_t; j = out.max_q; }
	}
	int n = 0, ovf = 0;
	ksw_walk_cols<G>(i, j, qlen, tlen, w, pitch, pmat, (uint32_t*)M.xvuy, M.cig, M.cig_cap, n, ovf);
	out.n_cigar = n;
	if (ovf) out.status = KSW_ST_CIGCAP;
	__syncwarp();
}
