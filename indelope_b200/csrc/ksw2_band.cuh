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
//
// MEASURED (round 2, profiles/r02_align_variants.txt): bit-exact, 10.4 G warp instructions against the shared-memory variant's 11.2 G on
// the chr1 workload, but 13.8 ms against 12.9 ms: both are bound by the ALU pipe (one warp instruction per two cycles per scheduler:
// LOP3 / PRMT / ISETP / SEL / SHF all issue there), and this variant turns the other one's shared-memory loads and stores (LSU pipe) into
// selects, masks and slot dispatch on the ALU pipe (65 % of its instructions against ~50 %), at 4 instead of 6 warps per scheduler.
// A block-owned layout of the same ring (thread = one 16-lane block, one SHFL per diagonal, 128-bit backtrack stores, straight-line code)
// was also written and verified bit-exact: 10.9 G instructions, ALU pipe 84 % busy, 14.2 ms.  The shared-memory variant therefore stays
// the default of the banded call-site; this one runs with IDL_BAND_REGS=1 and in the parity tests.
#pragma once
#include "ksw2.cuh"

// widest rounded band this variant serves: 128 ring columns - 4 (left neighbour) - 16 (score overrun) - 16 (entering block)
#define KSW_BAND_MAX_NCOL 96
__host__ __device__ inline bool ksw_band_serves(int qlen, int tlen, int w) { return w >= 0 && ksw_ncol(qlen > 0 ? qlen : 1, tlen > 0 ? tlen : 1, w) <= KSW_BAND_MAX_NCOL; }
// shared memory of a group: the reversed, padded query; later the 1 KB backtrack tile in the same bytes
// + 32 bytes for the end-of-query / end-of-target scores (KswBandEz)
__host__ __device__ inline size_t ksw_band_smem(int seq_cap) { size_t b = (size_t)((seq_cap + 15) & ~15); if (b < KSW_BTILE_BYTES) b = KSW_BTILE_BYTES; return (b + 32 + 127) & ~(size_t)127; }

// ksw_backtrack (:47-79, is_rot = 1) over p[r][t - off[r]] with the 32 x 32 tile prefetch; ALL 32 threads of the warp call it, G
// threads per alignment (i = j = -1 for a group without a path).  The walker is thread 0 of the group; n / ovf come back in every
// thread of the group.  `tile` = 1 KB of the group's shared memory.
template <int G>
__device__ __forceinline__ void ksw_walk_cols(int i, int j, int qlen, int tlen, int w, int pitch, const uint8_t *pmat, uint32_t *tile, uint32_t *cig, int cig_cap, int &n_out, int &ovf_out)
{
	const int lane = lane_id(), gl = lane & (G - 1);
	int n = 0, ovf = 0, state = 0;
	unsigned cur_op = 0xffu, cur_len = 0;
	const int p_hi = ksw_rows_bound(qlen, tlen, w) * pitch + KSW_PMAT_PAD;
	while (__any_sync(FULL_MASK, i >= 0 && j >= 0)) {
		const bool more = i >= 0 && j >= 0;
		const int i0 = i, r0 = i + j;
		ksw_tile_fetch(more, [&](int k) -> int {
			const int rr = r0 - k;
			if (rr < 0) return INT_MIN;
			int s0, e0;
			ksw_band(rr, qlen, tlen, w, s0, e0);
			return rr * pitch + (i0 - (s0 & ~15)); // byte offset of column i0 in row rr
		}, pmat, -KSW_PMAT_PAD, p_hi, tile);
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
				const unsigned tmp = force_state < 0 ? tb[(r0 - r) * KSW_BTILE_ROW + 32 + (i0 & 3) - (i0 - i)] : 0u;
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
	int st, en, en0;                 // rounded band (:205) and the last exact column
	int cq;                          // lane t meets the reversed-query byte cq + t
	unsigned uon, onm;               // slots with a word of [st, max(en, bend-1)] in this thread (onm) / in any thread of the warp (uon)
	uint32_t A4, B4, E4;             // st0 - st, bend - st, en0 - st, replicated into the four bytes (all <= 111)
	uint32_t nb_first;               // what the word at st reads instead of its left neighbour when the boundary is constant (x1 = 0, v1 << 8), else ~0
	uint8_t *prow;                   // backtrack row of this diagonal, column st
	const uint32_t *qrp;             // reversed query words (shared memory), KSW_QR_PAD bytes of zeros in front
	bool wild, fast_ok;
	int r, st0, last_st0, last_en0, qlen, tlen, w, qe;
};

// One slot of one diagonal, every case in one branch-free form (byte masks from lane numbers relative to st, all below 128):
// lanes [st0, bend) get fresh scores and the others keep their stale ones (:215-228); words up to en run the core (:262-284), the word
// at st with the constant boundary when the left neighbour is not valid (:207-211); words right of en only had their scores
// updated; the exact scores of lanes [st0, en0) are updated (:323-348); the word that holds en0 builds H[en0] (:318).
template <int G, int NS, int S>
__device__ __forceinline__ void ksw_band_slot(const KswParams &P, const KswBandDiag &dg, const uint32_t recv, const int (&tcol)[NS], uint32_t (&x)[NS], uint32_t (&v)[NS], uint32_t (&u)[NS],
                                              uint32_t (&y)[NS], uint32_t (&Sc)[NS], const uint32_t (&T)[NS], uint32_t (&gx)[NS], uint32_t (&gy)[NS], uint32_t &bh2)
{
	const int t = tcol[S], d0 = t - dg.st;                       // d0 in [0, 124]
	const int cqt = dg.cq + t;                                   // >= -KSW_QR_PAD for every lane the score blocks touch
	const uint32_t *qw = dg.qrp + (cqt >> 2);
	const uint32_t sq2 = __funnelshift_r(qw[0], qw[1], 8 * (cqt & 3));
	const uint32_t sq = T[S];
	uint32_t sc = sel4(msb_to_mask4((sq ^ sq2) + 0x7f7f7f7fu), P.misq_4, P.maxsc_4);
	if (dg.wild) sc = ksw_wild_score(sq, sq2, sc, P.qe2_4);
	const uint32_t O80 = (uint32_t)d0 * 0x01010101u + 0x83828180u; // 0x80 + lane number relative to st, per byte
	const uint32_t mA = msb_to_mask4(O80 - dg.A4);               // lanes >= st0
	const uint32_t mB = msb_to_mask4(O80 - dg.B4);               // lanes >= bend
	const uint32_t z0 = sel4(mA & ~mB, sc, Sc[S]);
	Sc[S] = z0;
	if (t <= dg.en) { // core over the rounded band
		const uint32_t nb = (d0 == 0 && dg.nb_first != 0xffffffffu) ? dg.nb_first : recv; // x, v of the lane left of this word
		const uint32_t xt1 = __byte_perm(nb, x[S], 0x6540), vt1 = __byte_perm(nb, v[S], 0x6541);
		uint32_t d, un, vn, xn, yn;
		ksw_core_word(P, dg.fast_ok, z0, xt1, vt1, u[S], y[S], xn, vn, un, yn, d);
		x[S] = xn; v[S] = vn; u[S] = un; y[S] = yn;
		KSW_PSTORE((uint32_t*)(dg.prow + d0), d);
		const uint32_t mG = mA & ~msb_to_mask4(O80 - dg.E4);       // lanes in [st0, en0): g[t] += v8[t] (:323-348)
		const uint32_t vm = vn & mG;
		gx[S] += __byte_perm(vm, 0u, 0x4140); gy[S] += __byte_perm(vm, 0u, 0x4342);
		bh2 = __vimax3_u16x2(bh2, gx[S] & __byte_perm(mG, 0u, 0x1100), gy[S] & __byte_perm(mG, 0u, 0x3322));
	}
}

// H[en0] (:318 / :349), kept as g: H[en0] = H_old[en0 - 1] + u_new[en0] - (q+e) (v_new[0] on column 0).  Runs once per diagonal, after the
// slots, in the thread that owns column en0; `gprev` = g of column max(en0 - 1, 0) before this diagonal's update, re-based by the caller.
template <int NS, int S>
__device__ __forceinline__ unsigned ksw_band_en0(int en0, unsigned gprev, const uint32_t (&u)[NS], const uint32_t (&v)[NS], uint32_t (&gx)[NS], uint32_t (&gy)[NS])
{
	const int hi = en0 & 3;
	const uint32_t add = en0 > 0 ? u[S] : v[S]; // + u8[en0] or + v8[en0]
	const unsigned gh = gprev + ((add >> (8 * hi)) & 0xffu);
	const uint32_t ins = (gh & 0xffffu) << (16 * (hi & 1)), keep = 0xffffu << (16 * ((hi & 1) ^ 1));
	if (hi & 2) gy[S] = (gy[S] & keep) | ins; else gx[S] = (gx[S] & keep) | ins;
	return gh;
}
template <int NS, int S>
__device__ __forceinline__ unsigned ksw_band_glane(int c, const uint32_t (&gx)[NS], const uint32_t (&gy)[NS]) { return (((c & 2) ? gy[S] : gx[S]) >> (16 * (c & 1))) & 0xffffu; }

// the slots of one diagonal, unrolled by template recursion (constant register indices from the start)
template <int G, int NS, int S>
struct KswBandSlots {
	static __device__ __forceinline__ void run(const KswParams &P, const KswBandDiag &dg, const int srcl, const int gl, const int (&tcol)[NS], uint32_t (&x)[NS], uint32_t (&v)[NS],
	                                           uint32_t (&u)[NS], uint32_t (&y)[NS], uint32_t (&Sc)[NS], const uint32_t (&T)[NS], uint32_t (&gx)[NS], uint32_t (&gy)[NS],
	                                           const uint32_t (&C)[NS], uint32_t &bh2)
	{
		if (dg.uon & (1u << S)) {
			// the neighbour exchange: the word left of thread 0's word in slot s sits in the previous slot of thread G-1.  C holds the values of
			// the previous diagonal for every slot, so the order in which the slots are updated does not matter.
			const uint32_t recv = __shfl_sync(FULL_MASK, gl == G - 1 ? C[(S + NS - 1) % NS] : C[S], srcl);
			if (dg.onm & (1u << S)) ksw_band_slot<G, NS, S>(P, dg, recv, tcol, x, v, u, y, Sc, T, gx, gy, bh2);
		}
		KswBandSlots<G, NS, S + 1>::run(P, dg, srcl, gl, tcol, x, v, u, y, Sc, T, gx, gy, C, bh2);
	}
};
template <int G, int NS>
struct KswBandSlots<G, NS, NS> {
	static __device__ __forceinline__ void run(const KswParams &, const KswBandDiag &, const int, const int, const int (&)[NS], uint32_t (&)[NS], uint32_t (&)[NS], uint32_t (&)[NS],
	                                           uint32_t (&)[NS], uint32_t (&)[NS], const uint32_t (&)[NS], uint32_t (&)[NS], uint32_t (&)[NS], const uint32_t (&)[NS], uint32_t &) {}
};

// end-of-query / end-of-target scores of ksw_extz_t, kept in shared memory while the DP runs (they change on few diagonals and would
// otherwise hold five registers for the whole loop)
struct KswBandEz { int mqe, mqe_t, mte, mte_q, score, pad[3]; };

// ALL 32 threads of a warp call this together: 32 / G groups of G threads, one alignment per group (valid = 0 for a group without one),
// the anti-diagonal loop in lockstep over the groups.  The caller has checked ksw_band_serves(qlen, tlen, P.w) for every valid group.
// M.seq (seq_cap bytes) stages the reversed query; M.xvuy points at 1 KB of shared memory for the backtrack tile (it may alias M.seq);
// the last 32 bytes of the group's shared-memory region (M.region_bytes) hold a KswBandEz.
template <int G, bool EZ_FULL>
__device__ void ksw2_band(bool valid, int qlen, const KswQuery query, int tlen, const uint8_t *target, const KswParams P, const KswMem M, KswOut &out)
{
	constexpr int NS = 32 / G;
	static_assert(G == 8, "8 threads per alignment (the tile prefetch deals eight words of a row to a group)");
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
		else if (ksw_seq_bytes(qlen, tlen) + sizeof(KswBandEz) > (size_t)M.seq_cap) { out.status = KSW_ST_SEQCAP; live = false; } // the staging area ends 32 bytes early: KswBandEz
		else if ((size_t)ksw_rows_bound(qlen, tlen, w) * (size_t)pitch + 2 * KSW_PMAT_PAD > M.p_cap) { out.status = KSW_ST_PCAP; live = false; }
		else if ((long long)(qlen + tlen + 2) * qe + (long long)(qlen < tlen ? qlen : tlen) * (P.match > 0 ? P.match : 0) + gbias >= 0xF000) { out.status = KSW_ST_HCAP; live = false; }
	}
	const bool run = live;
	uint8_t *pmat = M.pmat + KSW_PMAT_PAD;
	uint8_t *qrp = M.seq;
	KswBandEz *ez = (KswBandEz*)((unsigned char*)M.xvuy + (M.region_bytes - (int)sizeof(KswBandEz)));
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
	if (gl == 0) { ez->mqe = ez->mte = ez->score = KSW_NEG_INF; ez->mqe_t = ez->mte_q = -1; }
	wild = ksw_gany<G>(wild, lane);
	__syncwarp();

	KswBandDiag dg;
	dg.qrp = (const uint32_t*)qrp + (KSW_QR_PAD >> 2); dg.qlen = qlen; dg.tlen = tlen; dg.w = w; dg.qe = qe;
	dg.fast_ok = P.match + 2 * qe <= 63 && P.q >= 0 && P.q + 2 * P.e + min_sc >= 0;
	const int srcl = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1)); // the thread that owns the word left of mine
	int last_st0 = 0, last_en0 = -1, base = 0;                     // ring = columns [base, base + 128)
	unsigned g_high = 0, cells = 0;
	int st0 = 0, en0 = 0;
	uint8_t *prow = pmat;
	for (int r = 0; ; ++r, prow += pitch) {
		bool act = live && r < nr;
		if (act && st0 > en0) { out.zdropped = 1; live = false; act = false; } // :200-203
		if (!__any_sync(FULL_MASK, act)) break;
		const int st = st0 & ~15, en = en0 | 15;    // :205
		if (act) cells += (unsigned)(en0 - st0 + 1);
		const int bend = st0 + (int)((((unsigned)(en0 - st0) >> 4) + 1u) << 4);
		const int wl_last = en > bend - 1 ? en : bend - 1;
		const int goff = qe * (r + 1) + gbias;
		// the values the neighbours will ask for, from every slot, before any word of this diagonal is updated: top x lane, top v lane,
		// top exact score
		uint32_t C[NS];
#pragma unroll
		for (int s = 0; s < NS; ++s) C[s] = __byte_perm(__byte_perm(x[s], v[s], 0x0073), gy[s], 0x7610);
		{ // which of my slots hold a word of this diagonal
			// my words are gl + G k: those of [st, wl_last] are the rounds k in [kA, kB], at most NS of them, in slots k % NS
			const int kA = ((st >> 2) - gl + G - 1) / G, kB = ((wl_last >> 2) - gl) >= 0 ? ((wl_last >> 2) - gl) / G : -1; // st >= 0: no negative division on the left
			const int nk = act ? kB - kA + 1 : 0;
			const unsigned m = nk > 0 ? (1u << nk) - 1u : 0u, sh = (unsigned)kA % NS;
			const unsigned onm = ((m << sh) | (m << sh >> NS)) & ((1u << NS) - 1u);
			dg.onm = onm; dg.uon = __reduce_or_sync(FULL_MASK, onm);
		}
		{
			// boundary conditions :207-211 (values of the previous diagonal): the word left of st supplies x[st-1], v[st-1] when it was
			// inside the previous rounded band, else x1 = 0 and v1 = q (0 on diagonal 0, and 0 when st > 0)
			const bool keep_prev = st > 0 && st - 1 >= (last_st0 & ~15) && st - 1 <= (last_en0 | 15) && last_en0 >= 0;
			dg.nb_first = keep_prev ? 0xffffffffu : ((st == 0 && r) ? ((uint32_t)(P.q & 0xff) << 8) : 0u);
		}
		dg.st = st; dg.en = en; dg.en0 = en0; dg.st0 = st0; dg.cq = qlen - 1 - r; dg.r = r; dg.last_st0 = last_st0; dg.last_en0 = last_en0;
		dg.A4 = (uint32_t)(st0 - st) * 0x01010101u; dg.B4 = (uint32_t)(bend - st) * 0x01010101u; dg.E4 = (uint32_t)(en0 - st) * 0x01010101u;
		dg.prow = prow; dg.wild = wild;
		// H[en0] is built from the OLD H[en0-1] (:318): fetch it from its owner before this diagonal's updates.  Column word wi always
		// sits in slot (wi / G) % NS of thread wi % G (recycling moves a slot by NS * G words), so the slot is a uniform switch.
		unsigned gprev = 0;
		{
			const int c = en0 > 0 ? en0 - 1 : 0;
			switch (((c >> 2) / G) % NS) {
			case 0: gprev = ksw_band_glane<NS, 0>(c, gx, gy); break;
			case 1: gprev = ksw_band_glane<NS, 1 % NS>(c, gx, gy); break;
			case 2: gprev = ksw_band_glane<NS, 2 % NS>(c, gx, gy); break;
			case 3: gprev = ksw_band_glane<NS, 3 % NS>(c, gx, gy); break;
			case 4: gprev = ksw_band_glane<NS, 4 % NS>(c, gx, gy); break;
			case 5: gprev = ksw_band_glane<NS, 5 % NS>(c, gx, gy); break;
			case 6: gprev = ksw_band_glane<NS, 6 % NS>(c, gx, gy); break;
			default: gprev = ksw_band_glane<NS, 7 % NS>(c, gx, gy); break;
			}
			gprev = __shfl_sync(FULL_MASK, gprev, (lane & ~(G - 1)) | ((c >> 2) & (G - 1)));
			if (act && r > 0 && (c < last_st0 || c > last_en0)) { // the column was not part of the previous band: its offset is older (the band left the matrix on the right)
				int rr = r - 1;
				for (; rr > 0; --rr) { int s_, e_; ksw_band(rr, qlen, tlen, w, s_, e_); if (s_ <= c && c <= e_) break; }
				gprev += (unsigned)(qe * (r - 1 - rr));
			}
			if (r == 0) gprev = (unsigned)qe; // H[0] = v[0] - 2(q+e): g = H + (q+e) + bias = v[0] + (q+e)
		}
		uint32_t bh2 = 0;
		KswBandSlots<G, NS, 0>::run(P, dg, srcl, gl, tcol, x, v, u, y, Sc, T, gx, gy, C, bh2);
		// band max of the updated columns, then the en0 cell (:318 / :349): its owner computes it, everyone needs it
		unsigned mg = (bh2 & 0xffffu) > (bh2 >> 16) ? (bh2 & 0xffffu) : (bh2 >> 16);
		mg = ksw_gmax<G>(mg);
		unsigned ghen = 0;
		if (act && gl == ((en0 >> 2) & (G - 1))) {
			switch (((en0 >> 2) / G) % NS) {
			case 0: ghen = ksw_band_en0<NS, 0>(en0, gprev, u, v, gx, gy); break;
			case 1: ghen = ksw_band_en0<NS, 1 % NS>(en0, gprev, u, v, gx, gy); break;
			case 2: ghen = ksw_band_en0<NS, 2 % NS>(en0, gprev, u, v, gx, gy); break;
			case 3: ghen = ksw_band_en0<NS, 3 % NS>(en0, gprev, u, v, gx, gy); break;
			case 4: ghen = ksw_band_en0<NS, 4 % NS>(en0, gprev, u, v, gx, gy); break;
			case 5: ghen = ksw_band_en0<NS, 5 % NS>(en0, gprev, u, v, gx, gy); break;
			case 6: ghen = ksw_band_en0<NS, 6 % NS>(en0, gprev, u, v, gx, gy); break;
			default: ghen = ksw_band_en0<NS, 7 % NS>(en0, gprev, u, v, gx, gy); break;
			}
		}
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
				const int en1 = st0 + (((en0 - st0) >> 2) << 2);
				const unsigned bandw = (unsigned)(en0 - st0);
				const uint32_t mg2 = mg * 0x10001u;
#pragma unroll
				for (int s = 0; s < NS; ++s) {
					const uint32_t xa = gx[s] ^ mg2, xb = gy[s] ^ mg2; // a zero halfword = a column that holds the band max (if it is in the band)
					if ((((xa - 0x00010001u) & ~xa) | ((xb - 0x00010001u) & ~xb)) & 0x80008000u) {
#pragma unroll
						for (int c = 0; c < 4; ++c) {
							const unsigned val = ((c < 2 ? gx[s] : gy[s]) >> (16 * (c & 1))) & 0xffffu;
							const int t = tcol[s] + c;
							if ((unsigned)(t - st0) < bandw && val == mg) { const unsigned rk = ksw_tie_rank(t, st0, en1); best = rk < best ? rk : best; }
						}
					}
				}
			}
			const unsigned rk = ksw_gmin<G>(best);
			if (need_t) { max_H = mh; max_t = st0 + (int)((rk - 1) & 0xfffffu); }
		}
		if (EZ_FULL) {
			const bool at_tend = act && en0 == tlen - 1, at_qend = act && r - st0 == qlen - 1;
			if (__any_sync(FULL_MASK, at_tend || at_qend)) {
				unsigned gv = 0; // H[st0] of the last query row (:353-354): its owner has it
#pragma unroll
				for (int s = 0; s < NS; ++s) if (tcol[s] == (st0 & ~3)) gv = (((st0 & 2) ? gy[s] : gx[s]) >> (16 * (st0 & 1))) & 0xffffu;
				gv = __shfl_sync(FULL_MASK, gv, (lane & ~(G - 1)) | ((st0 >> 2) & (G - 1)));
				if (gl == 0) {
					if (at_tend) {
						if (hen > ez->mte) { ez->mte = hen; ez->mte_q = r - en; }
						if (r == qlen + tlen - 2) ez->score = hen; // H[tlen-1]
					}
					if (at_qend) {
						const int Hst0 = st0 == en0 ? hen : (int)gv - goff;
						if (Hst0 > ez->mqe) { ez->mqe = Hst0; ez->mqe_t = st0; }
					}
				}
			}
		}
		// band of the next diagonal (:196-199): the block it enters, if any, and its :212 patch are applied now
		int st0n, en0n;
		ksw_band(r + 1, qlen, tlen, w, st0n, en0n);
		const bool nxt = act && r + 1 < nr && st0n <= en0n;
		// the ring must reach 16 lanes past the rounded band (the score overrun): when the next band needs more, recycle the 16 leftmost
		// columns -- dead by then -- for the 16 columns right of the ring; they read as never written (calloc), scores as s = 0
		const bool enter = nxt && base + 128 < (en0n | 15) + 17;
		if (__any_sync(FULL_MASK, enter)) {
			bool w4 = false;
			if (enter) {
				base += 16;
#pragma unroll
				for (int s = 0; s < NS; ++s)
					if (tcol[s] < base) {
						tcol[s] += 128;
						x[s] = v[s] = u[s] = y[s] = 0u; Sc[s] = P.qe2_4;
						T[s] = ksw_target_word(target, tlen, tcol[s]); w4 |= ksw_has4(T[s]);
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
		last_st0 = st0; last_en0 = en0;
		st0 = st0n; en0 = en0n;
	}
	if (g_high) out.status = KSW_ST_HCAP;
	out.cells = (long long)cells;
	__syncwarp();
	if (EZ_FULL && run) { out.mqe = ez->mqe; out.mqe_t = ez->mqe_t; out.mte = ez->mte; out.mte_q = ez->mte_q; out.score = ez->score; }
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

