// indelope_b200/csrc/ksw2.cuh -- kernel 2: ksw2 extension alignment, one warp per alignment.
//
// Replaces ksw_extz2_sse (src/ksw2/csrc/ksw2_extz2_sse.c:113-388, flag == 0) bit for bit: the anti-diagonal
// difference recurrence on wrapping int8 lanes, the 16-lane rounding of the band and the stale lanes it leaves
// behind (they feed real cells), the exact 32-bit max with the SSE 4-accumulator tie
// order, z-drop, and ksw_backtrack (:47-79) to a BAM-style CIGAR.
//
// Mapping: lane l of the warp owns columns t == l (mod 32) of the current anti-diagonal; u,v,x,y,s live in
// shared memory as persistent per-column int8 arrays (they must survive between diagonals, stale values
// included); the x[t-1]/v[t-1] neighbour exchange is a __shfl_up_sync with a carried value between 32-lane
// chunks; H (exact scores) is a ring of the live band in shared memory; the backtrack matrix goes to global
// memory (80-180 KB per alignment, L2 resident).  All arithmetic is integer; no tensor cores.
#pragma once
#include "common.cuh"

struct KswOut {
	int max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score;
	int n_cigar;      // ops left in the scratch buffer, REVERSED (last op first)
	int status;       // 0 ok, 1 early return (:147/:171), negative: capacity
	long long cells;  // exact in-band cells over executed diagonals
};

#define KSW_ST_EARLY 1
#define KSW_ST_TCAP (-2)
#define KSW_ST_PCAP (-3)
#define KSW_ST_CIGCAP (-4)
#define KSW_ST_HCAP (-5)

__device__ __forceinline__ void ksw_reset(KswOut &o) // ksw_reset_extz :81-86
{
	o.max_q = o.max_t = o.mqe_t = o.mte_q = -1;
	o.max = 0; o.score = o.mqe = o.mte = KSW_NEG_INF;
	o.n_cigar = 0; o.zdropped = 0; o.status = 0; o.cells = 0;
}

// bytes of lane storage for targets up to t_cap (multiple of 16): u,v,x,y and s (+16 spare lanes)
__host__ __device__ inline size_t ksw_lane_bytes(int t_cap) { return (size_t)5 * t_cap + 16; }

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

// All 32 lanes call this with identical arguments; `out` comes back identical in every lane.
// sm: this warp's lane storage (ksw_lane_bytes(t_cap) bytes, 16-byte aligned; shared memory, or a global-memory
// spill area for targets that do not fit); H: ring of hr ints (power of two) in shared memory; pmat: backtrack matrix
// workspace of p_cap bytes in global memory; cig: CIGAR scratch of cig_cap ops in global memory.
__device__ void ksw2_warp(int qlen, const uint8_t *query, int tlen, const uint8_t *target, KswParams P,
                          int8_t *sm, int t_cap, int *H, int hr, uint8_t *pmat, size_t p_cap, uint32_t *cig, int cig_cap, KswOut &out)
{
	const int lane = lane_id();
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
	int n_col = qlen < tlen ? qlen : tlen; // :164-165 (bytes)
	const int bandmax = n_col < w + 1 ? n_col : w + 1;
	n_col = ((bandmax + 15) / 16 + 1) * 16;
	if (T16 > t_cap) { out.status = KSW_ST_TCAP; return; }
	if (bandmax + 2 > hr) { out.status = KSW_ST_HCAP; return; }
	if ((size_t)(qlen + tlen - 1) * (size_t)n_col > p_cap) { out.status = KSW_ST_PCAP; return; }
	const int hmask = hr - 1;
	const int8_t qe2 = (int8_t)(qe * 2), max_sc8 = (int8_t)(P.match + qe * 2);
	int8_t *u = sm, *v = u + t_cap, *x = v + t_cap, *y = x + t_cap, *s = y + t_cap;

	// calloc :173 -- lanes that were never computed must read as zero
	for (int a = 0; a < 5; ++a) {
		uint32_t *z = (uint32_t*)(sm + a * t_cap);
		const int nwords = (T16 + (a == 4 ? 16 : 0)) >> 2;
		for (int i = lane; i < nwords; i += 32) z[i] = 0;
	}
	__syncwarp();

	int last_st = -1, last_en = -1;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st0, en0;
		ksw_band(r, qlen, tlen, w, st0, en0);
		if (st0 > en0) { out.zdropped = 1; break; } // :200-203
		const int st = st0 & ~15, en = en0 | 15;    // :205
		out.cells += en0 - st0 + 1;
		// boundary conditions :207-212 (values of the previous diagonal)
		int x1, v1;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) { x1 = x[st - 1]; v1 = v[st - 1]; }
			else x1 = v1 = 0;
		} else { x1 = 0; v1 = r ? P.q : 0; }
		if (en >= r && lane == 0) { y[r] = 0; u[r] = r ? P.q : 0; }
		// scores: 16-lane blocks anchored at the exact st0 (:215-228); lanes outside keep stale s
		{
			const int total = ((en0 - st0) / 16 + 1) * 16;
			for (int i = lane; i < total; i += 32) {
				const int tt = st0 + i;
				const int sq = tt < tlen ? target[tt] : 0;
				const int sq2 = tt <= r ? query[r - tt] : 0; // qr[qlen-1-r+tt], zero padded past qlen
				s[tt] = (sq == 4 || sq2 == 4) ? 0 : (sq == sq2 ? P.match : P.mismatch);
			}
		}
		__syncwarp();
		// core lanes st..en (:262-284)
		{
			uint8_t *pr = pmat + (size_t)r * n_col - st;
			int xc = x1, vc = v1;
			for (int t0 = st; t0 <= en; t0 += 32) {
				const int t = t0 + lane;
				const bool act = t <= en;
				int xo = 0, vo = 0, ut = 0, yt = 0, sc = 0;
				if (act) { xo = x[t]; vo = v[t]; ut = u[t]; yt = y[t]; sc = s[t]; }
				int xt1 = __shfl_up_sync(FULL_MASK, xo, 1), vt1 = __shfl_up_sync(FULL_MASK, vo, 1);
				if (lane == 0) { xt1 = xc; vt1 = vc; }
				xc = __shfl_sync(FULL_MASK, xo, 31); vc = __shfl_sync(FULL_MASK, vo, 31);
				if (act) {
					int8_t z = (int8_t)(sc + qe2);
					int8_t a = (int8_t)(xt1 + vt1);
					int8_t b = (int8_t)(yt + ut);
					int d = a > z ? 1 : 0;
					z = a > z ? a : z;                                   // signed max
					d = b > z ? 2 : d;                                   // signed compare
					z = (uint8_t)z > (uint8_t)b ? z : b;                 // unsigned max
					z = (uint8_t)z < (uint8_t)max_sc8 ? z : max_sc8;     // unsigned min
					u[t] = (int8_t)(z - vt1);
					v[t] = (int8_t)(z - ut);
					z = (int8_t)(z - P.q);
					a = (int8_t)(a - z);
					b = (int8_t)(b - z);
					x[t] = a > 0 ? a : (int8_t)0; if (a > 0) d |= 0x08;
					y[t] = b > 0 ? b : (int8_t)0; if (b > 0) d |= 0x10;
					pr[t] = (uint8_t)d;
				}
			}
		}
		__syncwarp();
		// exact max :312-357
		int max_H, max_t, Hen0, Hst0;
		if (r > 0) {
			const int hen = en0 > 0 ? H[(en0 - 1) & hmask] + (int)(uint8_t)u[en0] - qe : H[en0 & hmask] + (int)(uint8_t)v[en0] - qe;
			__syncwarp(); // every lane has read the old H[en0-1] before anyone updates
			const int en1 = st0 + ((en0 - st0) / 4) * 4;
			int bh = hen; unsigned br = 0; // rank 0 = the initial candidate (H[en0], en0): wins every tie
			for (int t = st0 + lane; t < en0; t += 32) {
				const int h = H[t & hmask] + (int)(uint8_t)v[t] - qe;
				H[t & hmask] = h;
				// tie order of the SSE code: 4 strided accumulators (lower accumulator, then lower t), then the scalar tail
				const unsigned rank = 1u + ((t < en1 ? (unsigned)((t - st0) & 3) : 4u) << 20) + (unsigned)(t - st0);
				if (h > bh || (h == bh && rank < br)) { bh = h; br = rank; }
			}
			if (lane == 0) H[en0 & hmask] = hen;
			const unsigned key = (unsigned)bh ^ 0x80000000u;
			const unsigned mk = __reduce_max_sync(FULL_MASK, key);
			const unsigned mr = __reduce_min_sync(FULL_MASK, key == mk ? br : 0xffffffffu);
			max_H = (int)(mk ^ 0x80000000u);
			max_t = mr == 0 ? en0 : st0 + (int)((mr - 1) & 0xfffffu);
			Hen0 = hen;
			__syncwarp();
			Hst0 = H[st0 & hmask];
		} else {
			const int h0 = (int)(uint8_t)v[0] - qe - qe;
			if (lane == 0) H[0] = h0;
			max_H = h0; max_t = 0; Hen0 = h0; Hst0 = h0;
			__syncwarp();
		}
		if (en0 == tlen - 1 && Hen0 > out.mte) { out.mte = Hen0; out.mte_q = r - en; }
		if (r - st0 == qlen - 1 && Hst0 > out.mqe) { out.mqe = Hst0; out.mqe_t = st0; }
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
		if (r == qlen + tlen - 2 && en0 == tlen - 1) out.score = Hen0; // H[tlen-1]
		last_st = st; last_en = en;
	}
	__syncwarp();
	// backtrack :380-385 -> ksw_backtrack :47-79 (is_rot = 1, left-aligned gaps)
	int i, j;
	if (!out.zdropped) { i = tlen - 1; j = qlen - 1; }
	else if (out.max_t >= 0 && out.max_q >= 0) { i = out.max_t; j = out.max_q; }
	else return;
	int n = 0, ovf = 0;
	if (lane == 0) {
		int state = 0;
		unsigned cur_op = 0xffu, cur_len = 0;
		while (i >= 0 && j >= 0) {
			const int r = i + j;
			int st0, en0;
			ksw_band(r, qlen, tlen, w, st0, en0);
			const int off = st0 & ~15, off_end = en0 | 15;
			int force_state = -1;
			if (i < off) force_state = 2;
			if (i > off_end) force_state = 1;
			const unsigned tmp = force_state < 0 ? pmat[(size_t)r * n_col + i - off] : 0u;
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
	n = __shfl_sync(FULL_MASK, n, 0);
	ovf = __shfl_sync(FULL_MASK, ovf, 0);
	out.n_cigar = n;
	if (ovf) out.status = KSW_ST_CIGCAP;
	__syncwarp();
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
