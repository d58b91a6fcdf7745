/* oracle/ksw2_lane.c -- TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
 *
 * Lane-by-lane scalar restatement of the reference's 16-lane SSE kernel
 * src/ksw2/csrc/ksw2_extz2_sse.c (flag == 0: with CIGAR, gaps left-aligned,
 * exact max).  One "lane" here is one int8 element of the reference's __m128i
 * vectors; the 16-lane block rounding of the band, the stale values that the
 * SSE code leaves in lanes outside the exact band, and the 4-accumulator
 * arg-max tie order are all reproduced because they change results
 * (SURVEY.md appendix B).  Pinned against the compiled reference in
 * tests/test_oracle_ksw2.py.
 */
#include <stdlib.h>
#include <string.h>
#include "ksw2_lane.h"

static inline int8_t max_s8(int8_t a, int8_t b) { return a > b ? a : b; }
static inline int8_t max_u8(int8_t a, int8_t b) { return (uint8_t)a > (uint8_t)b ? a : b; }
static inline int8_t min_u8(int8_t a, int8_t b) { return (uint8_t)a < (uint8_t)b ? a : b; }

/* ksw_reset_extz, ksw2_extz2_sse.c:81-86 */
static void reset_ez(orc_ez_t *ez)
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0; ez->score = ez->mqe = ez->mte = ORC_NEG_INF;
	ez->n_cigar = 0; ez->zdropped = 0;
	ez->cells = 0; ez->diagonals = 0; ez->status = 0;
}

/* ksw_apply_zdrop with is_rot=1, ksw2_extz2_sse.c:88-104 */
static int apply_zdrop(orc_ez_t *ez, int32_t H, int r, int t, int zdrop, int8_t e)
{
	if (H > ez->max) {
		ez->max = H; ez->max_t = t; ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
		l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && ez->max - H > zdrop + l * e) {
			ez->zdropped = 1;
			return 1;
		}
	}
	return 0;
}

/* ksw_push_cigar, :31-41 (growth policy is irrelevant to results) */
static int push_cigar(uint32_t *cigar, int cap, int *n, uint32_t op, int len)
{
	if (*n == 0 || op != (cigar[*n - 1] & 0xf)) {
		if (*n == cap) return -1;
		cigar[(*n)++] = (uint32_t)len << 4 | op;
	} else cigar[*n - 1] += (uint32_t)len << 4;
	return 0;
}

/* ksw_backtrack with is_rot=1, is_rev=0, with_N=0, :47-79 */
static int backtrack(const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0,
                     uint32_t *cigar, int cap, int *n_cigar_)
{
	int n = 0, i = i0, j = j0, r, state = 0, k, rc = 0;
	uint32_t tmp;
	while (i >= 0 && j >= 0) {
		int force_state = -1;
		r = i + j;
		if (i < off[r]) force_state = 2;
		if (i > off_end[r]) force_state = 1;
		tmp = force_state < 0 ? p[(size_t)r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (force_state >= 0) state = force_state;
		if (state == 0) rc |= push_cigar(cigar, cap, &n, 0, 1), --i, --j;
		else if (state == 1 || state == 3) rc |= push_cigar(cigar, cap, &n, 2, 1), --i;
		else rc |= push_cigar(cigar, cap, &n, 1, 1), --j;
		if (rc) return -1;
	}
	if (i >= 0) rc |= push_cigar(cigar, cap, &n, 2, i + 1);
	if (j >= 0) rc |= push_cigar(cigar, cap, &n, 1, j + 1);
	if (rc) return -1;
	for (k = 0; k < n >> 1; ++k)
		tmp = cigar[k], cigar[k] = cigar[n - 1 - k], cigar[n - 1 - k] = tmp;
	*n_cigar_ = n;
	return 0;
}

void orc_ksw2_lane(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int8_t match, int8_t mismatch, int8_t q, int8_t e, int w, int zdrop,
                   orc_ez_t *ez, uint32_t *cigar, int cigar_cap)
{
	int r, t, qe = q + e, n_col, T16, last_st, last_en, min_sc;
	int8_t qe2 = (int8_t)((q + e) * 2), max_sc8 = (int8_t)(match + (q + e) * 2);
	int8_t *u, *v, *x, *y, *s;
	uint8_t *sf, *qr, *p;
	int32_t *H;
	int *off, *off_end;

	reset_ez(ez);
	if (qlen <= 0 || tlen <= 0) { ez->status = 1; return; }            /* :147 (m is always 5) */
	if (w < 0) w = tlen > qlen ? tlen : qlen;                            /* :161 */
	T16 = (tlen + 15) / 16 * 16;
	n_col = qlen < tlen ? qlen : tlen;                                   /* :164-165, in bytes here */
	n_col = (((n_col < w + 1 ? n_col : w + 1) + 15) / 16 + 1) * 16;
	min_sc = mismatch < 0 ? mismatch : 0;                                /* the matrix holds match, mismatch and 0 */
	if (match < min_sc) min_sc = match;
	if (-min_sc > 2 * (q + e)) { ez->status = 1; return; }               /* :171 */

	/* persistent, zero-initialised lanes (calloc :173); s and sf get 16 spare lanes
	 * for the unaligned 16-wide score blocks (:215-228) */
	u = (int8_t*)calloc((size_t)T16 * 4 + (T16 + 16), 1);
	v = u + T16; x = v + T16; y = x + T16; s = y + T16;
	sf = (uint8_t*)calloc((size_t)T16 + 32, 1);
	qr = (uint8_t*)calloc((size_t)qlen + 32, 1);
	H = (int32_t*)malloc(sizeof(int32_t) * T16);
	for (t = 0; t < T16; ++t) H[t] = ORC_NEG_INF;                        /* :177-178 */
	p = (uint8_t*)malloc((size_t)(qlen + tlen - 1) * n_col);
	off = (int*)malloc(sizeof(int) * 2 * (qlen + tlen - 1));
	off_end = off + qlen + tlen - 1;
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];             /* :187 */
	memcpy(sf, target, tlen);                                            /* :188 */

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, v1;
		const int qro = qlen - 1 - r;              /* qrr[t] = qr[qro + t]; t >= st0 >= -qro */
		uint8_t *u8 = (uint8_t*)u, *v8 = (uint8_t*)v;
		uint8_t *pr;
		/* :196-199 */
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) { ez->zdropped = 1; break; }                        /* :200-203 */
		st0 = st; en0 = en;
		st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;                 /* :205 */
		ez->cells += en0 - st0 + 1; ez->diagonals = r + 1;
		/* boundary conditions :207-212 */
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], v1 = v[st - 1];
			else x1 = v1 = 0;
		} else x1 = 0, v1 = r ? q : 0;
		if (en >= r) y[r] = 0, u[r] = r ? q : 0;
		/* scores, 16-lane blocks starting at the exact st0 (:215-228) */
		for (t = st0; t <= en0; t += 16) {
			int k;
			for (k = 0; k < 16; ++k) {
				uint8_t sq = sf[t + k], sq2 = qr[qro + t + k];
				s[t + k] = (sq == 4 || sq2 == 4) ? 0 : (sq == sq2 ? match : mismatch);
			}
		}
		/* core lanes st..en (:262-284); every input is the previous diagonal's value */
		off[r] = st; off_end[r] = en;
		pr = p + (size_t)r * n_col - st;
		{
			int8_t xc = x1, vc = v1; /* x[t-1], v[t-1] of the previous diagonal */
			for (t = st; t <= en; ++t) {
				int8_t z, a, b, d, xt1 = xc, vt1 = vc, ut = u[t];
				xc = x[t]; vc = v[t];
				z = (int8_t)(s[t] + qe2);
				a = (int8_t)(xt1 + vt1);
				b = (int8_t)(y[t] + ut);
				d = a > z ? 1 : 0;
				z = max_s8(z, a);
				d = b > z ? 2 : d;
				z = max_u8(z, b);
				z = min_u8(z, max_sc8);
				u[t] = (int8_t)(z - vt1);
				v[t] = (int8_t)(z - ut);
				z = (int8_t)(z - q);
				a = (int8_t)(a - z);
				b = (int8_t)(b - z);
				x[t] = a > 0 ? a : 0; if (a > 0) d |= 0x08;
				y[t] = b > 0 ? b : 0; if (b > 0) d |= 0x10;
				pr[t] = (uint8_t)d;
			}
		}
		/* exact max :312-357 */
		{
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u8[en0] - qe : H[en0] + v8[en0] - qe;
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += (int32_t)v8[t + i] - qe;
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (t = en1; t < en0; ++t) {
					H[t] += (int32_t)v8[t] - qe;
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v8[0] - qe - qe, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (apply_zdrop(ez, max_H, r, max_t, zdrop, e)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		}
		last_st = st; last_en = en;
	}
	/* :380-385 */
	{
		int rc = 0;
		if (!ez->zdropped)
			rc = backtrack(p, off, off_end, n_col, tlen - 1, qlen - 1, cigar, cigar_cap, &ez->n_cigar);
		else if (ez->max_t >= 0 && ez->max_q >= 0)
			rc = backtrack(p, off, off_end, n_col, ez->max_t, ez->max_q, cigar, cigar_cap, &ez->n_cigar);
		if (rc) ez->status = -1;
	}
	free(u); free(sf); free(qr); free(H); free(p); free(off);
}
