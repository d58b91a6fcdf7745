/* oracle/ref_shim.c -- TEST INFRASTRUCTURE. Thin adapter compiled TOGETHER WITH the
 * reference's own, unmodified src/ksw2/csrc/ksw2_extz2_sse.c (taken from /root/reference at
 * build time, never copied into this repo) into oracle/_ref/libksw2_ref.so.  It flattens the
 * reference's ksw_extz_t (src/ksw2/csrc/ksw2.h:22-30) into orc_ez_t so Python/ctypes and the
 * oracle can call the real reference DP.  The 5x5 matrix is the one src/ksw2/ksw2.nim:135-140
 * builds. */
#include <stdlib.h>
#include <string.h>
#include "ksw2.h"      /* the reference header, -I/root/reference/src/ksw2/csrc */
#include "ksw2_lane.h"

void orc_ksw2_ref(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                  int8_t match, int8_t mismatch, int8_t q, int8_t e, int w, int zdrop, int flag,
                  orc_ez_t *out, uint32_t *cigar, int cigar_cap)
{
	int8_t mat[25];
	int i, j;
	ksw_extz_t ez;
	memset(&ez, 0, sizeof(ez));
	for (i = 0; i < 5; ++i)
		for (j = 0; j < 5; ++j)
			mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : (i == j ? match : mismatch);
	ksw_extz2_sse(0, qlen, query, tlen, target, 5, mat, q, e, w, zdrop, flag, &ez);
	out->max = ez.max; out->zdropped = ez.zdropped;
	out->max_q = ez.max_q; out->max_t = ez.max_t;
	out->mqe = ez.mqe; out->mqe_t = ez.mqe_t;
	out->mte = ez.mte; out->mte_q = ez.mte_q;
	out->score = ez.score; out->n_cigar = ez.n_cigar;
	out->cells = 0; out->diagonals = 0; out->status = 0;
	if (ez.n_cigar > cigar_cap) out->status = -1;
	else if (ez.n_cigar > 0) memcpy(cigar, ez.cigar, sizeof(uint32_t) * ez.n_cigar);
	free(ez.cigar);
}
