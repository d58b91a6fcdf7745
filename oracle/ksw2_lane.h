/* oracle/ksw2_lane.h -- TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
 *
 * Plain-C, lane-by-lane restatement of the reference's SSE extension alignment
 * (reference: src/ksw2/csrc/ksw2_extz2_sse.c:113-388, flag==0 branch only, plus
 * ksw_backtrack :47-79, ksw_push_cigar :31-41, ksw_apply_zdrop :88-104).
 * Parity is pinned against the reference C file itself compiled into
 * oracle/_ref/libksw2_ref.so (see oracle/Makefile, tests/test_oracle_ksw2.py).
 */
#ifndef ORC_KSW2_LANE_H
#define ORC_KSW2_LANE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NEG_INF (-0x40000000)

/* the fields of the reference's ksw_extz_t (src/ksw2/csrc/ksw2.h:22-30), unpacked */
typedef struct {
	int32_t max, zdropped;
	int32_t max_q, max_t;
	int32_t mqe, mqe_t;
	int32_t mte, mte_q;
	int32_t score;
	int32_t n_cigar;     /* full CIGAR length (BAM-style ops, len<<4|op) */
	int64_t cells;       /* oracle counter: exact in-band cells over executed diagonals */
	int32_t diagonals;   /* executed anti-diagonals */
	int32_t status;      /* 0 ok, 1 early return (:147 or :171), -1 cigar buffer too small */
} orc_ez_t;

/* query/target are 0..4 codes (4 = wildcard), match/mismatch fill the 5x5 matrix
 * as src/ksw2/ksw2.nim:135-140 does. Only flag==0 is modelled. */
void orc_ksw2_lane(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int8_t match, int8_t mismatch, int8_t q, int8_t e, int w, int zdrop,
                   orc_ez_t *ez, uint32_t *cigar, int cigar_cap);

#ifdef __cplusplus
}
#endif
#endif
