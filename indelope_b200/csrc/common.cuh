// indelope_b200/csrc/common.cuh -- shared device/host definitions of libindelope_cuda.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "indelope_cuda.h"

#define IDL_WARP 32
#define FULL_MASK 0xffffffffu
#define KSW_NEG_INF (-0x40000000)

// device-side global counters of one ticket (zeroed before the chain starts)
struct DevCounters {
	unsigned int region_next;      // work queue heads
	unsigned int region_next2;
	unsigned int aln_next;
	unsigned int al_next;
	unsigned int n_contigs;        // bump allocators of the result pools
	unsigned int n_contig_bases;
	unsigned int n_alns;
	unsigned int n_events;
	unsigned int n_cigar_ops;
	unsigned int n_al_items;       // AL fallback: reads that passed the window tests (two DP tasks each)
	unsigned long long al_pack;    // (#AL events << 40) | total AL work items
	unsigned int overflow;         // result pool overflow flags
	unsigned long long offsets_tested, dp_cells_a, dp_cells_b, dp_a, dp_b, kmer_reads, kmer_bytes, al_events;
	unsigned int n_regions_in, pad_;  // regions of the batch (written by region_key_kernel: the count sort_scatter_kernel reads)
};

// alignment scoring/band parameters of one call-site (src/ksw2/ksw2.nim:142,151-157)
struct KswParams {
	int8_t match, mismatch, q, e;
	int w, zdrop;
	// byte-replicated constants of the int8 recurrence, made once on the host so that the kernels read them straight from
	// the constant bank: 2(q+e), match + 2(q+e) (also the clamp), q, mismatch + 2(q+e), (match + 2(q+e)) | 0x80
	uint32_t qe2_4, maxsc_4, q_4, misq_4, maxsc_h80;
};
inline KswParams ksw_make_params(int match, int mismatch, int q, int e, int w, int zdrop)
{
	KswParams p;
	p.match = (int8_t)match; p.mismatch = (int8_t)mismatch; p.q = (int8_t)q; p.e = (int8_t)e; p.w = w; p.zdrop = zdrop;
	const int qe = p.q + p.e;
	auto rep = [](int v) { return (uint32_t)(v & 0xff) * 0x01010101u; };
	p.qe2_4 = rep(qe * 2); p.maxsc_4 = rep(p.match + qe * 2); p.q_4 = rep(p.q); p.misq_4 = rep(p.mismatch + qe * 2);
	p.maxsc_h80 = p.maxsc_4 | 0x80808080u;
	return p;
}

#define SORT_BUCKETS 16384
// bucket sort of alignment tasks, longest first: hist/start/cursor hold SORT_BUCKETS entries.  Kernel 2 runs the four
// alignments of a warp in lockstep and takes its cheap interior path only when the words of all four lie inside their
// bands, so the key carries the whole SHAPE of the task: call-site A (banded: the band depends on the diagonal alone
// until the sequences end) sorts by the exact number of anti-diagonals, call-site B (unbanded: the band is
// [r - qlen + 1, min(r, tlen - 1)]) by anti-diagonals and query length, so neighbours have the same (qlen, tlen).
struct SortBufs { unsigned *hist, *start, *cursor; uint16_t *keys; unsigned *order; };
__device__ __forceinline__ uint16_t sort_key_a(int diagonals) { return (uint16_t)(diagonals > SORT_BUCKETS - 1 ? SORT_BUCKETS - 1 : (diagonals < 0 ? 0 : diagonals)); }
__device__ __forceinline__ uint16_t sort_key_b(int diagonals, int qlen)
{
	const int d = diagonals > 1023 ? 1023 : (diagonals < 0 ? 0 : diagonals);
	return (uint16_t)((d << 4) | (qlen & 15));
}
// executed anti-diagonals of one extension alignment are bounded by the band running out (ksw2_extz2_sse.c:196-203)
__device__ __forceinline__ int est_diagonals(int qlen, int tlen, int w)
{
	if (qlen <= 0 || tlen <= 0) return 0;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	int d = qlen + tlen - 1;
	if (d > 2 * qlen + w) d = 2 * qlen + w;
	if (d > 2 * tlen + w) d = 2 * tlen + w;
	return d;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// 32 bits of a bit-plane starting at bit position `pos` (pos may be negative: missing low bits read as 0).
// The caller guarantees one padding word after the last word it can touch.
// the same for a position known to be >= 0 (no branch: the inner loops of the assembler are instruction-cache bound)
__device__ __forceinline__ uint32_t get32p(const uint32_t *plane, int pos)
{
	const int w = pos >> 5;
	return __funnelshift_r(plane[w], plane[w + 1], pos & 31);
}
__device__ __forceinline__ uint32_t get32(const uint32_t *plane, int pos)
{
	if (pos >= 0) {
		const int w = pos >> 5, sh = pos & 31;
		return __funnelshift_r(plane[w], plane[w + 1], sh);
	}
	const int sh = -pos;
	return sh < 32 ? plane[0] << sh : 0u;
}
