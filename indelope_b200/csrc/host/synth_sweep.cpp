// indelope_b200/csrc/host/synth_sweep.cpp
//
// (1) Seeded synthetic data: reference, planted indels / tandem-repeat events, reads with an aligner
//     model (no aligner or BAM library exists in this image; SURVEY.md 8d describes the model).
// (2) The host sweep that stays outside the GPU library: a C++ stand-in for the reference's
//     gen_roi (src/indelope.nim:430-545).  Same read order, same chunking, same region bounds.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "indelope_host.h"
#include "dataset.h"

namespace {

// ---- deterministic RNG (xoshiro256**, seeded by splitmix64): identical streams on every box ----
struct Rng {
	uint64_t s[4];
	static uint64_t splitmix(uint64_t &x) { uint64_t z = (x += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
	explicit Rng(uint64_t seed) { for (auto &v : s) v = splitmix(seed); }
	static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
	uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
	double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
	int64_t below(int64_t n) { return n <= 0 ? 0 : (int64_t)(next() % (uint64_t)n); }
	int64_t range(int64_t lo, int64_t hi) { return lo + below(hi - lo + 1); } // inclusive
	int poisson_exp(double expl) { // Knuth with exp(-lambda) precomputed; lambda is small (coverage / read_len)
		double p = 1; int k = 0;
		do { ++k; p *= uni(); } while (p > expl);
		return k - 1;
	}
};

using Event = IdlhEvent;
using ReadRec = IdlhReadRec;

} // namespace

namespace {

const char ACGT[] = "ACGT";

void emit_read(idlh_dataset &D, Rng &rng, int chrom, const std::string &seq0, int64_t start, int64_t ref_consumed, const std::vector<uint32_t> &cig)
{
	const idlh_synth_params &P = D.P;
	std::string seq = seq0;
	// sparse errors: jump from one hit to the next with geometric gaps instead of one draw per base
	auto sprinkle = [&](double rate, bool to_n) {
		if (rate <= 0) return;
		const double lg = std::log1p(-rate);
		for (double at = std::floor(std::log(1.0 - rng.uni()) / lg); at < (double)seq.size(); at += 1.0 + std::floor(std::log(1.0 - rng.uni()) / lg)) {
			char &c = seq[(size_t)at];
			if (to_n) c = 'N';
			else { char o = c; while (o == c) o = ACGT[rng.below(4)]; c = o; }
		}
	};
	sprinkle(P.sub_rate, false);
	sprinkle(P.n_base_rate, true);
	ReadRec r;
	r.chrom = chrom; r.start = (int32_t)start; r.stop = (int32_t)(start + ref_consumed); r.len = (int32_t)seq.size();
	r.mapq = 60;
	if (rng.uni() < P.low_mapq_fraction) { static const uint8_t lows[4] = {0, 5, 10, 19}; r.mapq = lows[rng.below(4)]; }
	r.flag = (uint16_t)(rng.below(2) ? 16 : 0);
	if (rng.uni() < P.dup_fraction) r.flag |= 0x400;
	r.seq_off = (int64_t)D.bases.size(); r.cig_off = (int64_t)D.cigars.size(); r.n_cig = (int32_t)cig.size();
	r.order = D.reads.size();
	D.bases.insert(D.bases.end(), seq.begin(), seq.end());
	size_t q0 = D.quals.size();
	D.quals.resize(q0 + seq.size(), 30);
	if (P.qual_levels > 1) {   // per-base qualities from a hash: the main random stream (and with it every read) is unchanged
		uint64_t h = (uint64_t)D.reads.size() * 0x9E3779B97F4A7C15ULL + P.seed;
		for (size_t i = 0; i < seq.size(); ++i) {
			h ^= h >> 27; h *= 0x3C79AC492BA7B653ULL; h ^= h >> 33; h += 0x1C69B3F74AC4AE35ULL;
			D.quals[q0 + i] = (uint8_t)(15 + (h >> 40) % (uint64_t)P.qual_levels * 26 / (uint64_t)(P.qual_levels - 1));
		}
	}
	if (rng.uni() < P.lowq_tail_fraction) {
		int n = (int)rng.range(1, 15);
		if (n > (int)seq.size()) n = (int)seq.size();
		bool at_start = rng.uni() < 0.3;
		for (int i = 0; i < n; ++i) D.quals[q0 + (at_start ? (size_t)i : seq.size() - 1 - (size_t)i)] = 2;
	}
	D.cigars.insert(D.cigars.end(), cig.begin(), cig.end());
	D.reads.push_back(r);
}

inline uint32_t cg(int len, int op) { return (uint32_t)len << 4 | (uint32_t)op; }

// Build one read of haplotype `hap` whose first base is reference base `s` (k < 0) or base k of the insertion
// of event `ev` (k >= 0), and "align" it.  At most one event can fall inside a read (events are spaced apart).
void make_read(idlh_dataset &D, Rng &rng, int chrom, const std::vector<uint8_t> &ref, const Event *ev, bool carries, int64_t s, int k)
{
	const idlh_synth_params &P = D.P;
	const int L = P.read_len;
	const int64_t clen = (int64_t)ref.size();
	std::string seq; seq.reserve(L);
	int64_t p = s; int lf = 0, in_ins = 0, rf = 0; bool hit = false;
	if (k >= 0) { // starts inside the insertion
		hit = true;
		for (int i = k; i < (int)ev->ins.size() && (int)seq.size() < L; ++i) { seq.push_back(ev->ins[i]); ++in_ins; }
		p = ev->pos + ev->dlen; // an event with both ins and del replaces bases
	}
	while ((int)seq.size() < L && p < clen) {
		if (carries && !hit && ev && p == ev->pos && lf > 0) { // a read starting AT pos begins after the event
			hit = true;
			for (int i = 0; i < (int)ev->ins.size() && (int)seq.size() < L; ++i) { seq.push_back(ev->ins[i]); ++in_ins; }
			p += ev->dlen;
			continue;
		}
		seq.push_back((char)ref[p]); ++p;
		if (hit) ++rf; else ++lf;
	}
	if ((int)seq.size() < L) return; // ran off the contig end
	std::vector<uint32_t> cig;
	int64_t start = s, consumed = 0;
	if (!hit || (in_ins == 0 && (ev == nullptr || ev->dlen == 0))) {
		cig.push_back(cg(L, 0)); consumed = L;
	} else if (!ev->ins.empty()) { // insertion
		const int ilen = (int)ev->ins.size();
		if (lf > 0 && rf > 0 && in_ins == ilen && ilen <= P.max_cigar_indel && lf >= P.min_cigar_flank && rf >= P.min_cigar_flank) {
			cig = {cg(lf, 0), cg(ilen, 1), cg(rf, 0)}; consumed = lf + rf;
		} else if (lf > 0 && (rf == 0 || lf >= rf)) {
			cig = {cg(lf, 0), cg(in_ins + rf, 4)}; consumed = lf;
		} else if (rf > 0) {
			cig = {cg(lf + in_ins, 4), cg(rf, 0)}; start = ev->pos; consumed = rf;
		} else return; // entirely inside the insertion: would not map
	} else { // deletion
		const int dlen = ev->dlen;
		if (rf == 0) { cig.push_back(cg(L, 0)); consumed = L; }
		else if (dlen <= P.max_cigar_indel && lf >= P.min_cigar_flank && rf >= P.min_cigar_flank) {
			cig = {cg(lf, 0), cg(dlen, 2), cg(rf, 0)}; consumed = lf + dlen + rf;
		} else if (lf >= rf) { cig = {cg(lf, 0), cg(rf, 4)}; consumed = lf; }
		else { cig = {cg(lf, 4), cg(rf, 0)}; start = ev->pos + dlen; consumed = rf; }
	}
	emit_read(D, rng, chrom, seq, start, consumed, cig);
}

} // namespace

extern "C" {

void idlh_default_synth(idlh_synth_params *p)
{
	memset(p, 0, sizeof *p);
	p->seed = 20171101; p->n_chroms = 1; p->chrom_len = 1000000; p->n_events = 200; p->min_indel = 5; p->max_indel = 300;
	p->coverage = 30; p->read_len = 150; p->sub_rate = 0.001; p->tr_fraction = 0; p->tr_max_unit = 6; p->het_fraction = 0.5;
	p->lowq_tail_fraction = 0.10; p->low_mapq_fraction = 0.05; p->dup_fraction = 0.01; p->n_base_rate = 0;
	p->locus_only = 0; p->locus_flank = 0; p->max_cigar_indel = 30; p->min_cigar_flank = 20;
}

idlh_dataset *idlh_synth(const idlh_synth_params *pp)
{
	idlh_dataset *Dp = new idlh_dataset();
	idlh_dataset &D = *Dp;
	D.P = *pp;
	const idlh_synth_params &P = D.P;
	const int L = P.read_len;
	const int flank = P.locus_flank > 0 ? P.locus_flank : 2 * L + P.max_indel;
	{ // reserve: avoids re-copying gigabytes while the read table grows
		const double per_base = P.coverage / (double)L;
		const double span = P.locus_only ? (double)P.n_events * (2.0 * flank + (P.min_indel + P.max_indel) / 2.0) : (double)P.chrom_len;
		const size_t est = (size_t)(span * per_base * P.n_chroms * 1.1) + 1024;
		D.reads.reserve(est); D.bases.reserve(est * (size_t)L); D.quals.reserve(est * (size_t)L); D.cigars.reserve(est * 2);
	}
	for (int c = 0; c < P.n_chroms; ++c) {
		Rng rng(P.seed * 1000003ULL + (uint64_t)(P.chrom_first + c) * 7919ULL + 17);
		D.names.push_back("chrS" + std::to_string(P.chrom_first + c + 1));
		D.chroms.emplace_back((size_t)P.chrom_len);
		std::vector<uint8_t> &ref = D.chroms.back();
		for (int64_t i = 0; i < P.chrom_len; i += 32) { // 2 bits per base out of each 64-bit draw
			uint64_t x = rng.next();
			for (int j = 0; j < 32 && i + j < P.chrom_len; ++j) ref[i + j] = (uint8_t)ACGT[(x >> (2 * j)) & 3];
		}
		// events: evenly spread with jitter, at least 2 read lengths + the longest event apart
		std::vector<Event> evs;
		const int64_t spacing = P.n_events > 0 ? P.chrom_len / (P.n_events + 1) : 0;
		for (int e = 0; e < P.n_events; ++e) {
			Event ev; ev.chrom = c; ev.dlen = 0; ev.tr = rng.uni() < P.tr_fraction; ev.hom = rng.uni() >= P.het_fraction;
			int64_t jit = spacing / 4 > 0 ? rng.range(-spacing / 4, spacing / 4) : 0;
			ev.pos = (e + 1) * spacing + jit;
			if (ev.pos < 2 * L + 700 || ev.pos > P.chrom_len - 2 * L - 1400) continue;
			if (ev.tr) { // plant a tandem repeat in the reference, then expand/contract it by whole units
				int unit = (int)rng.range(1, P.tr_max_unit), copies = (int)rng.range(6, 24), delta = (int)rng.range(1, 10);
				std::string u; for (int i = 0; i < unit; ++i) u.push_back(ACGT[rng.below(4)]);
				for (int i = 0; i < unit * copies; ++i) ref[ev.pos + i] = (uint8_t)u[i % unit];
				if (rng.below(2) && delta < copies - 1) ev.dlen = unit * delta;
				else for (int i = 0; i < unit * delta; ++i) ev.ins.push_back(u[i % unit]);
			} else {
				int len = (int)rng.range(P.min_indel, P.max_indel);
				if (rng.below(2)) ev.dlen = len;
				else for (int i = 0; i < len; ++i) ev.ins.push_back(ACGT[rng.below(4)]);
			}
			evs.push_back(ev);
		}
		// reads: Poisson(coverage / L) starts per reference base; each read picks a haplotype; het events live on hap 1
		const double lam = std::exp(-P.coverage / (double)L); // exp(-lambda) of the per-base Poisson
		size_t first_read = D.reads.size();
		auto sim_range = [&](int64_t lo, int64_t hi, const Event *ev) {
			if (lo < 0) lo = 0;
			if (hi > P.chrom_len - L) hi = P.chrom_len - L;
			for (int64_t s = lo; s < hi; ++s) {
				if (ev && s >= ev->pos && s < ev->pos + ev->dlen) { // bases that exist only on the non-carrier haplotype
					int n = rng.poisson_exp(lam);
					for (int i = 0; i < n; ++i) { bool hap1 = rng.below(2) != 0; bool carries = ev->hom || hap1; if (!carries) make_read(D, rng, c, ref, ev, false, s, -1); }
					continue;
				}
				int n = rng.poisson_exp(lam);
				for (int i = 0; i < n; ++i) {
					bool hap1 = rng.below(2) != 0;
					bool carries = ev && (ev->hom || hap1);
					make_read(D, rng, c, ref, ev, carries, s, -1);
				}
				if (ev && s == ev->pos) // reads that start inside the inserted sequence (carrier haplotypes only)
					for (int k = 0; k < (int)ev->ins.size(); ++k) {
						int m = rng.poisson_exp(lam);
						for (int i = 0; i < m; ++i) { bool hap1 = rng.below(2) != 0; if (ev->hom || hap1) make_read(D, rng, c, ref, ev, true, s, k); }
					}
			}
		};
		if (P.locus_only) {
			for (const Event &ev : evs) sim_range(ev.pos - flank, ev.pos + ev.dlen + flank, &ev);
		} else {
			int64_t at = 0;
			for (const Event &ev : evs) {
				int64_t lo = ev.pos - 2 * L - 8; if (lo < at) lo = at;
				sim_range(at, lo, nullptr);
				int64_t hi = ev.pos + ev.dlen + 8;
				sim_range(lo, hi, &ev);
				at = hi;
			}
			sim_range(at, P.chrom_len, nullptr);
		}
		std::stable_sort(D.reads.begin() + (long)first_read, D.reads.end(), [](const ReadRec &a, const ReadRec &b) { return a.start < b.start; });
		D.events.insert(D.events.end(), evs.begin(), evs.end());
	}
	return Dp;
}

void idlh_dataset_free(idlh_dataset *d) { delete d; }

void idlh_dataset_counts(const idlh_dataset *d, int64_t counts[4])
{
	counts[0] = (int64_t)d->reads.size(); counts[1] = (int64_t)d->bases.size(); counts[2] = (int64_t)d->events.size(); counts[3] = (int64_t)d->chroms.size();
}

int64_t idlh_dataset_truth(const idlh_dataset *d, int64_t *out, int64_t cap)
{
	int64_t n = 0;
	for (const Event &e : d->events) {
		if (n >= cap) break;
		int64_t *o = out + 6 * n++;
		o[0] = e.chrom; o[1] = e.pos; o[2] = (int64_t)e.ins.size(); o[3] = e.dlen; o[4] = e.hom; o[5] = e.tr;
	}
	return (int64_t)d->events.size();
}

// ------------------------------------------------------------------------------------------------
// gen_roi (src/indelope.nim:515-545) with gen_roi_internal (:461-499), event_locations (:430-442),
// overlaps (:449-452), skippable (:40-47), cache_t (:502-513)
// ------------------------------------------------------------------------------------------------
idlh_rois *idlh_sweep(const idlh_dataset *d, int32_t min_event_support, int32_t min_read_coverage, int32_t max_read_coverage)
{
	idlh_rois *R = new idlh_rois();
	const size_t n = d->reads.size();
	R->start.resize(n); R->stop.resize(n); R->len.resize(n); R->mapq.resize(n); R->flag.resize(n); R->seq_off.resize(n);
	for (size_t i = 0; i < n; ++i) {
		const ReadRec &r = d->reads[i];
		R->start[i] = r.start; R->stop[i] = r.stop; R->len[i] = r.len; R->mapq[i] = r.mapq; R->flag[i] = r.flag; R->seq_off[i] = r.seq_off;
	}
	R->bases = d->bases.data(); R->quals = d->quals.data();
	const uint8_t min_evidence = (uint8_t)min_event_support;
	size_t ri = 0;
	for (size_t c = 0; c < d->chroms.size(); ++c) {
		const int64_t tlen = (int64_t)d->chroms[c].size();
		std::vector<uint8_t> evidence((size_t)tlen + 1, 0);           // :522
		std::vector<size_t> cache; int64_t cache_stop = 0;             // :523
		int64_t last_start = 0;                                         // :525
		const bool skip_chrom = idlh_skippable_chrom(d->names[c]);
		auto gen_roi_internal = [&](int64_t cache_start, int64_t cache_end) { // :461-499
			bool in_roi = false; int64_t roi_start = 0, roi_end = 0;
			auto flush = [&]() {
				std::vector<size_t> reads;
				for (size_t k : cache) {
					const ReadRec &r = d->reads[k];
					if (!(r.start > roi_end) && !(r.stop < roi_start)) { // overlaps :449-452
						reads.push_back(k);
						if ((int64_t)reads.size() > max_read_coverage) break;
					}
					if (r.start > roi_end) break;
				}
				if ((int64_t)reads.size() >= min_read_coverage && (int64_t)reads.size() <= max_read_coverage) {
					R->roi_chrom.push_back((int32_t)c); R->roi_start.push_back((int32_t)roi_start); R->roi_stop.push_back((int32_t)roi_end);
					R->roi_read_begin.push_back((int64_t)R->read_idx.size()); R->roi_n_reads.push_back((int32_t)reads.size());
					for (size_t k : reads) R->read_idx.push_back((int64_t)k);
				}
			};
			for (int64_t i = cache_start; i < cache_end; ++i) {
				if (evidence[(size_t)i] >= min_evidence) {
					if (!in_roi) { in_roi = true; roi_start = i; }
					roi_end = i;
					continue;
				}
				if (in_roi) { flush(); in_roi = false; }
			}
			if (in_roi) flush();
		};
		for (; ri < n && d->reads[ri].chrom == (int32_t)c; ++ri) { // b.querys(t.name) :527
			const ReadRec &r = d->reads[ri];
			if (!cache.empty() && r.start > cache_stop) { // :529-534
				gen_roi_internal(last_start, r.start);
				last_start = r.start;
				cache.clear(); cache_stop = 0;
			}
			if (skip_chrom || idlh_skippable_flag(r.flag)) continue; // skippable :40-47
			cache.push_back(ri); cache_stop = std::max<int64_t>(cache_stop, r.stop); // :504-506,537
			int64_t off = 0; // event_locations :430-442
			for (int32_t k = 0; k < r.n_cig; ++k) {
				uint32_t op = d->cigars[(size_t)r.cig_off + k] & 0xf, len = d->cigars[(size_t)r.cig_off + k] >> 4;
				bool cons = op == 0 || op == 2 || op == 3 || op == 7 || op == 8;
				if (op != 0) {
					int64_t es = r.start + off, ee = cons ? es + len : es + 1;
					for (int64_t i = es; i < ee && i <= tlen; ++i) { // :539-543, saturating
						uint8_t &ev = evidence[(size_t)i];
						ev += 1; if (ev == 0) ev = 255;
					}
				}
				if (cons) off += len;
			}
		}
		gen_roi_internal(last_start, (int64_t)evidence.size()); // :544
	}
	for (size_t c = 0; c < d->chroms.size(); ++c) {
		R->name_ptrs.push_back(d->names[c].c_str()); R->seq_ptrs.push_back(d->chroms[c].data()); R->chrom_len.push_back((int64_t)d->chroms[c].size());
	}
	idlh_roiset &v = R->view;
	v.n_reads = (int64_t)n; v.start = R->start.data(); v.stop = R->stop.data(); v.mapq = R->mapq.data(); v.flag = R->flag.data(); v.len = R->len.data();
	v.seq_off = R->seq_off.data(); v.bases = R->bases; v.quals = R->quals;
	v.n_rois = (int64_t)R->roi_start.size(); v.roi_chrom = R->roi_chrom.data(); v.roi_start = R->roi_start.data(); v.roi_stop = R->roi_stop.data();
	v.roi_read_begin = R->roi_read_begin.data(); v.roi_n_reads = R->roi_n_reads.data(); v.read_idx = R->read_idx.data();
	v.n_chroms = (int32_t)d->chroms.size(); v.chrom_name = R->name_ptrs.data(); v.chrom_seq = R->seq_ptrs.data(); v.chrom_len = R->chrom_len.data();
	return R;
}

idlh_chrom_reads *idlh_dataset_chrom(const idlh_dataset *d, int32_t chrom)
{
	if (chrom < 0 || (size_t)chrom >= d->chroms.size()) return nullptr;
	size_t a = 0, n = d->reads.size();
	while (a < n && d->reads[a].chrom < chrom) ++a;
	size_t b = a;
	while (b < n && d->reads[b].chrom == chrom) ++b;
	idlh_chrom_reads *c = (idlh_chrom_reads*)calloc(1, sizeof *c);
	const size_t m = b - a;
	size_t nc = 0;
	for (size_t i = a; i < b; ++i) nc += (size_t)d->reads[i].n_cig;
	c->first_read = (int64_t)a; c->n_reads = (int64_t)m; c->chrom_len = (int32_t)d->chroms[(size_t)chrom].size();
	c->start = (int32_t*)malloc((m + 1) * 4); c->stop = (int32_t*)malloc((m + 1) * 4); c->flag = (uint16_t*)malloc((m + 1) * 2);
	c->cigar = (uint32_t*)malloc((nc + 1) * 4); c->cig_off = (uint64_t*)malloc((m + 1) * 8);
	size_t k = 0;
	for (size_t i = 0; i < m; ++i) {
		const ReadRec &r = d->reads[a + i];
		c->start[i] = r.start; c->stop[i] = r.stop; c->flag[i] = r.flag; c->cig_off[i] = k;
		for (int32_t j = 0; j < r.n_cig; ++j) c->cigar[k++] = d->cigars[(size_t)r.cig_off + j];
	}
	c->cig_off[m] = k;
	return c;
}
void idlh_chrom_free(idlh_chrom_reads *c)
{
	if (!c) return;
	free(c->start); free(c->stop); free(c->flag); free(c->cigar); free(c->cig_off); free(c);
}

int32_t idlh_dataset_n_chroms(const idlh_dataset *d) { return (int32_t)d->chroms.size(); }
const char *idlh_dataset_chrom_name(const idlh_dataset *d, int32_t c) { return c >= 0 && (size_t)c < d->names.size() ? d->names[(size_t)c].c_str() : nullptr; }

void idlh_rois_free(idlh_rois *r) { delete r; }
const idlh_roiset *idlh_rois_view(const idlh_rois *r) { return &r->view; }

} // extern "C"
