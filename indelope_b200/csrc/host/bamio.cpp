// indelope_b200/csrc/host/bamio.cpp -- file I/O of the host stand-in, written from the SAM/BAM specification on zlib
// (no htslib / hts-nim in this image):
//
//   * FASTA + .fai reader  (the reference opens it with hts-nim's open_fai and slices it with fai.get,
//     src/indelope.nim:213-220,414,421,583) and a writer for the synthetic references
//   * BGZF + BAM reader    (the reference iterates `b.querys(target.name)` over a coordinate-sorted, indexed BAM,
//     src/indelope.nim:527,593-602; a sequential pass over a sorted BAM visits the same records in the same order)
//     with multi-threaded block inflation: the `-t/--threads` option of the reference's CLI (:566)
//   * BGZF + BAM writer    so that the synthetic configs exist as real .bam/.fa/.fai files
//
// hts-nim field semantics reproduced here: start = pos (0-based), stop = bam_endpos (pos + reference
// bases consumed, at least pos + 1), sequence() = the 4-bit codes through "=ACMGRSVTWYHKDBN" (soft clips included),
// base_qualities() = raw phred bytes, cigar ops as in the file.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>
#include "indelope_host.h"
#include "dataset.h"

namespace {

void set_err(char *err, size_t n, const std::string &msg) { if (err && n) { snprintf(err, n, "%s", msg.c_str()); } }

bool read_file(const char *path, std::vector<uint8_t> &out, std::string &why)
{
	FILE *f = fopen(path, "rb");
	if (!f) { why = std::string("cannot open ") + path; return false; }
	fseek(f, 0, SEEK_END);
	const long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	if (n < 0) { fclose(f); why = std::string("cannot size ") + path; return false; }
	out.resize((size_t)n);
	const size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
	fclose(f);
	if (got != (size_t)n) { why = std::string("short read on ") + path; return false; }
	return true;
}

inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }
inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline void put16(std::vector<uint8_t> &v, uint16_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }
inline void put32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

// ---------------------------------------------------------------------------------------------------------------
// BGZF (SAM spec 4.1): gzip members of at most 64 KiB with a 'BC' extra subfield holding the member size
// ---------------------------------------------------------------------------------------------------------------
struct BgzfBlock { size_t off, csize; size_t uoff, usize; };

bool bgzf_index(const std::vector<uint8_t> &file, std::vector<BgzfBlock> &blocks, size_t &total, std::string &why)
{
	size_t at = 0; total = 0;
	while (at < file.size()) {
		if (file.size() - at < 18) { why = "truncated BGZF header"; return false; }
		const uint8_t *h = file.data() + at;
		if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { why = "not a BGZF file (is it BAM? CRAM is not supported by this stand-in)"; return false; }
		const unsigned xlen = le16(h + 10);
		if (file.size() - at < 12 + xlen) { why = "truncated BGZF extra field"; return false; }
		int bsize = -1;
		for (unsigned x = 0; x + 4 <= xlen;) {
			const uint8_t *e = h + 12 + x; const unsigned slen = le16(e + 2);
			if (e[0] == 'B' && e[1] == 'C' && slen == 2) bsize = le16(e + 4);
			x += 4 + slen;
		}
		if (bsize < 0) { why = "BGZF block without BC subfield"; return false; }
		const size_t csize = (size_t)bsize + 1;
		if (csize < 12 + xlen + 8 || file.size() - at < csize) { why = "truncated BGZF block"; return false; }
		const size_t usize = le32(h + csize - 4);
		blocks.push_back({at, csize, total, usize});
		total += usize; at += csize;
	}
	return true;
}

bool bgzf_inflate_block(const std::vector<uint8_t> &file, const BgzfBlock &b, uint8_t *dst)
{
	if (b.usize == 0) return true;
	const uint8_t *h = file.data() + b.off;
	const unsigned xlen = le16(h + 10);
	z_stream zs; memset(&zs, 0, sizeof zs);
	if (inflateInit2(&zs, -15) != Z_OK) return false;
	zs.next_in = const_cast<Bytef*>(h + 12 + xlen); zs.avail_in = (uInt)(b.csize - 12 - xlen - 8);
	zs.next_out = dst; zs.avail_out = (uInt)b.usize;
	const int rc = inflate(&zs, Z_FINISH);
	inflateEnd(&zs);
	if (rc != Z_STREAM_END || zs.avail_out != 0) return false;
	return (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, (uInt)b.usize) == le32(h + b.csize - 8);
}

bool bgzf_read_all(const char *path, int threads, std::vector<uint8_t> &out, std::string &why)
{
	std::vector<uint8_t> file;
	if (!read_file(path, file, why)) return false;
	std::vector<BgzfBlock> blocks; size_t total = 0;
	if (!bgzf_index(file, blocks, total, why)) return false;
	out.resize(total);
	std::atomic<size_t> next(0); std::atomic<bool> ok(true);
	auto work = [&]() {
		for (;;) {
			const size_t i = next.fetch_add(16);
			if (i >= blocks.size() || !ok.load()) return;
			for (size_t k = i; k < std::min(i + 16, blocks.size()); ++k)
				if (!bgzf_inflate_block(file, blocks[k], out.data() + blocks[k].uoff)) { ok.store(false); return; }
		}
	};
	if (threads < 1) threads = 1;
	std::vector<std::thread> pool;
	for (int t = 1; t < threads; ++t) pool.emplace_back(work);
	work();
	for (auto &t : pool) t.join();
	if (!ok.load()) { why = "BGZF block failed to inflate (corrupt data or CRC mismatch)"; return false; }
	return true;
}

struct BgzfWriter {
	FILE *f = nullptr; std::vector<uint8_t> buf; int level = 1; bool ok = true;
	void flush_block(const uint8_t *p, size_t n)
	{
		std::vector<uint8_t> out(18 + compressBound((uLong)n) + 8);
		z_stream zs; memset(&zs, 0, sizeof zs);
		if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; return; }
		zs.next_in = const_cast<Bytef*>(p); zs.avail_in = (uInt)n;
		zs.next_out = out.data() + 18; zs.avail_out = (uInt)(out.size() - 18 - 8);
		const int rc = deflate(&zs, Z_FINISH);
		const size_t clen = zs.total_out;
		deflateEnd(&zs);
		if (rc != Z_STREAM_END) { ok = false; return; }
		const size_t bsize = 18 + clen + 8;
		static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
		memcpy(out.data(), hdr, 16);
		out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
		const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n);
		for (int i = 0; i < 4; ++i) { out[18 + clen + i] = (uint8_t)(crc >> (8 * i)); out[18 + clen + 4 + i] = (uint8_t)((uint32_t)n >> (8 * i)); }
		if (fwrite(out.data(), 1, bsize, f) != bsize) ok = false;
	}
	void write(const uint8_t *p, size_t n)
	{
		const size_t CAP = 0xff00; // uncompressed bytes per block, the value htslib uses
		while (n) {
			const size_t take = std::min(n, CAP - buf.size());
			buf.insert(buf.end(), p, p + take); p += take; n -= take;
			if (buf.size() == CAP) { flush_block(buf.data(), buf.size()); buf.clear(); }
		}
	}
	void close()
	{
		if (!buf.empty()) { flush_block(buf.data(), buf.size()); buf.clear(); }
		static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
		if (fwrite(eof, 1, 28, f) != 28) ok = false;
		if (fclose(f) != 0) ok = false;
		f = nullptr;
	}
};

int reg2bin(int64_t beg, int64_t end) // SAM spec 5.3
{
	--end;
	if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
	if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
	if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
	if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
	if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
	return 0;
}

const char SEQ16[] = "=ACMGRSVTWYHKDBN";
inline uint8_t code16(uint8_t c)
{
	switch (c) {
	case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'G': case 'g': return 4; case 'T': case 't': return 8;
	case '=': return 0; case 'M': return 3; case 'R': return 5; case 'S': return 6; case 'V': return 7; case 'W': return 9; case 'Y': return 10;
	case 'H': return 11; case 'K': return 12; case 'D': return 13; case 'B': return 14; default: return 15;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// FASTA
// ---------------------------------------------------------------------------------------------------------------
bool load_fasta(const char *path, idlh_dataset &D, std::string &why)
{
	std::vector<uint8_t> file;
	if (!read_file(path, file, why)) return false;
	size_t i = 0; const size_t n = file.size();
	while (i < n) {
		if (file[i] != '>') { why = std::string("FASTA record does not start with '>' in ") + path; return false; }
		size_t e = i + 1;
		while (e < n && file[e] != '\n') ++e;
		std::string name((const char*)file.data() + i + 1, e - i - 1);
		const size_t ws = name.find_first_of(" \t\r");
		if (ws != std::string::npos) name.resize(ws);
		i = e < n ? e + 1 : n;
		std::vector<uint8_t> seq;
		while (i < n && file[i] != '>') {
			size_t le = i;
			while (le < n && file[le] != '\n') ++le;
			size_t re = le;
			while (re > i && (file[re - 1] == '\r' || file[re - 1] == ' ')) --re;
			seq.insert(seq.end(), file.begin() + (long)i, file.begin() + (long)re);
			i = le < n ? le + 1 : n;
		}
		D.names.push_back(name); D.chroms.push_back(std::move(seq));
	}
	return true;
}

} // namespace

extern "C" {

int idlh_write_fasta(const idlh_dataset *d, const char *path)
{
	FILE *f = fopen(path, "wb");
	if (!f) return -1;
	const std::string fai_path = std::string(path) + ".fai";
	FILE *fi = fopen(fai_path.c_str(), "wb");
	if (!fi) { fclose(f); return -1; }
	const int W = 60;
	long long off = 0;
	for (size_t c = 0; c < d->chroms.size(); ++c) {
		off += fprintf(f, ">%s\n", d->names[c].c_str());
		const std::vector<uint8_t> &s = d->chroms[c];
		fprintf(fi, "%s\t%zu\t%lld\t%d\t%d\n", d->names[c].c_str(), s.size(), off, W, W + 1);
		for (size_t i = 0; i < s.size(); i += W) {
			const size_t k = std::min<size_t>(W, s.size() - i);
			fwrite(s.data() + i, 1, k, f); fputc('\n', f);
			off += (long long)k + 1;
		}
	}
	const bool ok = fclose(f) == 0;
	return (fclose(fi) == 0 && ok) ? 0 : -1;
}

/* coordinate-sorted BAM of the dataset's reads: @HD SO:coordinate, one @SQ per chromosome, read names r<serial>, no mates, no tags */
int idlh_write_bam(const idlh_dataset *d, const char *path, int level)
{
	BgzfWriter w;
	w.f = fopen(path, "wb");
	if (!w.f) return -1;
	w.level = level < 0 ? 1 : (level > 9 ? 9 : level);
	std::vector<uint8_t> b;
	std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
	for (size_t c = 0; c < d->chroms.size(); ++c) text += "@SQ\tSN:" + d->names[c] + "\tLN:" + std::to_string(d->chroms[c].size()) + "\n";
	text += "@PG\tID:indelope_b200\tPN:indelope_b200\n";
	b.insert(b.end(), {'B', 'A', 'M', 1});
	put32(b, (uint32_t)text.size()); b.insert(b.end(), text.begin(), text.end());
	put32(b, (uint32_t)d->chroms.size());
	for (size_t c = 0; c < d->chroms.size(); ++c) {
		put32(b, (uint32_t)d->names[c].size() + 1); b.insert(b.end(), d->names[c].begin(), d->names[c].end()); b.push_back(0);
		put32(b, (uint32_t)d->chroms[c].size());
	}
	w.write(b.data(), b.size());
	for (const IdlhReadRec &r : d->reads) {
		b.clear();
		char name[32];
		const int ln = snprintf(name, sizeof name, "r%llu", (unsigned long long)r.order) + 1;
		const uint32_t block = 32 + (uint32_t)ln + 4u * (uint32_t)r.n_cig + (uint32_t)(r.len + 1) / 2 + (uint32_t)r.len;
		put32(b, block); put32(b, (uint32_t)r.chrom); put32(b, (uint32_t)r.start);
		b.push_back((uint8_t)ln); b.push_back(r.mapq); put16(b, (uint16_t)reg2bin(r.start, r.stop > r.start ? r.stop : r.start + 1));
		put16(b, (uint16_t)r.n_cig); put16(b, r.flag); put32(b, (uint32_t)r.len);
		put32(b, 0xffffffffu); put32(b, 0xffffffffu); put32(b, 0);
		b.insert(b.end(), name, name + ln);
		for (int k = 0; k < r.n_cig; ++k) put32(b, d->cigars[(size_t)r.cig_off + k]);
		const uint8_t *s = d->bases.data() + r.seq_off, *q = d->quals.data() + r.seq_off;
		for (int i = 0; i < r.len; i += 2) b.push_back((uint8_t)(code16(s[i]) << 4 | (i + 1 < r.len ? code16(s[i + 1]) : 0)));
		b.insert(b.end(), q, q + r.len);
		w.write(b.data(), b.size());
	}
	w.close();
	return w.ok ? 0 : -1;
}

/* reference FASTA + coordinate-sorted BAM -> dataset.  Records without a reference id are dropped (a per-target query never
 * returns them); everything else, flags included, is kept for the sweep's `skippable` test (src/indelope.nim:40-47). */
idlh_dataset *idlh_load(const char *fasta_path, const char *bam_path, int threads, char *err, size_t errlen)
{
	idlh_dataset *D = new idlh_dataset();
	memset(&D->P, 0, sizeof D->P);
	std::string why;
	auto fail = [&](const std::string &m) -> idlh_dataset* { set_err(err, errlen, m); delete D; return nullptr; };
	if (!load_fasta(fasta_path, *D, why)) return fail(why);
	std::vector<uint8_t> bam;
	if (!bgzf_read_all(bam_path, threads, bam, why)) return fail(std::string(bam_path) + ": " + why);
	const uint8_t *p = bam.data(); const size_t n = bam.size();
	if (n < 12 || memcmp(p, "BAM\1", 4) != 0) return fail(std::string(bam_path) + ": not a BAM file");
	size_t at = 4;
	const uint32_t l_text = le32(p + at); at += 4;
	if (at + l_text + 4 > n) return fail("truncated BAM header");
	at += l_text;
	const uint32_t n_ref = le32(p + at); at += 4;
	// map BAM reference ids onto FASTA records by name; the VCF header and the sweep follow the BAM's target order (:599-601)
	std::vector<std::string> names; std::vector<std::vector<uint8_t>> chroms;
	for (uint32_t r = 0; r < n_ref; ++r) {
		if (at + 4 > n) return fail("truncated BAM reference list");
		const uint32_t l_name = le32(p + at); at += 4;
		if (at + l_name + 4 > n || l_name == 0) return fail("truncated BAM reference list");
		const std::string name((const char*)p + at, l_name - 1); at += l_name;
		const uint32_t l_ref = le32(p + at); at += 4;
		size_t k = 0;
		while (k < D->names.size() && D->names[k] != name) ++k;
		if (k == D->names.size()) return fail("BAM target " + name + " is not in the FASTA");
		if (D->chroms[k].size() != l_ref) return fail("BAM target " + name + " has a different length than the FASTA record");
		names.push_back(name); chroms.push_back(D->chroms[k]);
	}
	D->names.swap(names); D->chroms.swap(chroms);
	int32_t last_ref = 0, last_pos = -1;
	while (at + 4 <= n) {
		const uint32_t block = le32(p + at); at += 4;
		if (block < 32 || at + block > n) return fail("truncated BAM record");
		const uint8_t *b = p + at; at += block;
		const int32_t ref_id = (int32_t)le32(b), pos = (int32_t)le32(b + 4);
		const unsigned l_name = b[8]; const uint8_t mapq = b[9];
		const unsigned n_cig = le16(b + 12); const uint16_t flag = le16(b + 14);
		const uint32_t l_seq = le32(b + 16);
		if (32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq > block) return fail("malformed BAM record");
		if (ref_id < 0) continue;
		if ((uint32_t)ref_id >= n_ref) return fail("BAM record with an unknown reference id");
		if (ref_id < last_ref || (ref_id == last_ref && pos < last_pos)) return fail("BAM is not coordinate sorted");
		last_ref = ref_id; last_pos = pos;
		IdlhReadRec r;
		r.chrom = ref_id; r.start = pos; r.mapq = mapq; r.flag = flag; r.len = (int32_t)l_seq;
		r.seq_off = (int64_t)D->bases.size(); r.cig_off = (int64_t)D->cigars.size(); r.n_cig = (int32_t)n_cig; r.order = D->reads.size();
		const uint8_t *cg = b + 32 + l_name;
		int64_t rlen = 0;
		for (unsigned k = 0; k < n_cig; ++k) {
			const uint32_t c = le32(cg + 4 * k); const unsigned op = c & 0xf;
			D->cigars.push_back(c);
			if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;
		}
		if ((flag & 4) || n_cig == 0 || rlen == 0) rlen = 1; // bam_endpos
		r.stop = (int32_t)(pos + rlen);
		const uint8_t *sq = cg + 4 * n_cig, *ql = sq + (l_seq + 1) / 2;
		for (uint32_t i = 0; i < l_seq; ++i) D->bases.push_back((uint8_t)SEQ16[(sq[i >> 1] >> ((~i & 1) << 2)) & 0xf]);
		D->quals.insert(D->quals.end(), ql, ql + l_seq);
		D->reads.push_back(r);
	}
	if (at != n) return fail("trailing bytes after the last BAM record");
	return D;
}

} // extern "C"
